"""
Bridge from PhiML's linear tracer to the device assembler.

``math.matrix_from_function`` traces a linear function into a ``ShiftLinTracer`` — one full-grid value array per stencil shift
(phiml/math/_lin_trace.py:57-168) — and then materialises it on the HOST: ``tracer_to_coo`` (:750-807) builds int64 coordinate
arrays for every stored entry, ``SparseCoordinateTensor.compress`` (phiml/math/_sparse.py:347-390) sorts them through SciPy.  That is
36.7 s / 17.5 GB at 256^3 and impossible from 512^3 on.  For what the BASELINE configs trace — constant
coefficients with ZERO boundaries (``laplace`` / central differences with ``extrapolation.ZERO``) — the tracer already says
everything: a handful of (offset, coefficient) pairs.  This module

* recognises that case on the tracer (``analyse_tracer``: every shift's value array must equal ONE constant on the cells whose
  shifted neighbour lies inside the grid and exactly 0 elsewhere; anything else — variable coefficients, periodic wrap, sliced or
  renamed dims, batched values — is left to the reference), and
* builds the SAME ``CompressedSparseMatrix`` the reference would return (same dims, int32 CSR with sorted columns, byte-identical
  arrays — ``phb200_stencil_to_csr``, pinned by tests/test_gpu_parity.py::test_device_assembly_bit_exact) with its index / pointer /
  value arrays living on the device, so nothing of size nnz ever exists on the host.

``phiml_plugin.install()`` re-binds ``phiml.math._lin_trace.tracer_to_coo`` to ``make_tracer_to_coo`` below; ``matrix_from_function``
then skips its own ``compress_rows()`` because the result is already compressed (:741-742).
"""
from typing import Callable, Optional

import numpy as np

MAX_POINTS = 16


def analyse_tracer(tracer, sparsify, separate_independent) -> Optional[dict]:
    """Returns the description of a ``ShiftLinTracer`` the device can assemble, or None.  Constant coefficients with ZERO boundaries give a
    stencil description (``offsets`` / ``coeffs``); anything else with one value per cell and shift — variable coefficients, PERIODIC wrap,
    non-zero edge values — gives the general description (``general=True``, ``values``: one full-grid array per shift)."""
    from phiml.backend import NUMPY
    from phiml.math._lin_trace import ShiftLinTracer, Shift, matrix_dims_for_tracer, pattern_dim_names
    from phiml.math._shape import batch, merge_shapes, non_batch
    if separate_independent or not isinstance(tracer, ShiftLinTracer):
        return None
    try:
        pattern_dims = tracer._source.shape.only(pattern_dim_names(tracer))
        if not batch(pattern_dims).is_empty:
            return None
        out_shape, src_shape, typed_src_shape, missing_dims, sliced_src_shape = matrix_dims_for_tracer(tracer, sparsify)
    except Exception:
        return None
    if missing_dims or non_batch(out_shape).is_empty or out_shape.rank > 3 or out_shape.rank < 1:
        return None
    if any(tracer._out_name_to_original.get(d, d) != d for d in out_shape.names):
        return None          # renamed dims
    if out_shape.names != typed_src_shape.names or out_shape.sizes != typed_src_shape.sizes:
        return None
    if not merge_shapes(*tracer.val.values()).without(out_shape).is_empty:
        return None          # batched values
    if any(v.default_backend is not NUMPY for v in tracer.val.values()):
        return None
    grid = tuple(int(s) for s in out_shape.sizes)
    offsets, coeffs = [], []
    all_offsets, all_values = [], []
    constant = True          # constant coefficients, ZERO boundaries (the stencil form)
    dtype = None
    for shift_, shift_val in tracer.val.items():
        assert isinstance(shift_, Shift)
        if any(d not in out_shape.names for d in shift_.by_dim):
            return None
        off = tuple(int(shift_[d]) for d in out_shape.names)
        v = np.asarray(shift_val.native([*out_shape]))
        if v.shape != grid or v.dtype.kind != 'f':
            return None
        if dtype is not None and v.dtype != dtype:
            return None
        dtype = v.dtype
        if not np.any(v):
            continue          # all-zero shifts leave no entries (np.nonzero, _lin_trace.py:774-776); ZERO_GRADIENT traces carry many of them
        all_offsets.append(off)
        all_values.append(v)
        if len(all_offsets) > MAX_POINTS:
            return None
        if not constant:
            continue
        box = tuple(slice(max(0, -o), n - max(0, o)) for o, n in zip(off, grid))
        inner = v[box]
        if inner.size == 0:
            constant = False
            continue
        c = inner.flat[0]
        nz = int(np.count_nonzero(v))
        if not np.all(inner == c) or (c == 0 and nz) or (c != 0 and nz != inner.size):
            constant = False  # variable coefficients / non-zero values at wrapped positions (periodic and other boundaries)
            continue
        offsets.append(off)
        coeffs.append(float(c))
    if dtype not in (np.float32, np.float64):
        return None
    common = dict(grid=grid, dtype=np.dtype(dtype), out_shape=out_shape, src_shape=src_shape, m_rank=out_shape.volume - tracer.min_rank_deficiency())
    if not all_offsets:
        return None
    if constant:
        return dict(common, general=False, offsets=offsets, coeffs=coeffs)
    if any(abs(o) >= n for off in all_offsets for o, n in zip(off, grid)):
        return None           # a shift as long as the grid wraps onto its own column
    return dict(common, general=True, offsets=all_offsets, values=all_values)


def cuda_assemble(grid, offsets, coeffs, dtype: np.dtype, device=None):
    """(rowptr int32, col int32, values) on the device, byte-identical to the reference's compress_rows() (phb200_stencil_to_csr)."""
    import torch
    from .sparse import Matrix
    dev = torch.device(device if device is not None else 'cuda')
    st = Matrix.stencil_matrix(grid, offsets, coeffs, torch.float32 if dtype == np.float32 else torch.float64, dev)
    if st.nnz() >= 2 ** 31:
        return None          # int32 CSR (the reference's native) cannot hold it
    csr = st.to_csr()
    return csr.rowptr, csr.col, csr.values


def cuda_assemble_general(grid, offsets, values, dtype: np.dtype, device=None):
    """(rowptr int32, col int32, values) on the device for a ``ShiftLinTracer`` with one value per cell and shift (variable coefficients, PERIODIC wrap):
    the per-shift arrays are uploaded once ((shifts, N) values) and ``phb200_shift_count`` / ``phb200_shift_fill`` do what ``tracer_to_coo`` +
    ``compress_rows()`` do on the host — drop the zeros, wrap the columns, order every row by column — byte-identical, without any array of size nnz
    on the host.  None if no int32 CSR exists (the caller then falls back to the reference)."""
    import ctypes
    import torch
    from . import _engine as E
    dev = torch.device(device if device is not None else 'cuda')
    if dev.index is None:
        dev = torch.device('cuda', torch.cuda.current_device())
    n = int(np.prod(grid))
    rank = len(grid)
    tdt = torch.float32 if dtype == np.float32 else torch.float64
    with torch.cuda.device(dev):
        vals = torch.empty((len(offsets), n), dtype=tdt, device=dev)
        for k, v in enumerate(values):                      # shift by shift: no second full copy on the host
            vals[k].copy_(torch.as_tensor(np.ascontiguousarray(v).reshape(-1)))
        g = (ctypes.c_int64 * rank)(*[int(s) for s in grid])
        flat = [int(o) for off in offsets for o in off]
        sh = (ctypes.c_int32 * len(flat))(*flat)
        rowptr = torch.empty(n + 1, dtype=torch.int32, device=dev)
        nnz = ctypes.c_int64()
        try:
            E.check(E.lib().phb200_shift_count(rank, g, len(offsets), sh, E.ptr(vals), E.dtype_code(tdt), E.ptr(rowptr), ctypes.byref(nnz), E.stream_ptr(dev)))
            col = torch.empty(nnz.value, dtype=torch.int32, device=dev)
            out = torch.empty(nnz.value, dtype=tdt, device=dev)
            E.check(E.lib().phb200_shift_fill(rank, g, len(offsets), sh, E.ptr(vals), E.dtype_code(tdt), E.ptr(rowptr), E.ptr(col), E.ptr(out), E.stream_ptr(dev)))
        except (ValueError, NotImplementedError):
            return None
    return rowptr, col, out


def build_compressed(info: dict, natives):
    """The ``CompressedSparseMatrix`` of ``SparseCoordinateTensor.compress_rows()`` (phiml/math/_sparse.py:383-390) around device arrays."""
    from phiml.math._shape import instance
    from phiml.math._sparse import CompressedSparseMatrix
    from phiml.math._tensors import wrap
    from phiml.math._ops import reshaped_tensor
    rowptr, col, values = natives
    indices = wrap(col, instance('sp_entries'))
    pointers = wrap(rowptr, instance('pointers'))
    vals = reshaped_tensor(values, [instance('sp_entries')], convert=False)
    m = CompressedSparseMatrix(indices, pointers, vals, info['src_shape'], info['out_shape'], True, uncompressed_indices=None,
                               uncompressed_indices_perm=None, m_rank=info['m_rank'])
    m._phb200_device_assembled = True      # a constant, like the reference's NumPy-backed matrix (phiml_plugin.attach_gradient_solve)
    register_owner(m, values)
    return m


def coordinates_of_compressed(m):
    """The coordinates ``SparseCoordinateTensor.compress_rows()`` would have kept for the compressed matrix ``m`` (phiml/math/_sparse.py:383-390), rebuilt from its
    CSR arrays: one row of per-dimension indices per stored entry, in CSR order, labels = uncompressed (dual) names followed by compressed (primal) names."""
    import torch
    from phiml.math._shape import instance, channel
    from phiml.math._ops import reshaped_tensor
    col = torch.as_tensor(m._indices.native('sp_entries')).to(torch.int64)
    ptr = torch.as_tensor(m._pointers.native('pointers')).to(torch.int64).to(col.device)
    rows = torch.repeat_interleave(torch.arange(ptr.numel() - 1, device=col.device, dtype=torch.int64), ptr[1:] - ptr[:-1])
    parts = []
    for flat, dims in ((col, m._uncompressed_dims), (rows, m._compressed_dims)):
        comps, rem = [], flat
        for size in reversed(dims.sizes):
            comps.append(rem % size)
            rem = rem // size
        parts += comps[::-1]
    native = torch.stack(parts, -1).cpu().numpy()      # NumPy-backed like every coordinate array of the reference (its packing / slicing utilities insist: _sparse.py:394)
    labels = m._uncompressed_dims.names + m._compressed_dims.names
    return reshaped_tensor(native, [instance('sp_entries'), channel(sparse_idx=labels)], convert=False)


# PhiML hands constants around as detached copies (stop_gradient in key_from_args, phiml/math/_functional.py:147): new tensor objects over
# the SAME storage on every call.  The residency caches of the plug-in need one stable object per matrix to tie their entries to:
# the bridged matrix itself (kept alive by the LinearFunction that traced it), found again through the address of its value array.
_OWNERS = {}


def register_owner(matrix, values):
    import weakref
    key = int(values.data_ptr())
    _OWNERS[key] = weakref.ref(matrix, lambda _r, k=key: _OWNERS.pop(k, None))


def owner_of(values):
    """The live bridged matrix whose value array shares storage (address, version, size) with ``values``, or None."""
    try:
        ref = _OWNERS.get(int(values.data_ptr()))
    except Exception:
        return None
    m = ref() if ref is not None else None
    if m is None:
        return None
    own = m._values._native
    return m if (own.data_ptr() == values.data_ptr() and own._version == values._version and own.numel() == values.numel() and own.dtype == values.dtype) else None


STATS = {'device_assemblies': 0, 'general_assemblies': 0, 'fallbacks': 0}


def _caller_target_backend():
    """``matrix_from_function`` determines the backend its matrix is FOR from the tensor it traces (``target_backend = backend_for(*tensors)``,
    phiml/math/_lin_trace.py:714) — NumPy-backed inputs are solved on the NumPy backend even while torch is the default one.  ``tracer_to_coo`` is not
    told; read from the calling frame like ``auto_compress`` below (``None`` for any other caller)."""
    import sys
    try:
        frame = sys._getframe(2)
        if frame.f_code.co_name == 'matrix_from_function':
            return frame.f_locals.get('target_backend')
    except ValueError:
        pass
    return None


def _caller_wants_compressed() -> bool:
    """``matrix_from_function(..., auto_compress=False)`` — the default of the public function — returns the coordinate form (SparseCoordinateTensor,
    phiml/math/_lin_trace.py:680,743); only ``auto_compress=True`` (``jit_compile_linear`` / ``solve_linear``, phiml/math/_functional.py:436) asks for the
    compressed rows the device assembler produces.  ``tracer_to_coo`` is not told, so the flag is read from the calling frame (two frames up from here);
    any other caller gets the compressed form."""
    import sys
    try:
        frame = sys._getframe(2)
        if frame.f_code.co_name == 'matrix_from_function':
            return bool(frame.f_locals.get('auto_compress', True))
    except ValueError:
        pass
    return True


def make_tracer_to_coo(original: Callable, eligible: Callable[[], bool], assemble: Callable = cuda_assemble, assemble_general: Callable = cuda_assemble_general):
    """Wrapper for ``phiml.math._lin_trace.tracer_to_coo``.  ``eligible([target_backend])`` says whether the matrix is wanted on the CUDA torch backend
    right now (the wrapped function is not told the target backend: it is read from ``matrix_from_function``'s frame and handed over when ``eligible`` takes an argument); ``assemble`` is injectable so that the CPU tests can check the
    PhiML-side structure against the reference's own matrix with arrays from the oracle."""
    import inspect
    takes_backend = len(inspect.signature(eligible).parameters) >= 1

    def tracer_to_coo(tracer, sparsify, separate_independent):
        info = None
        wanted = eligible(_caller_target_backend()) if takes_backend else eligible()
        if wanted and _caller_wants_compressed():
            info = analyse_tracer(tracer, sparsify, separate_independent)
        if info is not None:
            if info['general']:
                natives = assemble_general(info['grid'], info['offsets'], info['values'], info['dtype'])
            else:
                natives = assemble(info['grid'], info['offsets'], info['coeffs'], info['dtype'])
            if natives is not None:
                STATS['general_assemblies' if info['general'] else 'device_assemblies'] += 1
                return build_compressed(info, natives), tracer._bias
        STATS['fallbacks'] += 1
        return original(tracer, sparsify, separate_independent)
    return tracer_to_coo
