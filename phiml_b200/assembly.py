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
    """Returns the stencil description of a constant-coefficient ZERO-boundary ``ShiftLinTracer`` or None."""
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
    if len(tracer.val) > MAX_POINTS or any(v.default_backend is not NUMPY for v in tracer.val.values()):
        return None
    grid = tuple(int(s) for s in out_shape.sizes)
    offsets, coeffs = [], []
    dtype = None
    for shift_, shift_val in tracer.val.items():
        assert isinstance(shift_, Shift)
        if any(d not in out_shape.names for d in shift_.by_dim):
            return None
        off = tuple(int(shift_[d]) for d in out_shape.names)
        v = np.asarray(shift_val.native([*out_shape]))
        if v.shape != grid or v.dtype.kind != 'f':
            return None
        dtype = v.dtype if dtype is None else np.promote_types(dtype, v.dtype)
        box = tuple(slice(max(0, -o), n - max(0, o)) for o, n in zip(off, grid))
        inner = v[box]
        if inner.size == 0:
            if np.any(v):
                return None
            continue
        c = inner.flat[0]
        if not np.all(inner == c):
            return None      # variable coefficients
        nz = int(np.count_nonzero(v))
        if c == 0:
            if nz:
                return None
            continue          # the reference drops all-zero shifts as well (np.nonzero, _lin_trace.py:774-776)
        if nz != inner.size:
            return None      # non-zero values at wrapped positions: periodic / other boundaries
        offsets.append(off)
        coeffs.append(float(c))
    if not offsets or dtype not in (np.float32, np.float64):
        return None
    if any(v.dtype != dtype for v in (np.asarray(t.native([*out_shape])) for t in tracer.val.values())):
        return None
    return dict(grid=grid, offsets=offsets, coeffs=coeffs, dtype=np.dtype(dtype), out_shape=out_shape, src_shape=src_shape,
                m_rank=out_shape.volume - tracer.min_rank_deficiency())


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


STATS = {'device_assemblies': 0, 'fallbacks': 0}


def make_tracer_to_coo(original: Callable, eligible: Callable[[], bool], assemble: Callable = cuda_assemble):
    """Wrapper for ``phiml.math._lin_trace.tracer_to_coo``.  ``eligible()`` says whether the matrix is wanted on the CUDA torch backend
    right now (the wrapped function is not told the target backend); ``assemble`` is injectable so that the CPU tests can check the
    PhiML-side structure against the reference's own matrix with arrays from the oracle."""
    def tracer_to_coo(tracer, sparsify, separate_independent):
        info = None
        if eligible():
            info = analyse_tracer(tracer, sparsify, separate_independent)
        if info is not None:
            natives = assemble(info['grid'], info['offsets'], info['coeffs'], info['dtype'])
            if natives is not None:
                STATS['device_assemblies'] += 1
                return build_compressed(info, natives), tracer._bias
        STATS['fallbacks'] += 1
        return original(tracer, sparsify, separate_independent)
    return tracer_to_coo
