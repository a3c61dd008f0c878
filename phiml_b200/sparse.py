"""
Device matrices for the phb200 engine.

Natives at the boundary are what the reference's torch backend produces
(phiml/backend/torch/_torch_backend.py:857-896): ``torch.sparse_csr_tensor`` (crow int32/int64, col, values),
``torch.sparse_csc_tensor`` (backward pass: CSR arrays of A^T) and ``torch.sparse_coo_tensor`` (int64 indices).
``as_matrix`` wraps them without copying values; a matrix that turns out to be a constant-coefficient 5/7-point
(or any <=16 point) ZERO-boundary stencil — what ``matrix_from_function`` emits for Laplace / central differences —
is additionally recognised on the device and then runs through the stencil-specialised kernels.
"""
import ctypes
from collections import OrderedDict
from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from . import _engine as E


class Matrix:
    """Owner of a ``phb200_matrix`` handle; keeps the tensors it views alive."""

    def __init__(self, handle, dtype: torch.dtype, shape: Tuple[int, int], keep, kind: str, device):
        self._h = handle
        self.dtype = dtype
        self.shape = tuple(int(s) for s in shape)
        self._keep = keep
        self.kind = kind          # 'csr' | 'csc' | 'stencil' | 'coded'
        self.device = device
        self.stencil: Optional['Matrix'] = None   # detected stencil twin of a CSR matrix
        self.coded: Optional['Matrix'] = None     # pattern-coded twin (variable coefficients / non-ZERO boundaries), products only
        self._stencil_checked = False
        self._coded_checked = False
        self.grid = None

    @property
    def handle(self):
        return self._h

    def __del__(self):
        try:
            if self._h is not None and E._lib is not None:
                E._lib.phb200_matrix_destroy(self._h)
        except Exception:
            pass
        self._h = None

    # --- constructors ---
    @staticmethod
    def csr(rowptr: torch.Tensor, col: torch.Tensor, values: torch.Tensor, shape, transposed=False) -> 'Matrix':
        E.require_cuda(rowptr, col, values)
        rowptr = rowptr.to(torch.int32).contiguous()
        col = col.to(torch.int32).contiguous()
        values = values.contiguous()
        rows_stored = shape[1] if transposed else shape[0]
        cols_stored = shape[0] if transposed else shape[1]
        assert rowptr.numel() == rows_stored + 1, f"row pointer length {rowptr.numel()} does not match {rows_stored} rows"
        h = ctypes.c_void_p()
        E.check(E.lib().phb200_matrix_csr(ctypes.byref(h), E.ptr(rowptr), E.ptr(col), E.ptr(values), rows_stored, cols_stored,
                                          values.numel(), E.dtype_code(values.dtype), 1 if transposed else 0))
        m = Matrix(h, values.dtype, shape, (rowptr, col, values), 'csc' if transposed else 'csr', values.device)
        m.rowptr, m.col, m.values = rowptr, col, values
        return m

    @staticmethod
    def stencil_matrix(grid: Sequence[int], offsets: Sequence[Sequence[int]], coeffs: Sequence[float], dtype: torch.dtype, device='cuda') -> 'Matrix':
        """Constant-coefficient stencil, ZERO boundaries.  ``offsets[p]`` = source - output cell, one int per grid dim."""
        rank = len(grid)
        g = (ctypes.c_int64 * rank)(*[int(v) for v in grid])
        flat = [int(o) for off in offsets for o in off]
        o = (ctypes.c_int32 * len(flat))(*flat)
        c = (ctypes.c_double * len(coeffs))(*[float(v) for v in coeffs])
        h = ctypes.c_void_p()
        E.check(E.lib().phb200_matrix_stencil(ctypes.byref(h), rank, g, len(coeffs), o, c, E.dtype_code(dtype)))
        n = int(np.prod(grid))
        m = Matrix(h, dtype, (n, n), (), 'stencil', torch.device(device))
        m.grid = tuple(int(v) for v in grid)
        m.spec = ([tuple(int(o) for o in off) for off in offsets], [float(v) for v in coeffs])
        return m

    # --- stencil recognition ---
    @E.on_operand_device
    def detect_stencil(self, grid: Optional[Sequence[int]] = None) -> Optional['Matrix']:
        """Returns the stencil twin if every row matches bit for bit, else None.  ``grid`` defaults to an inference
        from the neighbour offsets of a few probe rows (nearest-neighbour stencils on C-ordered grids)."""
        if self.kind != 'csr' or self.shape[0] != self.shape[1] or self.shape[0] == 0:
            return None
        if self._stencil_checked and grid is None:
            return self.stencil
        if grid is None:
            grid = _infer_grid(self)
            self._stencil_checked = True
        if grid is None:
            return None
        g = (ctypes.c_int64 * len(grid))(*[int(v) for v in grid])
        h = ctypes.c_void_p()
        try:
            E.check(E.lib().phb200_stencil_from_csr(ctypes.byref(h), self._h, len(grid), g, E.stream_ptr(self.device)))
        except E.NotAStencil:
            return None
        st = Matrix(h, self.dtype, self.shape, (), 'stencil', self.device)
        st.grid = tuple(int(v) for v in grid)
        self.stencil = st
        return st

    @E.on_operand_device
    def detect_coded(self) -> Optional['Matrix']:
        """Pattern-coded twin (csrc/coded.cuh) of a CSR matrix that is NOT a constant-coefficient ZERO-boundary stencil but still has few
        distinct rows — ZERO_GRADIENT / PERIODIC boundaries, masks, piecewise-constant coefficients: one code byte per row instead of the
        matrix stream, products bit-identical to the CSR product.  None if the matrix does not qualify (``PHB200_CODED=0`` switches it off)."""
        import os
        if self.kind != 'csr' or self.shape[0] != self.shape[1] or self.shape[0] == 0 or os.environ.get('PHB200_CODED', '1') == '0':
            return None
        if self._coded_checked:
            return self.coded
        self._coded_checked = True
        h = ctypes.c_void_p()
        try:
            E.check(E.lib().phb200_coded_from_csr(ctypes.byref(h), self._h, E.stream_ptr(self.device)))
        except E.NotAStencil:
            return None
        tw = Matrix(h, self.dtype, self.shape, (self.rowptr, self.col, self.values), 'coded', self.device)     # views this matrix' CSR arrays
        k, p, sd = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        offs = (ctypes.c_int64 * 16)()
        E.check(E.lib().phb200_coded_info(h, ctypes.byref(k), ctypes.byref(p), offs, ctypes.byref(sd)))
        tw.offsets, tw.patterns = [int(offs[i]) for i in range(k.value)], int(p.value)
        tw.star_diag = bool(sd.value)      # star with a per-row diagonal: CG / CG-adaptive run on the fused TMA kernels
        self.coded = tw
        return tw

    @property
    def product_twin(self) -> 'Matrix':
        """The handle products and Krylov loops run on: stencil twin, else coded twin, else the matrix itself."""
        return self.stencil if self.stencil is not None else (self.coded if self.coded is not None else self)

    def nnz(self) -> int:
        n = ctypes.c_int64()
        E.check(E.lib().phb200_matrix_info(self._h, None, None, None, ctypes.byref(n), None))
        return int(n.value)

    @E.on_operand_device
    def to_csr(self) -> 'Matrix':
        """Stencil -> canonical CSR assembled on the device (byte-identical to the reference's compress())."""
        assert self.kind == 'stencil'
        n = self.shape[0]
        nnz = self.nnz()
        rowptr = torch.empty(n + 1, dtype=torch.int32, device=self.device)
        col = torch.empty(nnz, dtype=torch.int32, device=self.device)
        val = torch.empty(nnz, dtype=self.dtype, device=self.device)
        E.check(E.lib().phb200_stencil_to_csr(self._h, E.ptr(rowptr), E.ptr(col), E.ptr(val), E.stream_ptr(self.device)))
        return Matrix.csr(rowptr, col, val, self.shape)

    # --- product ---
    @E.on_operand_device
    def matvec(self, x: torch.Tensor, use_stencil=True) -> torch.Tensor:
        """(batch, cols) -> (batch, rows)"""
        E.require_cuda(x)
        assert x.dim() == 2 and x.shape[1] == self.shape[1], f"expected (batch, {self.shape[1]}), got {tuple(x.shape)}"
        x = x.to(self.dtype).contiguous()
        m = self.product_twin if use_stencil else self
        y = torch.empty((x.shape[0], self.shape[0]), dtype=self.dtype, device=x.device)
        for b0 in range(0, x.shape[0], 65535):
            xb, yb = x[b0:b0 + 65535], y[b0:b0 + 65535]
            E.check(E.lib().phb200_spmv(m._h, E.ptr(xb), E.ptr(yb), xb.shape[0], x.shape[1], y.shape[1], E.stream_ptr(x.device)))
        return y


def _infer_grid(m: Matrix) -> Optional[Tuple[int, ...]]:
    """Grid shape from the positive neighbour offsets {1, nz, ny*nz} seen in a few rows around the middle."""
    n = m.shape[0]
    probes = sorted({n // 2, n // 3, min(n - 1, n // 2 + 1), 0})
    rp = m.rowptr[torch.tensor([p for r in probes for p in (r, r + 1)], device=m.device)].cpu().numpy().reshape(-1, 2)
    offs = set()
    for r, (a, b) in zip(probes, rp):
        if b - a > E_MAX_STENCIL:
            return None
        cols = m.col[a:b].cpu().numpy().astype(np.int64)
        offs.update(int(c - r) for c in cols if c > r)
    offs = sorted(offs)
    if len(offs) == 0:
        return (n,)
    if len(offs) > 3 or offs[0] != 1:
        return None
    dims = []
    prev = 1
    for o in offs[1:]:
        if o % prev != 0:
            return None
        dims.append(o // prev)
        prev = o
    if n % prev != 0:
        return None
    grid = tuple([n // prev] + dims[::-1])
    return grid if int(np.prod(grid)) == n else None


E_MAX_STENCIL = 16


# ----------------------------------------------------------------------------------------------------------------------
# natives -> Matrix, with a tiny identity cache so that repeated products / solves with one native do not repeat the
# analysis (the cache holds references, so addresses cannot be recycled under it; in-place edits bump _version)
# ----------------------------------------------------------------------------------------------------------------------

_CACHE: 'OrderedDict[tuple, Matrix]' = OrderedDict()
_CACHE_SIZE = 4


def _cache_key(kind, a, b, c, shape):
    return (kind, a.data_ptr(), b.data_ptr(), c.data_ptr(), a._version, b._version, c._version, tuple(shape), c.dtype, c.numel())


def clear_cache():
    _CACHE.clear()


@E.on_operand_device
def as_matrix(native, detect_stencil=True, grid=None) -> Matrix:
    """torch sparse CSR / CSC / COO tensor (or Matrix) -> Matrix.  COO is converted to CSR once (sorted, duplicates summed)."""
    if isinstance(native, Matrix):
        return native
    if not isinstance(native, torch.Tensor):
        raise TypeError(f"phb200 needs a torch sparse matrix, got {type(native)}")
    E.require_cuda(native)
    shape = tuple(native.shape)
    if native.layout == torch.sparse_csr:
        comps = ('csr', native.crow_indices(), native.col_indices(), native.values())
    elif native.layout == torch.sparse_csc:
        comps = ('csc', native.ccol_indices(), native.row_indices(), native.values())
    elif native.layout == torch.sparse_coo:
        comps = ('coo', native._indices(), native._indices(), native._values())
    else:
        raise TypeError("dense matrices are not on the phb200 path")
    key = _cache_key(*comps, shape)
    hit = _CACHE.get(key)
    if hit is not None:
        _CACHE.move_to_end(key)
        if detect_stencil and hit.kind == 'csr' and (grid is not None or not hit._stencil_checked):
            if hit.detect_stencil(grid) is None:          # cached without recognition (detect_stencil=False) earlier
                hit.detect_coded()
        return hit
    kind, a, b, v = comps
    if kind == 'coo':
        csr = native.coalesce().to_sparse_csr()      # one-off format conversion at set-up, not on the hot path
        m = Matrix.csr(csr.crow_indices(), csr.col_indices(), csr.values(), shape)
        keep = (native, csr)
    else:
        m = Matrix.csr(a, b, v, shape, transposed=(kind == 'csc'))
        keep = (native,)
    m._keep = m._keep + keep
    if detect_stencil and m.kind == 'csr':
        st = m.detect_stencil(grid)
        if st is None and kind == 'coo':
            st = _detect_ignoring_stored_zeros(m, grid)
        if st is None:
            m.detect_coded()
    _CACHE[key] = m
    while len(_CACHE) > _CACHE_SIZE:
        _CACHE.popitem(last=False)
    return m


def _detect_ignoring_stored_zeros(m: Matrix, grid=None):
    """COO natives whose values carry gradients (solve_linear(..., grad_for_f=True)) keep the boundary entries of a ZERO-padded
    stencil as stored zeros at the wrapped positions (the tracer cannot drop traced values).  They do not change a product, so
    the stencil twin — used for products only — is looked for on the matrix without them; the CSR itself keeps its pattern
    (ILU(0) is defined on it)."""
    nz = m.values != 0
    if bool(nz.all()):
        return None
    rows = torch.repeat_interleave(torch.arange(m.shape[0], device=m.device), (m.rowptr[1:] - m.rowptr[:-1]).to(torch.int64))
    counts = torch.zeros(m.shape[0] + 1, dtype=torch.int64, device=m.device)
    counts[1:] = torch.cumsum(torch.bincount(rows[nz], minlength=m.shape[0]), 0)
    pruned = Matrix.csr(counts.to(torch.int32), m.col[nz], m.values[nz], m.shape)
    st = pruned.detect_stencil(grid)
    if st is not None:
        m.stencil = st
        m._keep = m._keep + (pruned,)
    return st


@E.on_operand_device
def plain_csr(m: Matrix) -> Matrix:
    """CSC-native (arrays of A^T) -> CSR of A, via one transposition on the device (set-up cost, cached on the matrix)."""
    if m.kind != 'csc':
        return m
    t = getattr(m, '_plain', None)
    if t is None:
        rows_stored, nnz = m.shape[1], int(m.values.numel())          # the stored arrays are the CSR of A^T
        rowptr = torch.empty(m.shape[0] + 1, dtype=torch.int32, device=m.device)
        col = torch.empty(nnz, dtype=torch.int32, device=m.device)
        val = torch.empty(nnz, dtype=m.dtype, device=m.device)
        E.check(E.lib().phb200_csr_transpose(E.ptr(m.rowptr), E.ptr(m.col), E.ptr(m.values), rows_stored, m.shape[0], nnz, E.dtype_code(m.dtype),
                                             E.ptr(rowptr), E.ptr(col), E.ptr(val), E.stream_ptr(m.device)))
        t = Matrix.csr(rowptr, col, val, m.shape)
        if t.detect_stencil() is None:
            t.detect_coded()
        m._plain = t
    return t
