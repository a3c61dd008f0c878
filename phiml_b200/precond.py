"""
Preconditioners of the phb200 path, mirroring the reference protocol ``Preconditioner``
(phiml/backend/_backend.py:36-47: apply / apply_transposed / apply_inv_l / apply_inv_u on (batch, N) tensors).

* ``JacobiPreconditioner`` — NEW; ``compute_preconditioner`` of the reference knows only 'auto' | 'ilu' | 'cluster'
  (phiml/math/_optimize.py:836-880).
* ``IncompleteLU`` — device counterpart of phiml/backend/_linalg.py:455-487.  By default the factors are the reference's own:
  the fixed-point sweeps of ``incomplete_lu_coo`` (_linalg.py:524-594) with the sweep count of ``compute_preconditioner``
  (phiml/math/_optimize.py:855-869), one gather-multiply-add kernel per sweep; ``sweeps='exact'`` gives the fixed point itself
  (exact ILU(0)).  The triangular solves run on the GPU (the reference's torch backend has none: ``solve_triangular_sparse``
  raises, _backend.py:1580-1581): structured wavefronts on recognised star stencils, level-scheduled CSR solves otherwise.
"""
import ctypes
from typing import Optional

import torch

from . import _engine as E
from .sparse import Matrix, as_matrix, plain_csr


class Preconditioner:
    """Same protocol as the reference's Preconditioner."""

    def apply(self, vec):
        raise NotImplementedError

    def apply_transposed(self, vec):
        raise NotImplementedError

    def apply_inv_l(self, vec):
        raise NotImplementedError

    def apply_inv_u(self, vec):
        raise NotImplementedError


class _Handle:
    def __init__(self):
        self._h = None

    @property
    def handle(self):
        return self._h

    def __del__(self):
        try:
            if self._h is not None and E._lib is not None:
                E._lib.phb200_precond_destroy(self._h)
        except Exception:
            pass
        self._h = None


class JacobiPreconditioner(Preconditioner, _Handle):
    """M^-1 = diag(A)^-1; the split used by biCG-stab(1) is (L^-1, U^-1) = (D^-1, I)."""

    @E.on_operand_device
    def __init__(self, matrix):
        _Handle.__init__(self)
        m = as_matrix(matrix)
        assert m.shape[0] == m.shape[1], "Jacobi needs a square matrix"
        self.matrix = m
        src = m.stencil if m.stencil is not None else m
        self.inv_diag = torch.empty(m.shape[0], dtype=m.dtype, device=m.device)
        E.check(E.lib().phb200_jacobi_setup(src.handle, E.ptr(self.inv_diag), E.stream_ptr(m.device)))
        h = ctypes.c_void_p()
        # a stencil matrix has a constant diagonal: the fused kernels then take 1/diag as a scalar
        E.check(E.lib().phb200_precond_jacobi(ctypes.byref(h), E.ptr(self.inv_diag), m.shape[0], E.dtype_code(m.dtype), 1 if src.kind == 'stencil' else 0))
        self._h = h

    @E.on_operand_device
    def apply(self, vec):
        return vec * self.inv_diag

    apply_transposed = apply
    apply_inv_l = apply

    @E.on_operand_device
    def apply_inv_u(self, vec):
        return vec

    def __repr__(self):
        return "jacobi"


def ilu_sweeps(n: int, nnz: int) -> int:
    """Number of fixed-point sweeps ``compute_preconditioner`` picks in eager mode (phiml/math/_optimize.py:855-869)."""
    import math
    d = (nnz / n - 1) / 2
    if d < 1:
        return 1
    return int(math.ceil(math.sqrt(d * n ** (1 / d))))


class IncompleteLU(Preconditioner, _Handle):
    """ILU on the matrix' own pattern with the triangular solves on the GPU.  ``lu`` holds L (strict lower, unit diagonal
    implied) and U (upper incl. diagonal) in A's CSR layout.

    ``sweeps``: ``None`` — the reference's own factors: ``ilu_sweeps(n, nnz)`` fixed-point sweeps of ``incomplete_lu_coo``
    (phiml/backend/_linalg.py:524-594), reported as ``'ilu (iter=k)'`` like the reference's; an int — that many sweeps;
    ``'exact'`` — the fixed point itself (sequential-equivalent ILU(0), level-scheduled), reported as ``'ilu (exact)'``.
    ``safe``: ``divide_no_nan`` for the lower entries (rank-deficient matrices, _linalg.py:576)."""

    @E.on_operand_device
    def __init__(self, matrix, sweeps=None, safe: bool = False, rank_deficiency=0):
        _Handle.__init__(self)
        m = plain_csr(as_matrix(matrix))
        assert m.kind == 'csr' and m.shape[0] == m.shape[1], "incomplete_lu_coo only implemented for square matrices"
        self.matrix = m
        self.rank_deficiency = rank_deficiency
        self.safe = bool(safe)
        self.lower_unit_diagonal, self.upper_unit_diagonal = True, False
        self.lu = torch.empty_like(m.values)
        h = ctypes.c_void_p()
        twin = m.stencil.handle if m.stencil is not None else None
        if sweeps == 'exact':
            self.sweeps = None
            self.source = "exact"
            E.check(E.lib().phb200_precond_ilu0(ctypes.byref(h), m.handle, E.ptr(self.lu), twin, E.stream_ptr(m.device)))
        else:
            self.sweeps = int(sweeps) if sweeps is not None else ilu_sweeps(m.shape[0], int(m.values.numel()))
            self.source = f"iter={self.sweeps}"
            E.check(E.lib().phb200_precond_ilu_sweeps(ctypes.byref(h), m.handle, E.ptr(self.lu), self.sweeps, 1 if safe else 0, twin, E.stream_ptr(m.device)))
        self._h = h
        self._factors = None
        self._factors_t = None

    @property
    def star_wavefront(self) -> bool:
        """True if the triangular solves run as the structured wavefront (csrc/star_trsv.cuh) instead of CSR level solves."""
        kind, star = ctypes.c_int(), ctypes.c_int()
        E.check(E.lib().phb200_precond_info(self._h, ctypes.byref(kind), ctypes.byref(star)))
        return bool(star.value)

    @property
    def levels(self):
        lo, up = ctypes.c_int64(), ctypes.c_int64()
        E.check(E.lib().phb200_precond_levels(self._h, ctypes.byref(lo), ctypes.byref(up)))
        return int(lo.value), int(up.value)

    def factors(self):
        """(L, U) as torch sparse CSR tensors: L with explicit unit diagonal, U including the diagonal — the layout
        ``incomplete_lu_coo`` returns."""
        if self._factors is not None:
            return self._factors
        m = self.matrix
        n = m.shape[0]
        rows = torch.repeat_interleave(torch.arange(n, device=m.device, dtype=torch.int64), (m.rowptr[1:] - m.rowptr[:-1]).to(torch.int64))
        cols = m.col.to(torch.int64)
        lower, diag = rows > cols, rows == cols
        l_val = torch.where(lower, self.lu, torch.ones_like(self.lu))
        keep_l = lower | diag
        L = torch.sparse_coo_tensor(torch.stack([rows[keep_l], cols[keep_l]]), l_val[keep_l], m.shape).coalesce().to_sparse_csr()
        U = torch.sparse_coo_tensor(torch.stack([rows[~lower], cols[~lower]]), self.lu[~lower], m.shape).coalesce().to_sparse_csr()
        self._factors = (L, U)
        return self._factors

    @property
    def lower(self):
        return self.factors()[0]

    @property
    def upper(self):
        return self.factors()[1]

    @E.on_operand_device
    def apply(self, vec):
        vec = vec.to(self.matrix.dtype).contiguous()
        out = torch.empty_like(vec)
        E.check(E.lib().phb200_precond_apply(self._h, E.ptr(vec), E.ptr(out), vec.shape[0], vec.shape[1], E.stream_ptr(vec.device)))
        return out

    @E.on_operand_device
    def apply_transposed(self, vec):
        """(LU)^-T vec = L^-T (U^-T vec), phiml/backend/_linalg.py:481-485: a lower solve with U^T followed by an upper solve with
        L^T; the transposed factors are built once on the device (phb200_csr_transpose)."""
        from .solve import solve_triangular_sparse
        if self._factors_t is None:
            L, U = self.factors()
            self._factors_t = (plain_csr(as_matrix(torch.sparse_csc_tensor(L.crow_indices(), L.col_indices(), L.values(), tuple(L.shape)), detect_stencil=False)),
                               plain_csr(as_matrix(torch.sparse_csc_tensor(U.crow_indices(), U.col_indices(), U.values(), tuple(U.shape)), detect_stencil=False)))
        Lt, Ut = self._factors_t
        intermediate = solve_triangular_sparse(Ut, vec, lower=True, unit_diagonal=self.upper_unit_diagonal)
        return solve_triangular_sparse(Lt, intermediate, lower=False, unit_diagonal=self.lower_unit_diagonal)

    @E.on_operand_device
    def apply_inv_l(self, vec):
        from .solve import solve_triangular_sparse
        return solve_triangular_sparse(self.factors()[0], vec, lower=True, unit_diagonal=True)

    @E.on_operand_device
    def apply_inv_u(self, vec):
        from .solve import solve_triangular_sparse
        return solve_triangular_sparse(self.factors()[1], vec, lower=False, unit_diagonal=False)

    def __repr__(self):
        return f"ilu ({self.source})"


def cluster_assignment(grid, cluster_count: int, dtype: torch.dtype, device) -> tuple:
    """Cell -> cluster map of ``explicit_coarse`` for grid matrices (phiml/math/_optimize.py:947-960) built on the device: one cluster
    per cell if the grid has no more cells than clusters, else boxes with ``round(size * (count / volume)^(1/rank))`` clusters per axis;
    the cluster coordinate is ``int(index / size * clusters_along_axis)`` evaluated in the working precision, like the reference."""
    import numpy as np
    grid = tuple(int(g) for g in grid)
    n = int(np.prod(grid))
    if cluster_count >= n:
        return torch.arange(n, dtype=torch.int32, device=device), n
    by_axis = np.round(np.asarray(grid) * (cluster_count / n) ** (1 / len(grid))).astype(np.int32)
    clusters = torch.zeros(grid, dtype=torch.int64, device=device)
    for axis, (g, c) in enumerate(zip(grid, by_axis)):
        coord = (torch.arange(g, device=device).to(dtype) / torch.tensor(g, dtype=dtype, device=device) * torch.tensor(int(c), dtype=dtype, device=device)).to(torch.int32).to(torch.int64)
        shape = [1] * len(grid)
        shape[axis] = g
        clusters = clusters * int(c) + coord.reshape(shape)
    return clusters.reshape(-1).to(torch.int32), int(np.prod(by_axis))


class ClusterPreconditioner(Preconditioner, _Handle):
    """Device counterpart of ``ExplicitClusterSolve`` (phiml/backend/_linalg.py:734-766), the reference's 'cluster' preconditioner:
    ``M^-1 v = P C^-1 R v + (v - P (R v / size))`` with R = sums over the cells of a cluster, C = coarse matrix (entries of A summed by
    cluster of row / column), P = gather by cluster.  It is applied inside the device-resident Krylov loop by three kernels
    (phb200_precond_cluster).  The parts are those the reference's object holds: ``inv_coarse_matrix (C, C)``, ``clusters (N,)``,
    ``cluster_count``, ``cluster_size (C,)``; ``from_matrix`` builds them like ``explicit_coarse`` / ``coarse_explicit_preconditioner_coo``
    (:769-799; the dense inverse is taken with NumPy on the host, as the reference does on every backend)."""

    @E.on_operand_device
    def __init__(self, inv_coarse_matrix: torch.Tensor, clusters: torch.Tensor, cluster_count: int, cluster_size: torch.Tensor):
        _Handle.__init__(self)
        E.require_cuda(inv_coarse_matrix, clusters, cluster_size)
        dt = inv_coarse_matrix.dtype
        assert dt in (torch.float32, torch.float64)
        self.cluster_count = int(cluster_count)
        self.inv_coarse_matrix = inv_coarse_matrix.contiguous()
        self.clusters = clusters.reshape(-1).to(torch.int32).contiguous()
        self.cluster_size = cluster_size.to(dt).contiguous()
        assert tuple(self.inv_coarse_matrix.shape) == (self.cluster_count, self.cluster_count) and self.cluster_size.numel() == self.cluster_count
        n = int(self.clusters.numel())
        # cells grouped by cluster, ascending inside a cluster (the order bincount adds them in)
        self.members = torch.argsort(self.clusters.to(torch.int64), stable=True).to(torch.int32).contiguous()
        counts = torch.bincount(self.clusters.to(torch.int64), minlength=self.cluster_count)
        self.member_ptr = torch.zeros(self.cluster_count + 1, dtype=torch.int32, device=self.clusters.device)
        self.member_ptr[1:] = torch.cumsum(counts, 0).to(torch.int32)
        h = ctypes.c_void_p()
        E.check(E.lib().phb200_precond_cluster(ctypes.byref(h), E.ptr(self.clusters), E.ptr(self.members), E.ptr(self.member_ptr), n, self.cluster_count,
                                               E.ptr(self.inv_coarse_matrix), E.ptr(self.cluster_size), E.dtype_code(dt)))
        self._h = h

    @classmethod
    @E.on_operand_device
    def from_matrix(cls, matrix, grid=None, cluster_count: int = 3 ** 6):
        import numpy as np
        m = plain_csr(as_matrix(matrix))
        assert m.kind == 'csr' and m.shape[0] == m.shape[1]
        if grid is None:
            grid = m.stencil.grid if m.stencil is not None else (m.shape[0],)
        assert int(np.prod(grid)) == m.shape[0], f"grid {grid} does not match a matrix with {m.shape[0]} rows"
        clusters, count = cluster_assignment(grid, cluster_count, m.dtype, m.device)
        rows = torch.repeat_interleave(torch.arange(m.shape[0], device=m.device, dtype=torch.int64), (m.rowptr[1:] - m.rowptr[:-1]).to(torch.int64))
        cl = clusters.to(torch.int64)
        lin = cl[rows] * count + cl[m.col.to(torch.int64)]
        coarse = torch.zeros(count * count, dtype=torch.float64, device=m.device).index_add_(0, lin, m.values.to(torch.float64))
        coarse = coarse.reshape(count, count).to(m.dtype)
        inv = torch.as_tensor(np.linalg.inv(coarse.cpu().numpy()), device=m.device)        # "inverting matrix ... using NumPy", _linalg.py:793-795
        size = torch.bincount(cl, minlength=count).to(m.dtype)
        return cls(inv, clusters, count, size)

    @E.on_operand_device
    def apply(self, vec):
        vec = vec.to(self.inv_coarse_matrix.dtype).contiguous()
        out = torch.empty_like(vec)
        E.check(E.lib().phb200_precond_apply(self._h, E.ptr(vec), E.ptr(out), vec.shape[0], vec.shape[1], E.stream_ptr(vec.device)))
        return out

    @E.on_operand_device
    def apply_transposed(self, vec):
        raise NotImplementedError      # as the reference (_linalg.py:755-756)

    @E.on_operand_device
    def apply_inv_l(self, vec):
        return vec

    @E.on_operand_device
    def apply_inv_u(self, vec):
        return self.apply(vec)

    def __repr__(self):
        return f"cluster ({self.cluster_count})"


@E.on_operand_device
def incomplete_lu_coo_device(indices, values, shape, iterations: int, safe: bool):
    """``incomplete_lu_coo`` (phiml/backend/_linalg.py:524-594) for ONE matrix with ONE channel whose values are a CUDA tensor, on the device sweeps:
    returns ``((l_indices, l_values), (u_indices, u_values))`` in the reference's layout — NumPy index arrays (1, nnz_L|U, 2) holding the lower + diagonal /
    the upper + diagonal coordinates in the order they were given, value tensors (1, nnz_L|U, 1) with L's unit diagonal stored — or ``None`` when the
    call is not of that kind (several matrices or channels, CPU values, repeated coordinates), which the caller hands to the reference."""
    import numpy as np
    if not (isinstance(values, torch.Tensor) and values.is_cuda and values.dim() == 3 and values.shape[0] == 1 and values.shape[2] == 1 and values.dtype in (torch.float32, torch.float64)):
        return None
    if int(shape[0]) != int(shape[1]) or tuple(indices.shape) != (1, values.shape[1], 2):
        return None
    with torch.cuda.device(values.device):
        n, dev = int(shape[0]), values.device
        idx = torch.as_tensor(indices, device=dev)[0].to(torch.int64)
        row, col = idx[:, 0].contiguous(), idx[:, 1].contiguous()
        key = row * n + col
        order = None
        if key.numel() > 1 and not bool((key[1:] > key[:-1]).all()):
            key, order = torch.sort(key, stable=True)
            if bool((key[1:] == key[:-1]).any()):
                return None                                   # repeated coordinates: not a CSR pattern
            row, col = row[order], col[order]
        vals = values[0, :, 0].detach()
        vals = (vals[order] if order is not None else vals).contiguous()
        rowptr = torch.searchsorted(row, torch.arange(n + 1, device=dev, dtype=torch.int64)).to(torch.int32)
        ilu = IncompleteLU(Matrix.csr(rowptr, col.to(torch.int32), vals, (n, n)), sweeps=int(iterations), safe=bool(safe))
        lu = ilu.lu
        if order is not None:
            lu = torch.empty_like(lu).index_copy_(0, order, lu)
        row, col = idx[:, 0], idx[:, 1]
        lower, diag = row > col, row == col
        keep_l = lower | diag
        l_val = torch.where(lower, lu, torch.ones_like(lu))[keep_l]
        u_val = lu[~lower]
        idx_host = np.asarray(indices.detach().cpu().numpy() if isinstance(indices, torch.Tensor) else indices)
        keep_l_h, upper_h = keep_l.cpu().numpy(), (~lower).cpu().numpy()
        return (idx_host[:, keep_l_h], l_val[None, :, None]), (idx_host[:, upper_h], u_val[None, :, None])


def compute_preconditioner(method: Optional[str], matrix, rank_deficiency=0, **_ignored) -> Optional[Preconditioner]:
    """'jacobi' | 'ilu' | 'ilu-exact' | 'cluster' | 'auto' | None on a native matrix (torch sparse tensor or Matrix).  'auto' resolves to 'ilu'
    as in the reference for backends with sparse triangular solves (phiml/math/_optimize.py:838-846); 'ilu' builds the
    reference's own factors (fixed-point sweeps, eager sweep count); ``safe`` sweeps when a rank deficiency is declared (:869)."""
    if method is None:
        return None
    if method == 'jacobi':
        return JacobiPreconditioner(matrix)
    if method in ('ilu', 'auto'):
        import numpy as np
        return IncompleteLU(matrix, safe=bool(np.any(np.asarray(rank_deficiency) > 0)), rank_deficiency=rank_deficiency)
    if method == 'ilu-exact':
        return IncompleteLU(matrix, sweeps='exact')
    if method == 'cluster':
        return ClusterPreconditioner.from_matrix(matrix, grid=_ignored.get('grid'))
    raise NotImplementedError(f"preconditioner '{method}' is not on the phb200 path")
