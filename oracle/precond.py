"""
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

CPU restatement of the preconditioners on the path:

* ``Jacobi``              -> NEW plug-in (the reference has no 'jacobi', phiml/math/_optimize.py:836-880); oracle =
                             the reference's generic ``cg`` driven with this object.
* ``ilu_fixed_point``     -> phiml/backend/_linalg.py:524-594 (incomplete_lu_coo): Jacobi-style fixed-point sweeps
                             on the stored pattern; ``sum_l_u`` is the strict-lower x strict-upper product sampled
                             on the pattern (the reference enumerates the same products with index tables,
                             _linalg.py:597-634).
* ``ilu_sweeps``          -> phiml/math/_optimize.py:855-869 (auto sweep count, eager branch).
* ``ilu0_exact``          -> textbook IKJ ILU(0): the fixed point of the sweeps.
* ``IncompleteLU``        -> phiml/backend/_linalg.py:455-487 with SciPy ``spsolve_triangular`` as in
                             phiml/backend/_numpy_backend.py:579-580.
* ``sptrsv``              -> plain forward/backward substitution on CSR (checker for the GPU level-scheduled solve).
"""
import math

import numpy as np
import scipy.sparse as sp
from scipy.sparse.linalg import spsolve_triangular


class Jacobi:
    """M^-1 = diag(A)^-1.  apply_inv_l / apply_inv_u split it as (D^-1, I) for BiCGstab(1)."""

    def __init__(self, csr: sp.csr_matrix):
        self.inv_diag = (1.0 / csr.diagonal()).astype(csr.dtype)

    def apply(self, v): return v * self.inv_diag
    def apply_transposed(self, v): return v * self.inv_diag
    def apply_inv_l(self, v): return v * self.inv_diag
    def apply_inv_u(self, v): return v
    def __repr__(self): return "jacobi"


def ilu_sweeps(n: int, nnz: int) -> int:
    d = (nnz / n - 1) / 2
    if d < 1:
        return 1
    return int(math.ceil(math.sqrt(d * n ** (1 / d))))


def ilu_fixed_point(csr: sp.csr_matrix, sweeps: int, safe: bool = False):
    """
    Returns (L, U) as CSR: L strict lower + explicit unit diagonal, U upper including the diagonal.
    State ``lu`` lives on the stored pattern of ``csr`` (which must contain the full diagonal).
    """
    a = csr.tocoo()
    row, col, val = a.row, a.col, a.data
    n = csr.shape[0]
    lower = row > col
    is_diag = row == col
    diag_pos = np.full(n, -1, np.int64)
    diag_pos[row[is_diag]] = np.nonzero(is_diag)[0]
    assert np.all(diag_pos >= 0), "All diagonal elements must be present in sparse matrix."
    piv = diag_pos[np.minimum(row, col)]              # entry index of (min(i,j), min(i,j))
    lu = np.where(is_diag, val, 0) + np.where(lower, val / val[piv], 0)
    lu = lu.astype(val.dtype)
    for _ in range(sweeps):
        ls = sp.csr_matrix((np.where(lower, lu, 0), (row, col)), shape=(n, n))
        us = sp.csr_matrix((np.where(~lower & ~is_diag, lu, 0), (row, col)), shape=(n, n))
        prod = (ls @ us).tocsr()
        s = np.asarray(prod[row, col]).ravel().astype(val.dtype)
        rest = val - s
        pivv = lu[piv]
        if safe:
            with np.errstate(divide='ignore', invalid='ignore'):
                low = np.where(pivv == 0, 0, rest / pivv)
        else:
            low = rest / pivv
        lu = np.where(lower, low, rest).astype(val.dtype)
    l_val = np.where(lower, lu, is_diag.astype(val.dtype))
    keep_l = lower | is_diag
    L = sp.csr_matrix((l_val[keep_l], (row[keep_l], col[keep_l])), shape=(n, n))
    U = sp.csr_matrix((lu[~lower], (row[~lower], col[~lower])), shape=(n, n))
    return L, U


def ilu0_exact(csr: sp.csr_matrix):
    """Sequential IKJ ILU(0) on sorted CSR; returns (L with unit diagonal stored, U incl. diagonal) as CSR."""
    a = csr.copy().tocsr()
    a.sort_indices()
    ptr, idx, val = a.indptr, a.indices, a.data.copy()
    n = a.shape[0]
    diag = np.empty(n, np.int64)
    for i in range(n):
        lo, hi = ptr[i], ptr[i + 1]
        pos = {int(idx[p]): p for p in range(lo, hi)}
        for p in range(lo, hi):
            k = int(idx[p])
            if k >= i:
                break
            val[p] = val[p] / val[diag[k]]
            for q in range(diag[k] + 1, ptr[k + 1]):
                t = pos.get(int(idx[q]))
                if t is not None:
                    val[t] = val[t] - val[p] * val[q]
        diag[i] = pos[i]
    coo_row = np.repeat(np.arange(n), np.diff(ptr))
    lower = coo_row > idx
    is_diag = coo_row == idx
    l_val = np.where(lower, val, 1).astype(val.dtype)
    keep_l = lower | is_diag
    L = sp.csr_matrix((l_val[keep_l], (coo_row[keep_l], idx[keep_l])), shape=(n, n))
    U = sp.csr_matrix((val[~lower], (coo_row[~lower], idx[~lower])), shape=(n, n))
    return L, U


def sptrsv(tri: sp.csr_matrix, rhs: np.ndarray, lower: bool, unit_diagonal: bool) -> np.ndarray:
    """Substitution in row order; rhs (batch, N).  The stored diagonal is ignored when unit_diagonal."""
    t = tri.tocsr()
    t.sort_indices()
    ptr, idx, val = t.indptr, t.indices, t.data
    n = t.shape[0]
    x = np.array(rhs, copy=True)
    order = range(n) if lower else range(n - 1, -1, -1)
    for i in order:
        acc = x[:, i].copy()
        d = 1.0
        for p in range(ptr[i], ptr[i + 1]):
            j = idx[p]
            if j == i:
                d = val[p]
            else:
                acc = acc - val[p] * x[:, j]
        x[:, i] = acc if unit_diagonal else acc / d
    return x


class IncompleteLU:
    """apply = U^-1 L^-1 v through SciPy, like the reference's NumPy backend."""

    def __init__(self, L, U, source="exact"):
        self.L, self.U, self.source = L.tocsr(), U.tocsr(), source

    def apply_inv_l(self, v):
        return spsolve_triangular(self.L, v.T, lower=True, unit_diagonal=True).T.astype(v.dtype)

    def apply_inv_u(self, v):
        return spsolve_triangular(self.U, v.T, lower=False, unit_diagonal=False).T.astype(v.dtype)

    def apply(self, v):
        return self.apply_inv_u(self.apply_inv_l(v))

    def __repr__(self):
        return f"ilu ({self.source})"


def cluster_assignment(grid, cluster_count: int = 3 ** 6, dtype=np.float64):
    """Cell -> cluster map of ``explicit_coarse`` for grid matrices (phiml/math/_optimize.py:947-960): one cluster per cell when the
    grid has no more cells than clusters, else a box decomposition with ``round(size * (count / volume)^(1/rank))`` clusters per axis,
    cluster coordinate = int(index / size * clusters_along_axis) evaluated in the working precision.  Returns (clusters int32 (N,), count)."""
    grid = tuple(int(g) for g in grid)
    n = int(np.prod(grid))
    if cluster_count >= n:
        return np.arange(n, dtype=np.int32), n
    by_axis = np.round(np.asarray(grid) * (cluster_count / n) ** (1 / len(grid))).astype(np.int32)
    ft = np.dtype(dtype).type
    coords = [(np.arange(g).astype(ft) / ft(g) * ft(c)).astype(np.int32) for g, c in zip(grid, by_axis)]
    mesh = np.meshgrid(*coords, indexing='ij')
    clusters = np.ravel_multi_index([m.reshape(-1) for m in mesh], tuple(int(c) for c in by_axis)).astype(np.int32)
    return clusters, int(np.prod(by_axis))


class Cluster:
    """ExplicitClusterSolve (phiml/backend/_linalg.py:734-766) with the parts of coarse_explicit_preconditioner_coo (:769-799):
    coarse matrix = entries summed by (cluster of row, cluster of column), inverted densely with NumPy."""

    def __init__(self, csr: sp.csr_matrix, clusters: np.ndarray, cluster_count: int):
        self.clusters, self.cluster_count = clusters, cluster_count
        coo = csr.tocoo()
        coarse = np.zeros((cluster_count, cluster_count), csr.dtype)
        np.add.at(coarse, (clusters[coo.row], clusters[coo.col]), coo.data)
        self.inv_coarse = np.linalg.inv(coarse)
        self.cluster_size = np.bincount(clusters, minlength=cluster_count).astype(csr.dtype)

    def apply(self, v):
        coarse = np.stack([np.bincount(self.clusters, weights=row, minlength=self.cluster_count) for row in v])
        delta = v - (coarse / self.cluster_size)[:, self.clusters]
        solution = (self.inv_coarse @ coarse.T).T
        return (solution[:, self.clusters] + delta).astype(v.dtype)

    def apply_inv_l(self, v): return v
    def apply_inv_u(self, v): return self.apply(v)
    def __repr__(self): return f"cluster ({self.cluster_count})"
