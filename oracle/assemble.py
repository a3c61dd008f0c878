"""
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

CPU restatement of how the reference turns a traced stencil into sparse natives:

* ``stencil_coo``   -> phiml/math/_lin_trace.py:750-807 (tracer_to_coo): one block of entries per shift in the
                       tracer's insertion order, inside a block C-order over the output cells whose coefficient
                       is non-zero (``np.nonzero`` drops the cut-off boundary entries), source = out + shift.
                       Index columns are (source dims..., output dims...) exactly like the reference's
                       ``(~x,~y[,~z],x,y[,z])`` channel.
* ``coo_to_csr``    -> phiml/math/_sparse.py:347-390 (SparseCoordinateTensor.compress): ravel row/col C-order,
                       let SciPy build the CSR of the permutation ``1..nnz`` → sorted int32 ``indices``/``indptr``.
* ``laplace_shifts`` / ``advdiff_shifts`` give the shift tables that ``-math.laplace(x, padding=ZERO)`` and the
  central-difference advection-diffusion operator of SURVEY.md §8(d) produce (insertion order measured on the
  reference; tests/test_oracle_golden.py pins both tables against the reference's COO entry order).
"""
from typing import Dict, Sequence, Tuple

import numpy as np
import scipy.sparse as sp


def laplace_shifts(rank: int, scale: float = -1.0) -> Tuple[Tuple[Tuple[int, ...], float], ...]:
    """
    Shift table of ``scale * laplace`` (scale=-1: diag 2*rank, off-diag -1), in the reference's insertion order
    ``x-1, x+1, self, y-1, y+1[, z-1, z+1]``.
    """
    out = []
    for d in range(rank):
        lo = tuple(-1 if i == d else 0 for i in range(rank))
        hi = tuple(+1 if i == d else 0 for i in range(rank))
        out.append((lo, scale * 1.0))
        out.append((hi, scale * 1.0))
        if d == 0:
            out.append((tuple([0] * rank), scale * (-2.0 * rank)))
    return tuple(out)


def advdiff_shifts(rank: int, velocity: Sequence[float], c: float, nu: float, dtype=np.float64):
    """
    ``x + c*(u . grad x) - nu*laplace(x)`` with central differences (h=1): 2*rank+1 constant coefficients.
    Insertion order measured on the reference trace: ``self, x+1, x-1, y+1, y-1[, z+1, z-1]``; the coefficients are
    evaluated in ``dtype`` with the tracer's operation order so that fp32 values are bit-identical.
    """
    t = np.dtype(dtype).type
    out = [(tuple([0] * rank), t(1) - t(nu) * t(-2.0 * rank))]
    for d in range(rank):
        lo = tuple(-1 if i == d else 0 for i in range(rank))
        hi = tuple(+1 if i == d else 0 for i in range(rank))
        adv = t(c) * (t(0.5) * t(velocity[d]))
        out.append((hi, adv - t(nu)))
        out.append((lo, -adv - t(nu)))
    return tuple(out)


def stencil_coo(grid: Sequence[int], shifts, dtype=np.float64, periodic: bool = False):
    """
    COO entries of a constant-coefficient stencil with ZERO (entries dropped) or PERIODIC (wrapped) boundaries.

    Returns ``(indices int64 (nnz, 2*rank) [source dims, output dims], values (nnz,))``.
    """
    grid = tuple(int(g) for g in grid)
    rank = len(grid)
    cells = np.stack(np.meshgrid(*[np.arange(g) for g in grid], indexing='ij'), -1).reshape(-1, rank)  # C-order
    idx_blocks, val_blocks = [], []
    for shift, coeff in shifts:
        if coeff == 0:
            continue
        src = cells + np.asarray(shift)
        if periodic:
            keep = np.ones(len(cells), bool)
            src = src % np.asarray(grid)
        else:
            keep = np.all((src >= 0) & (src < np.asarray(grid)), axis=1)
        idx_blocks.append(np.concatenate([src[keep], cells[keep]], axis=1))
        val_blocks.append(np.full(int(keep.sum()), coeff, dtype))
    return np.concatenate(idx_blocks).astype(np.int64), np.concatenate(val_blocks)


def shift_values_coo(grid: Sequence[int], offsets, values):
    """
    General form of phiml/math/_lin_trace.py:750-807 (tracer_to_coo) for NumPy-backed shift values: ``values[k]`` is the full-grid value array of
    shift ``offsets[k]``; per shift the cells with a non-zero value (``np.nonzero(np.sum(abs(v), 0))``, :773-774) are kept in C order and the source
    index wraps periodically (``(component + shift) % size``, :775).  Covers variable coefficients and PERIODIC boundaries.
    """
    grid = tuple(int(g) for g in grid)
    rank = len(grid)
    idx_blocks, val_blocks = [], []
    for shift, v in zip(offsets, values):
        v = np.asarray(v).reshape(grid)
        out_idx = np.nonzero(np.abs(v))
        src_idx = [(component + int(shift[d])) % grid[d] for d, component in enumerate(out_idx)]
        idx_blocks.append(np.stack(list(src_idx) + list(out_idx), axis=1))
        val_blocks.append(v[out_idx])
    return np.concatenate(idx_blocks).astype(np.int64).reshape(-1, 2 * rank), np.concatenate(val_blocks)


def shift_values_csr(grid, offsets, values) -> sp.csr_matrix:
    idx, val = shift_values_coo(grid, offsets, values)
    return coo_to_csr(grid, idx, val)


def coo_to_csr(grid: Sequence[int], indices: np.ndarray, values: np.ndarray) -> sp.csr_matrix:
    """Canonical CSR (sorted columns, int32) through the same SciPy permutation trick the reference uses."""
    grid = tuple(int(g) for g in grid)
    rank = len(grid)
    n = int(np.prod(grid))
    col = np.ravel_multi_index(tuple(indices[:, :rank].T), grid)
    row = np.ravel_multi_index(tuple(indices[:, rank:].T), grid)
    perm_csr = sp.csr_matrix((np.arange(1, len(values) + 1), (row, col)), shape=(n, n))
    perm = perm_csr.data - 1
    return sp.csr_matrix((values[perm], perm_csr.indices.astype(np.int32), perm_csr.indptr.astype(np.int32)), shape=(n, n))


def stencil_csr(grid, shifts, dtype=np.float64, periodic=False) -> sp.csr_matrix:
    idx, val = stencil_coo(grid, shifts, dtype, periodic)
    return coo_to_csr(grid, idx, val)


def poisson_csr(grid, dtype=np.float64) -> sp.csr_matrix:
    """``-laplace`` with ZERO padding on ``grid`` — the matrix of BASELINE configs 1-4."""
    return stencil_csr(grid, laplace_shifts(len(grid)), dtype)
