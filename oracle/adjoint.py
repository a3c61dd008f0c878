"""
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

NumPy/SciPy restatement of the implicit-gradient (adjoint) step of ``math.solve_linear``:

* ``transpose``                -> phiml/math/_optimize.py:823-833 (transpose_matrix: the backward solve runs on A^T)
* ``implicit_gradient_solve``  -> phiml/math/_optimize.py:785-817: ``dy = solve(A^T, dx)`` with the forward method and tolerances
                                  (``Solve.gradient_solve``), and with ``grad_for_f`` the matrix gradient sampled on the stored
                                  entries, ``dm[k] = -sum_b dy[b, row[k]] * x[b, col[k]]`` (:803-807)

``row`` / ``col`` are given entry by entry (COO order as the reference stores the matrix while its values carry gradients;
for a CSR matrix ``decompress()`` (:800-801) expands the row pointers in place, i.e. the CSR order).
"""
import numpy as np
import scipy.sparse as sp

from . import krylov


def transpose(csr: sp.csr_matrix) -> sp.csr_matrix:
    t = csr.T.tocsr()
    t.sort_indices()
    return t


def matrix_gradient(row: np.ndarray, col: np.ndarray, dy: np.ndarray, x: np.ndarray) -> np.ndarray:
    """dm[k] = -sum_b dy[b,row[k]] * x[b,col[k]]  (_optimize.py:805-806), summed in batch order."""
    out = np.zeros(len(row), dtype=dy.dtype)
    for b in range(dy.shape[0]):
        out += dy[b, row] * x[b, col]
    return -out


def csr_rows(csr: sp.csr_matrix) -> np.ndarray:
    """row index of every stored entry of a CSR matrix (CompressedSparseMatrix.decompress, phiml/math/_sparse.py)."""
    return np.repeat(np.arange(csr.shape[0], dtype=np.int64), np.diff(csr.indptr))


def implicit_gradient_solve(method: str, csr: sp.csr_matrix, x: np.ndarray, dx: np.ndarray, rtol, atol, max_iter, pre=None, flavour: str = "numpy",
                            entries=None):
    """Returns (OracleResult of the adjoint solve A^T dy = dx, dm or None).  ``entries`` = (row, col) of the entries dm is wanted for."""
    res = krylov.linear_solve(method, transpose(csr), dx, np.zeros_like(dx), rtol, atol, max_iter, pre, flavour)
    dm = None
    if entries is not None:
        dm = matrix_gradient(np.asarray(entries[0]), np.asarray(entries[1]), res.x, x)
    return res, dm
