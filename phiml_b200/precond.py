"""
Preconditioners of the phb200 path, mirroring the reference protocol ``Preconditioner``
(phiml/backend/_backend.py:36-47: apply / apply_transposed / apply_inv_l / apply_inv_u on (batch, N) tensors).

* ``JacobiPreconditioner`` — NEW; ``compute_preconditioner`` of the reference knows only 'auto' | 'ilu' | 'cluster'
  (phiml/math/_optimize.py:836-880).
* ``IncompleteLU`` — device counterpart of phiml/backend/_linalg.py:455-487.  The factors are the exact ILU(0) of the
  stored pattern, i.e. the fixed point the reference's sweep iteration (``incomplete_lu_coo``, _linalg.py:524-594)
  converges to; the triangular solves are level-scheduled on the GPU (the reference's torch backend has none:
  ``solve_triangular_sparse`` raises, _backend.py:1580-1581).
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

    def apply(self, vec):
        return vec * self.inv_diag

    apply_transposed = apply
    apply_inv_l = apply

    def apply_inv_u(self, vec):
        return vec

    def __repr__(self):
        return "jacobi"


class IncompleteLU(Preconditioner, _Handle):
    """Exact ILU(0) on the matrix' own pattern with level-scheduled solves.  ``lu`` holds L (strict lower, unit diagonal
    implied) and U (upper incl. diagonal) in A's CSR layout."""

    def __init__(self, matrix, source: str = "exact"):
        _Handle.__init__(self)
        m = plain_csr(as_matrix(matrix))
        assert m.kind == 'csr' and m.shape[0] == m.shape[1], "incomplete_lu_coo only implemented for square matrices"
        self.matrix = m
        self.source = source
        self.lu = torch.empty_like(m.values)
        h = ctypes.c_void_p()
        twin = m.stencil.handle if m.stencil is not None else None
        E.check(E.lib().phb200_precond_ilu0(ctypes.byref(h), m.handle, E.ptr(self.lu), twin, E.stream_ptr(m.device)))
        self._h = h

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
        m = self.matrix
        n = m.shape[0]
        rows = torch.repeat_interleave(torch.arange(n, device=m.device, dtype=torch.int64), (m.rowptr[1:] - m.rowptr[:-1]).to(torch.int64))
        cols = m.col.to(torch.int64)
        lower, diag = rows > cols, rows == cols
        l_val = torch.where(lower, self.lu, torch.ones_like(self.lu))
        keep_l = lower | diag
        L = torch.sparse_coo_tensor(torch.stack([rows[keep_l], cols[keep_l]]), l_val[keep_l], m.shape).coalesce().to_sparse_csr()
        U = torch.sparse_coo_tensor(torch.stack([rows[~lower], cols[~lower]]), self.lu[~lower], m.shape).coalesce().to_sparse_csr()
        return L, U

    def apply(self, vec):
        vec = vec.to(self.matrix.dtype).contiguous()
        out = torch.empty_like(vec)
        E.check(E.lib().phb200_precond_apply(self._h, E.ptr(vec), E.ptr(out), vec.shape[0], vec.shape[1], E.stream_ptr(vec.device)))
        return out

    def apply_inv_l(self, vec):
        from .solve import solve_triangular_sparse
        return solve_triangular_sparse(self.factors()[0], vec, lower=True, unit_diagonal=True)

    def apply_inv_u(self, vec):
        from .solve import solve_triangular_sparse
        return solve_triangular_sparse(self.factors()[1], vec, lower=False, unit_diagonal=False)

    def __repr__(self):
        return f"ilu ({self.source})"


def compute_preconditioner(method: Optional[str], matrix, **_ignored) -> Optional[Preconditioner]:
    """'jacobi' | 'ilu' | None on a native matrix (torch sparse tensor or Matrix)."""
    if method is None:
        return None
    if method == 'jacobi':
        return JacobiPreconditioner(matrix)
    if method == 'ilu':
        return IncompleteLU(matrix)
    raise NotImplementedError(f"preconditioner '{method}' is not on the phb200 path")
