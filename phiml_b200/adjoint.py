"""
Implicit-gradient (adjoint) step of ``math.solve_linear`` on the phb200 path.

Mirrors ``implicit_gradient_solve`` (phiml/math/_optimize.py:785-817): the backward pass of a linear solve ``A x = y`` with incoming
gradient ``dx = dL/dx`` is

* one more solve on the transposed matrix with the forward method and tolerances (``Solve.gradient_solve``):  ``A^T dy = dx``,
  ``dy = dL/dy``.  The reference transposes by renaming dims (``transpose_matrix`` :823-833), so ``Backend.linear_solve`` receives
  the SAME arrays under the other compressed layout — CSR forward → CSC backward, COO → COO with swapped coordinates;
* with ``grad_for_f=True`` the matrix gradient on the stored entries, ``dm[k] = -sum_b dy[b,row[k]] * x[b,col[k]]`` (:803-807),
  here one kernel (``phb200_matrix_gradient_coo`` / ``_csr``) instead of two gathers, a product, a batch sum and a negation.

Nothing here is differentiable by itself: ``implicit_gradient_solve`` IS the derivative rule.  ``SolveLinear`` wraps it in a
``torch.autograd.Function`` for native-tensor users (PhiML installs its own rule through ``custom_gradient``, :819).
"""
from typing import Optional, Tuple

import numpy as np
import torch

from . import _engine as E
from . import solve as _solve
from .sparse import Matrix, as_matrix


@E.on_operand_device
def transpose(native):
    """A^T the way the reference's backward pass presents it to ``Backend.linear_solve``: zero-copy re-labelling of the stored
    arrays (CSR <-> CSC), swapped coordinates for COO, mirrored offsets for a constant-coefficient stencil handle."""
    if isinstance(native, Matrix):
        if native.kind == 'stencil':
            spec = getattr(native, 'spec', None)
            if spec is None:
                raise NotImplementedError("only stencil handles built by Matrix.stencil_matrix can be transposed; pass the native instead")
            offsets, coeffs = spec
            return Matrix.stencil_matrix(native.grid, [tuple(-int(o) for o in off) for off in offsets], coeffs, native.dtype, native.device)
        if native.kind == 'csr':
            return Matrix.csr(native.rowptr, native.col, native.values, (native.shape[1], native.shape[0]), transposed=True)
        return Matrix.csr(native.rowptr, native.col, native.values, (native.shape[1], native.shape[0]), transposed=False)
    E.require_cuda(native)
    rows, cols = native.shape
    if native.layout == torch.sparse_csr:
        return torch.sparse_csc_tensor(native.crow_indices(), native.col_indices(), native.values(), (cols, rows))
    if native.layout == torch.sparse_csc:
        return torch.sparse_csr_tensor(native.ccol_indices(), native.row_indices(), native.values(), (cols, rows))
    if native.layout == torch.sparse_coo:
        idx = native._indices()
        return torch.sparse_coo_tensor(torch.stack([idx[1], idx[0]]), native._values(), (cols, rows))
    raise TypeError("dense matrices are not on the phb200 path")


def _entries(native) -> Tuple[str, torch.Tensor, torch.Tensor, Tuple[int, int]]:
    if isinstance(native, Matrix):
        if native.kind != 'csr':
            raise NotImplementedError("matrix gradients are defined on the stored entries of a CSR / COO matrix")
        return 'csr', native.rowptr, native.col, native.shape
    if native.layout == torch.sparse_csr:
        return 'csr', native.crow_indices().to(torch.int32), native.col_indices().to(torch.int32), tuple(native.shape)
    if native.layout == torch.sparse_coo:
        idx = native._indices()
        return 'coo', idx[0].contiguous(), idx[1].contiguous(), tuple(native.shape)
    raise NotImplementedError("matrix gradients need a CSR or COO native (the reference decompresses CSR to COO, _optimize.py:800-801)")


@E.on_operand_device
def matrix_gradient(native, dy: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
    """``dm[k] = -sum_b dy[b,row[k]] * x[b,col[k]]`` for every stored entry k of ``native`` (entry order of the native)."""
    kind, a, b, (rows, cols) = _entries(native)
    E.require_cuda(dy, x)
    dt = dy.dtype
    assert dt in (torch.float32, torch.float64) and x.dtype == dt, (dy.dtype, x.dtype)
    assert dy.dim() == 2 and x.dim() == 2 and dy.shape[0] == x.shape[0] and dy.shape[1] == rows and x.shape[1] == cols, (tuple(dy.shape), tuple(x.shape), rows, cols)
    dy, x = dy.contiguous(), x.contiguous()
    nnz = int(b.numel())
    out = torch.empty(nnz, dtype=dt, device=dy.device)
    if kind == 'coo':
        E.check(E.lib().phb200_matrix_gradient_coo(E.ptr(a), E.ptr(b), nnz, rows, cols, E.ptr(dy), E.ptr(x), dy.shape[0], dy.shape[1], x.shape[1],
                                                   E.ptr(out), E.dtype_code(dt), E.stream_ptr(dy.device)))
    else:
        a, b = a.contiguous(), b.contiguous()
        E.check(E.lib().phb200_matrix_gradient_csr(E.ptr(a), E.ptr(b), rows, cols, nnz, E.ptr(dy), E.ptr(x), dy.shape[0], dy.shape[1], x.shape[1],
                                                   E.ptr(out), E.dtype_code(dt), E.stream_ptr(dy.device)))
    return out


@E.on_operand_device
def implicit_gradient_solve(method: str, native, x: torch.Tensor, dx: torch.Tensor, rtol, atol, max_iter, pre=None, matrix_adjoint: bool = False,
                            x0: Optional[torch.Tensor] = None):
    """Backward pass of ``x = solve(A, y)``.  Returns ``(SolveResult of A^T dy = dx, dm or None)``; ``.x`` of the result is dL/dy.
    ``pre`` must be a preconditioner of A^T (the reference rebuilds it from the transposed matrix, _optimize.py:644)."""
    x0 = torch.zeros_like(dx) if x0 is None else x0          # grad_solve.x0 defaults to zeros, _optimize.py:795
    res = _solve.linear_solve(method, transpose(native), dx, x0, rtol, atol, max_iter, pre, None)
    dm = matrix_gradient(native, res.x, x) if matrix_adjoint else None
    return res, dm


class SolveLinear(torch.autograd.Function):
    """``x = A^-1 y`` for native torch users, differentiable w.r.t. ``y`` and (COO / CSR natives) the stored matrix values.

        x = SolveLinear.apply(values, y, native_without_grad, method, rtol, atol, max_iter)

    ``values`` is the value array of ``native`` (pass it separately so autograd sees it); both solves run on phb200."""

    @staticmethod
    def forward(ctx, values, y, native, method, rtol, atol, max_iter):
        nb = y.shape[0]
        rt, at = np.full(nb, rtol, dtype=np.float64), np.full(nb, atol, dtype=np.float64)
        mi = np.full((1, nb), int(max_iter))
        res = _solve.linear_solve(method, native, y, torch.zeros_like(y), rt, at, mi, None, None)
        ctx.native, ctx.args = native, (method, rt, at, mi)
        ctx.save_for_backward(res.x)
        ctx.info = res
        return res.x

    @staticmethod
    def backward(ctx, dx):
        x, = ctx.saved_tensors
        method, rt, at, mi = ctx.args
        res, dm = implicit_gradient_solve(method, ctx.native, x, dx.contiguous(), rt, at, mi, None, matrix_adjoint=ctx.needs_input_grad[0])
        return dm, res.x, None, None, None, None, None
