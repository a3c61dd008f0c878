"""
The solve as ONE registered torch operator, ``phb200::linear_solve`` ("under ``math.jit_compile`` the call is traced by
``torch.jit.trace`` (phiml/backend/torch/_torch_backend.py:1193), so the extension entry must be a registered torch op").

The reference makes its own loops traceable by scripting them (``@torch.jit._script_if_tracing`` on ``torch_sparse_cg``,
_torch_backend.py:1353) — and warns that every other method "will always run the maximum number of iterations when JIT-compiled"
(:950, :967) because a traced Python ``while`` is unrolled.  Here the whole solve is a single graph node: ``torch.jit.trace`` records
``phb200::linear_solve(matrix arrays, y, x0, rtol, atol, ...)`` with the method string, the iteration limits and the preconditioner
*recipe* as constants, and every execution of the traced graph runs the device-resident loop with its on-device stopping test — same
iteration counts as the eager call.

Only tensors and plain constants cross the operator boundary, so the preconditioner travels as a recipe ('' | 'jacobi' |
'ilu:<sweeps>:<safe>' | 'ilu-exact') and is built inside the operator from the matrix arrays it receives (a matrix whose values are
graph inputs gets fresh factors on every execution; constant matrices hit the identity caches of ``sparse.as_matrix`` and of this module).
"""
from collections import OrderedDict
from typing import Sequence, Tuple

import numpy as np
import torch
from torch import Tensor

OP_NAME = "phb200::linear_solve"
_STATE = {'registered': False, 'calls': 0}
_PRE_CACHE: 'OrderedDict[tuple, object]' = OrderedDict()


def native_from_parts(parts: Sequence[Tensor], layout: str, rows: int, cols: int) -> Tensor:
    """Re-assembles the torch sparse native from the arrays that crossed the operator boundary (no copies)."""
    if layout == 'csr':
        return torch.sparse_csr_tensor(parts[0], parts[1], parts[2], (rows, cols))
    if layout == 'csc':
        return torch.sparse_csc_tensor(parts[0], parts[1], parts[2], (rows, cols))
    if layout == 'coo':
        return torch.sparse_coo_tensor(parts[0], parts[1], (rows, cols))
    raise ValueError(f"unknown sparse layout '{layout}'")


def parts_of_native(native: Tensor):
    """(arrays, layout) of a torch sparse tensor — the inverse of ``native_from_parts``; traceable (plain accessor ops)."""
    if native.layout == torch.sparse_csr:
        return [native.crow_indices(), native.col_indices(), native.values()], 'csr'
    if native.layout == torch.sparse_csc:
        return [native.ccol_indices(), native.row_indices(), native.values()], 'csc'
    if native.layout == torch.sparse_coo:
        return [native._indices(), native._values()], 'coo'
    raise TypeError("dense matrices are not on the phb200 path")


def pre_recipe(pre) -> str:
    """Preconditioner object -> the constant that can cross the operator boundary."""
    from .precond import IncompleteLU, JacobiPreconditioner
    if pre is None:
        return ''
    if isinstance(pre, JacobiPreconditioner):
        return 'jacobi'
    if isinstance(pre, IncompleteLU):
        return 'ilu-exact' if pre.sweeps is None else f"ilu:{pre.sweeps}:{int(getattr(pre, 'safe', False))}"
    raise NotImplementedError(f"preconditioner {type(pre).__name__} cannot cross the phb200::linear_solve operator boundary")


def build_pre(recipe: str, native: Tensor):
    from . import precond
    if not recipe:
        return None
    values = native.values() if native.layout != torch.sparse_coo else native._values()
    key = (recipe, values.data_ptr(), values._version, tuple(native.shape), values.dtype)
    hit = _PRE_CACHE.get(key)
    if hit is not None and hit[0] is values:
        _PRE_CACHE.move_to_end(key)
        return hit[1]
    if recipe == 'jacobi':
        pre = precond.JacobiPreconditioner(native)
    elif recipe == 'ilu-exact':
        pre = precond.IncompleteLU(native, sweeps='exact')
    elif recipe.startswith('ilu:'):
        _, sweeps, safe = recipe.split(':')
        pre = precond.IncompleteLU(native, sweeps=int(sweeps), safe=bool(int(safe)))
    else:
        raise NotImplementedError(f"preconditioner recipe '{recipe}'")
    _PRE_CACHE[key] = (values, pre)
    while len(_PRE_CACHE) > 4:
        _PRE_CACHE.popitem(last=False)
    return pre


def _solve_impl(matrix, layout, rows, cols, y, x0, rtol, atol, max_iter, checkpoints, method, pre):
    from . import solve as host
    _STATE['calls'] += 1
    native = native_from_parts(matrix, layout, rows, cols)
    mi = np.asarray(list(max_iter), dtype=np.int64).reshape(checkpoints, -1)
    r = host.linear_solve(method, native, y, x0, rtol, atol, mi, build_pre(pre, native), None)
    return r.x, r.residual, r.iterations, r.function_evaluations, r.converged, r.diverged


def register():
    """Idempotent registration of ``phb200::linear_solve`` with the torch dispatcher."""
    if _STATE['registered']:
        return torch.ops.phb200.linear_solve

    @torch.library.custom_op(OP_NAME, mutates_args=())
    def linear_solve(matrix: Sequence[Tensor], layout: str, rows: int, cols: int, y: Tensor, x0: Tensor, rtol: Tensor, atol: Tensor,
                     max_iter: Sequence[int], checkpoints: int, method: str, pre: str) -> Tuple[Tensor, Tensor, Tensor, Tensor, Tensor, Tensor]:
        return _solve_impl(matrix, layout, rows, cols, y, x0, rtol, atol, max_iter, checkpoints, method, pre)

    @linear_solve.register_fake
    def _(matrix, layout, rows, cols, y, x0, rtol, atol, max_iter, checkpoints, method, pre):
        lead = (checkpoints,) if checkpoints > 1 else ()
        nb = y.shape[0]
        vec = y.new_empty(lead + tuple(y.shape))
        i32 = lambda: y.new_empty(lead + (nb,), dtype=torch.int32)
        flag = lambda: y.new_empty(lead + (nb,), dtype=torch.bool)
        return vec, y.new_empty(lead + tuple(y.shape)), i32(), i32(), flag(), flag()

    _STATE['registered'] = True
    return torch.ops.phb200.linear_solve


def traced_linear_solve(method: str, lin: Tensor, y, x0, rtol, atol, max_iter, pre):
    """Calls the registered operator (what ``torch.jit.trace`` records as one node).  Returns the six result tensors."""
    op = register()
    parts, layout = parts_of_native(lin)
    values = parts[-1]
    dt = values.dtype
    y = torch.as_tensor(y).to(dt)
    x0 = torch.as_tensor(x0, device=y.device).to(dt)
    nb = y.shape[0]
    rtol = torch.as_tensor(rtol, device=y.device).to(dt).reshape(-1).expand(nb)
    atol = torch.as_tensor(atol, device=y.device).to(dt).reshape(-1).expand(nb)
    mi = np.asarray(max_iter)
    assert mi.ndim == 2, "max_iter must have shape (checkpoints, batch)"
    return op(parts, layout, int(lin.shape[0]), int(lin.shape[1]), y, x0, rtol, atol, [int(v) for v in mi.reshape(-1)], int(mi.shape[0]), method, pre_recipe(pre))
