"""
Host side of the path: the operator interface of the reference's ``Backend`` for sparse products and linear solves,
implemented on the phb200 C ABI.  Same names, argument meaning and error behaviour as

* ``Backend.linear_solve``              phiml/backend/_backend.py:1460-1516
* ``TorchBackend.conjugate_gradient``   phiml/backend/torch/_torch_backend.py:932-946 (fast-path selection rules)
* ``TorchBackend.conjugate_gradient_adaptive`` :948-963, ``bi_conjugate_gradient`` :965-968
* ``Backend.mul_matrix_batched_vector`` phiml/backend/torch/_torch_backend.py:541-547
* ``Backend.mul_csr_dense`` / ``mul_coo_dense``  :912-922 / phiml/backend/_backend.py:1303-1328
* ``Backend.solve_triangular_sparse``   phiml/backend/_backend.py:1580-1586

There is no CPU path: tensors must live on a CUDA device and libphb200.so must be loadable.
"""
import ctypes
import warnings
from dataclasses import dataclass
from typing import Any, List, Optional

import numpy as np
import torch

from . import _engine as E
from .precond import ClusterPreconditioner, IncompleteLU, JacobiPreconditioner, Preconditioner
from .sparse import Matrix, as_matrix, plain_csr


@dataclass
class SolveResult:
    """Field-for-field mirror of phiml/backend/_backend.py:24-33."""
    method: str
    x: Any
    residual: Any
    iterations: Any
    function_evaluations: Any
    converged: Any
    diverged: Any
    message: List[str]


BACKEND_NAME = 'torch'     # the name the reference's method strings carry for this backend

# 'auto': a cost model picks the on-chip (resident) CG kernel for small systems; 'never' / 'always' (when it fits) for tests
RESIDENT = {'mode': 'auto'}
_RESIDENT_CODE = {'auto': -1, 'never': 0, 'always': 1}
OPT_RESIDENT, OPT_RESIDENT_USED = 1, 2
LAST_RUN = {'resident': False}     # diagnostics of the most recent solve (tests, bench)


def _pre_str(pre) -> str:
    return f"with preconditioner '{pre}'" if pre is not None else "with preconditioner 'none'"


def method_string(kind: str, pre, trajectory: bool, poly_order: int = 2, traced: bool = False) -> str:
    """``SolveResult.method`` exactly as the reference's torch backend words it (phiml/backend/torch/_torch_backend.py:946,963;
    phiml/backend/_linalg.py:90,128,224,274).  ``kind``: 'CG' | 'CG-adaptive' | 'biCG-stab'."""
    fast = "TorchScript" if traced else "PyTorch*"
    if kind == 'CG-adaptive' and pre:
        kind = 'CG'                                   # cg_adaptive falls back to cg when a preconditioner is given (_linalg.py:97-99)
    if kind == 'CG':
        return f"Φ-ML CG ({BACKEND_NAME}) {_pre_str(pre)}" if (trajectory or pre) else f"Φ-ML CG ({fast})"
    if kind == 'CG-adaptive':
        return f"Φ-ML CG-adaptive ({BACKEND_NAME}) " if trajectory else f"Φ-ML CG ({fast})"
    if kind == 'biCG-stab':
        return f"Φ-ML biCG-stab {_pre_str(pre)}" if poly_order == 1 else f"Φ-ML biCG-stab({poly_order}) ({BACKEND_NAME}) {_pre_str(pre)}"
    raise NotImplementedError(kind)


class DeviceSolver:
    """Thin owner of a ``phb200_solver`` (one batched solve in flight)."""

    def __init__(self, matrix: Matrix, pre, method: int, poly_order: int, batch: int):
        self._h = ctypes.c_void_p()
        self.matrix, self.pre = matrix, pre
        E.check(E.lib().phb200_solver_create(ctypes.byref(self._h), matrix.handle, pre.handle if pre is not None else None, method, poly_order, batch))
        self.workspace = None
        self._keep = ()
        self.option(OPT_RESIDENT, _RESIDENT_CODE[RESIDENT['mode']])

    def option(self, key: int, value: int = -2) -> int:
        """Sets (value >= -1) or reads (value = -2) an integer solver option; returns the previous / current value."""
        prev = ctypes.c_int()
        E.check(E.lib().phb200_solver_option(self._h, key, value, ctypes.byref(prev)))
        return int(prev.value)

    @property
    def resident_used(self) -> bool:
        return bool(self.option(OPT_RESIDENT_USED))

    def workspace_bytes(self) -> int:
        return int(E.lib().phb200_solver_workspace_bytes(self._h))

    def start(self, y, x0, rtol, atol, max_iter, zero_init=False, buffers=None):
        """``buffers``: (x, r, workspace) allocated by the caller (native distributed mode: the slab neighbours map them beforehand)."""
        dev = y.device
        alloc = torch.zeros_like if zero_init else torch.empty_like     # slab mode: halo planes must start as zeros
        nbytes = self.workspace_bytes()
        if buffers is not None:
            self.x, self.r, self.workspace = buffers
            assert self.x.shape == y.shape and self.r.shape == y.shape and self.workspace.numel() >= nbytes
        else:
            self.x = alloc(y)
            self.r = alloc(y)
            self.workspace = (torch.zeros if zero_init else torch.empty)(nbytes, dtype=torch.uint8, device=dev)
        self._keep = (y, x0, rtol, atol, max_iter)
        E.check(E.lib().phb200_solver_start(self._h, E.ptr(y), E.ptr(x0), E.ptr(self.x), E.ptr(self.r), E.ptr(rtol), E.ptr(atol), E.ptr(max_iter),
                                            E.ptr(self.workspace), nbytes, E.stream_ptr(dev)))

    def run(self, cap: int):
        E.check(E.lib().phb200_solver_run(self._h, int(cap), E.stream_ptr(self.x.device)))

    def advance(self, n: int):
        E.check(E.lib().phb200_solver_advance(self._h, int(n), E.stream_ptr(self.x.device)))

    def any_continue(self) -> bool:
        flag = ctypes.c_int()
        E.check(E.lib().phb200_solver_poll(self._h, ctypes.byref(flag), E.stream_ptr(self.x.device)))
        return bool(flag.value)

    def result(self):
        dev = self.x.device
        nb = self.x.shape[0]
        its = torch.empty(nb, dtype=torch.int32, device=dev)
        fev = torch.empty(nb, dtype=torch.int32, device=dev)
        conv = torch.empty(nb, dtype=torch.uint8, device=dev)
        div = torch.empty(nb, dtype=torch.uint8, device=dev)
        E.check(E.lib().phb200_solver_result(self._h, E.ptr(its), E.ptr(fev), E.ptr(conv), E.ptr(div), E.stream_ptr(dev)))
        return its, fev, conv.bool(), div.bool()

    def get_tol(self) -> torch.Tensor:
        """Device pair {minimum squared tolerance of this solver's systems, sum of |y_b|^2 over them} (after start())."""
        t = torch.empty(2, dtype=torch.float64, device=self.x.device)
        E.check(E.lib().phb200_solver_tol_min(self._h, E.ptr(t), 0, E.stream_ptr(self.x.device)))
        return t

    get_tol_min = get_tol

    def set_yy_sum(self, t: torch.Tensor):
        """Installs t[1] = sum of |y_b|^2 over ALL systems of a batch that is split over several solvers / ranks."""
        self._keep = self._keep + (t,)
        E.check(E.lib().phb200_solver_tol_min(self._h, E.ptr(t), 2, E.stream_ptr(self.x.device)))

    def set_tol_min(self, t: torch.Tensor):
        """Installs t[0] = minimum tolerance over ALL systems; repeats the first progress check."""
        self._keep = self._keep + (t,)
        E.check(E.lib().phb200_solver_tol_min(self._h, E.ptr(t), 1, E.stream_ptr(self.x.device)))

    def profile(self, enable: bool):
        E.check(E.lib().phb200_solver_profile(self._h, 1 if enable else 0))

    def profile_read(self, max_slots=32):
        ms = (ctypes.c_double * max_slots)()
        cnt = (ctypes.c_int64 * max_slots)()
        n = ctypes.c_int()
        E.check(E.lib().phb200_solver_profile_read(self._h, ms, cnt, max_slots, ctypes.byref(n)))
        return [(ms[k], int(cnt[k])) for k in range(n.value)]

    def stats(self):
        p, l = ctypes.c_int64(), ctypes.c_int64()
        E.check(E.lib().phb200_solver_stats(self._h, ctypes.byref(p), ctypes.byref(l)))
        return int(p.value), int(l.value)

    def __del__(self):
        try:
            if self._h is not None and E._lib is not None:
                E._lib.phb200_solver_destroy(self._h)
        except Exception:
            pass
        self._h = None


def _float_type(*tensors) -> torch.dtype:
    for t in tensors:
        if isinstance(t, torch.Tensor) and t.dtype == torch.float64:
            return torch.float64
    return torch.float32


def _prepare(lin, y, x0, rtol, atol, max_iter):
    if callable(lin) and not isinstance(lin, (torch.Tensor, Matrix)):
        raise NotImplementedError("matrix-free solves are not on the phb200 path (build a matrix with matrix_from_function)")
    if isinstance(lin, (tuple, list)):
        raise AssertionError("Batched matrices are not yet supported")      # same as _torch_backend.py:936
    matrix = plain_csr(as_matrix(lin))
    y = torch.as_tensor(y)
    E.require_cuda(y)
    dt = matrix.dtype
    y = y.to(dt).contiguous()
    x0 = torch.as_tensor(x0, device=y.device).to(dt).contiguous()
    assert y.dim() == 2 and x0.shape == y.shape and y.shape[1] == matrix.shape[0], f"y, x0 must be (batch, {matrix.shape[0]})"
    nb = y.shape[0]
    rtol = torch.as_tensor(rtol, device=y.device).to(dt).reshape(-1).expand(nb).contiguous()
    atol = torch.as_tensor(atol, device=y.device).to(dt).reshape(-1).expand(nb).contiguous()
    max_iter = np.asarray(max_iter)
    assert max_iter.ndim == 2, "max_iter must have shape (checkpoints, batch)"
    return matrix, y, x0, rtol, atol, max_iter


def _device_pre(pre, matrix):
    if pre is None:
        return None
    if isinstance(pre, (JacobiPreconditioner, IncompleteLU, ClusterPreconditioner)):
        return pre
    raise NotImplementedError(f"preconditioner {type(pre).__name__} cannot run inside the fused device loop; "
                              f"use phiml_b200.compute_preconditioner('jacobi'|'ilu'|'cluster', matrix)")


def _run(code: int, poly_order: int, name: str, matrix, y, x0, rtol, atol, max_iter: np.ndarray, pre) -> SolveResult:
    nb = y.shape[0]
    # products go through the stencil twin when the matrix was recognised as one (the ILU factors keep the CSR pattern)
    solver_matrix = matrix.product_twin
    if nb <= 65535:
        x, r, its, fev, conv, div = _run_slab(code, poly_order, solver_matrix, y, x0, rtol, atol, max_iter, pre)
        return SolveResult(name, x, r, its, fev, conv, div, [""] * nb)
    assert max_iter.shape[0] == 1, "trajectories are limited to 65535 systems"
    # gridDim.y limit: one solver per slab of systems.  The stopping rule couples the whole batch (minimum tolerance; for the
    # generic cg_adaptive / bicg loops the sum of |y|^2 over the batch), so the slabs are started first, the two quantities are
    # combined over them — as multi_gpu.solve_sharded does over ranks — and only then do the loops run.
    solvers = []
    for b0 in range(0, nb, 65535):
        sl = slice(b0, min(nb, b0 + 65535))
        mi = max_iter[:, sl] if max_iter.shape[1] > 1 else max_iter
        mi_last = np.broadcast_to(mi[-1, :], (sl.stop - sl.start,))
        mi_dev = torch.as_tensor(np.ascontiguousarray(mi_last).astype(np.int32), device=y.device)
        solver = DeviceSolver(solver_matrix, pre, code, poly_order, sl.stop - sl.start)
        solver.start(y[sl], x0[sl], rtol[sl], atol[sl], mi_dev)
        solvers.append(solver)
    combine_tolerances(solvers)
    parts = []
    for solver in solvers:
        solver.run(int(np.max(max_iter)))
        parts.append((solver.x, solver.r, *solver.result()))
    LAST_RUN['resident'] = solvers[-1].resident_used
    x, r, its, fev, conv, div = [torch.cat(p, dim=0) for p in zip(*parts)]
    return SolveResult(name, x, r, its, fev, conv, div, [""] * nb)


def combine_tolerances(solvers, all_reduce=None):
    """Started solvers that together hold ONE batch: make every one of them stop on the batch-wide rule of stop_on_l2
    (phiml/backend/_linalg.py:26-29).  ``all_reduce(tensor, op)`` with op 'sum' | 'min' additionally combines over ranks."""
    pairs = [s.get_tol() for s in solvers]
    total = torch.stack(pairs).sum(dim=0)
    if all_reduce is not None:
        all_reduce(total, 'sum')
    for s in solvers:
        s.set_yy_sum(total)
    pairs = [s.get_tol() for s in solvers]
    least = torch.stack(pairs).min(dim=0).values
    if torch.isnan(torch.stack(pairs)[:, 0]).any():
        least[0] = float('nan')
    if all_reduce is not None:
        all_reduce(least, 'min')
    for s in solvers:
        s.set_tol_min(least)


def _run_slab(code, poly_order, matrix, y, x0, rtol, atol, max_iter, pre):
    nb = y.shape[0]
    mi_last = np.broadcast_to(max_iter[-1, :], (nb,)) if max_iter.shape[1] in (1, nb) else max_iter[-1, :]
    mi_dev = torch.as_tensor(np.ascontiguousarray(mi_last).astype(np.int32), device=y.device)
    solver = DeviceSolver(matrix, pre, code, poly_order, nb)
    solver.start(y, x0, rtol, atol, mi_dev)
    if max_iter.shape[0] == 1:
        solver.run(int(np.max(max_iter)))
        LAST_RUN['resident'] = solver.resident_used
        its, fev, conv, div = solver.result()
        return solver.x, solver.r, its, fev, conv, div
    # --- trajectory recording (Backend.while_loop with a list of checkpoints, _backend.py:722-735) ---
    assert np.all(max_iter == max_iter[:, :1]), "When recording a trajectory, max_iter must be equal for all batch entries"
    checkpoints = max_iter[:, 0].tolist()
    def snap():
        flags = solver.result()       # first: result() also writes the NaNs of systems that stopped on a non-finite residual (include/phb200.h)
        return (solver.x.clone(), solver.r.clone(), *flags)
    trj = [snap()] if 0 in checkpoints else []
    for i in range(1, max(checkpoints) + 1):
        solver.advance(1)
        if i in checkpoints:
            trj.append(snap())
        if not solver.any_continue():
            break
    trj.extend([trj[-1]] * (len(checkpoints) - len(trj)))
    return tuple(torch.stack(parts) for parts in zip(*trj))


# ----------------------------------------------------------------------------------------------------------------------
# Backend.linear_solve and friends
# ----------------------------------------------------------------------------------------------------------------------

@E.on_operand_device
def conjugate_gradient(lin, y, x0, rtol, atol, max_iter, pre=None, matrix_offset=None) -> SolveResult:
    assert matrix_offset is None, "matrix_offset is not supported"
    matrix, y, x0, rtol, atol, max_iter = _prepare(lin, y, x0, rtol, atol, max_iter)
    if max_iter.shape[0] > 1 or pre:
        pre = _device_pre(pre, matrix)
        return _run(E.CG, 0, method_string('CG', pre, True), matrix, y, x0, rtol, atol, max_iter, pre)
    return _run(E.CG_TORCH, 0, method_string('CG', None, False), matrix, y, x0, rtol, atol, max_iter, None)


@E.on_operand_device
def conjugate_gradient_adaptive(lin, y, x0, rtol, atol, max_iter, pre=None, matrix_offset=None) -> SolveResult:
    assert matrix_offset is None, "matrix_offset is not supported"
    if pre:
        warnings.warn("CG-adaptive does not yet support preconditioners. Using regular CG instead.", RuntimeWarning)
        return conjugate_gradient(lin, y, x0, rtol, atol, max_iter, pre, matrix_offset)
    matrix, y, x0, rtol, atol, max_iter = _prepare(lin, y, x0, rtol, atol, max_iter)
    if max_iter.shape[0] > 1:
        return _run(E.CG_ADAPTIVE, 0, method_string('CG-adaptive', None, True), matrix, y, x0, rtol, atol, max_iter, None)
    return _run(E.CG_ADAPTIVE_TORCH, 0, method_string('CG-adaptive', None, False), matrix, y, x0, rtol, atol, max_iter, None)


@E.on_operand_device
def bi_conjugate_gradient(lin, y, x0, rtol, atol, max_iter, pre=None, matrix_offset=None, poly_order=2) -> SolveResult:
    assert matrix_offset is None, "matrix_offset is not supported"
    if poly_order == 0:
        raise NotImplementedError("Generic non-stabilized biCG not supported. Use 'scipy-biCG' instead")
    matrix, y, x0, rtol, atol, max_iter = _prepare(lin, y, x0, rtol, atol, max_iter)
    if pre and poly_order > 1:
        warnings.warn(f"Φ-ML biCG-stab({poly_order}) with preconditioner is experimental and may diverge.", RuntimeWarning)
    pre = _device_pre(pre, matrix)
    name = method_string('biCG-stab', pre, False, poly_order)
    return _run(E.BICGSTAB, poly_order, name, matrix, y, x0, rtol, atol, max_iter, pre)


@E.on_operand_device
def generic_conjugate_gradient(lin, y, x0, rtol, atol, max_iter, pre=None) -> SolveResult:
    """The generic loop ``_linalg.cg`` regardless of the torch fast-path rules (what the NumPy backend runs)."""
    matrix, y, x0, rtol, atol, max_iter = _prepare(lin, y, x0, rtol, atol, max_iter)
    pre = _device_pre(pre, matrix)
    return _run(E.CG, 0, f"Φ-ML CG ({BACKEND_NAME}) {_pre_str(pre)}", matrix, y, x0, rtol, atol, max_iter, pre)


@E.on_operand_device
def generic_conjugate_gradient_adaptive(lin, y, x0, rtol, atol, max_iter) -> SolveResult:
    matrix, y, x0, rtol, atol, max_iter = _prepare(lin, y, x0, rtol, atol, max_iter)
    return _run(E.CG_ADAPTIVE, 0, f"Φ-ML CG-adaptive ({BACKEND_NAME}) ", matrix, y, x0, rtol, atol, max_iter, None)


@E.on_operand_device
def linear_solve(method: str, lin, y, x0, rtol, atol, max_iter, pre=None, matrix_offset=None) -> SolveResult:
    """Method-string dispatch of Backend.linear_solve."""
    if method == 'auto':
        return conjugate_gradient_adaptive(lin, y, x0, rtol, atol, max_iter, pre, matrix_offset)
    elif method.startswith('scipy-'):
        raise NotImplementedError(f"'{method}' runs on the CPU through SciPy in the reference and is out of scope of the phb200 path")
    elif method == 'CG':
        return conjugate_gradient(lin, y, x0, rtol, atol, max_iter, pre, matrix_offset)
    elif method == 'CG-adaptive':
        return conjugate_gradient_adaptive(lin, y, x0, rtol, atol, max_iter, pre, matrix_offset)
    elif method in ['biCG', 'biCG-stab(0)']:
        return bi_conjugate_gradient(lin, y, x0, rtol, atol, max_iter, pre, matrix_offset, poly_order=0)
    elif method == 'biCG-stab':
        return bi_conjugate_gradient(lin, y, x0, rtol, atol, max_iter, pre, matrix_offset, poly_order=1)
    elif method.startswith('biCG-stab('):
        order = int(method[len('biCG-stab('):-1])
        return bi_conjugate_gradient(lin, y, x0, rtol, atol, max_iter, pre, matrix_offset, poly_order=order)
    else:
        raise NotImplementedError(f"Method '{method}' not supported for linear solve.")


# ----------------------------------------------------------------------------------------------------------------------
# products
# ----------------------------------------------------------------------------------------------------------------------

@E.on_operand_device
def mul_matrix_batched_vector(A, b: torch.Tensor) -> torch.Tensor:
    """(A b^T)^T for b of shape (batch, N)."""
    return as_matrix(A).matvec(b)


linear = mul_matrix_batched_vector


@E.on_operand_device
def mul_csr_dense(column_indices, row_pointers, values, shape, dense) -> torch.Tensor:
    """
    column_indices (batch, nnz), row_pointers (batch, rows+1), values (batch, nnz, channels),
    dense (batch, cols, channels, dense_cols)  ->  (batch, rows, channels, dense_cols)

    Mirrors `TorchBackend.mul_csr_dense` (phiml/backend/torch/_torch_backend.py:912-922).  The reference loops over (batch, channel) pairs in
    Python and builds a native per pair; here ONE matrix with one channel runs on the solver's CSR product kernels and everything else is
    one launch of `phb200_csr_dense_batched`.  Both sum a row's entries in entry order with separately rounded products (SciPy's order).
    """
    values, dense = torch.as_tensor(values), torch.as_tensor(dense)
    E.require_cuda(values, dense)
    dt = _float_type(values, dense)
    values, dense = values.to(dt), dense.to(dt)
    column_indices = torch.as_tensor(column_indices, device=values.device)
    row_pointers = torch.as_tensor(row_pointers, device=values.device)
    batch_size, nnz, channels = values.shape
    rows, cols = int(shape[0]), int(shape[1])
    assert dense.shape[0] == batch_size and dense.shape[1] == cols and dense.shape[2] == channels, f"dense {tuple(dense.shape)} does not match values {tuple(values.shape)} and shape {shape}"
    if batch_size == 1 and channels == 1:
        m = Matrix.csr(row_pointers[0], column_indices[0], values[0, :, 0], (rows, cols))
        return m.matvec(dense[0, :, 0, :].t().contiguous()).t()[None, :, None, :].contiguous()
    return _csr_dense_batched(column_indices, row_pointers, values, dense, rows, cols)


def _csr_dense_batched(column_indices: torch.Tensor, row_pointers: torch.Tensor, values: torch.Tensor, dense: torch.Tensor, rows: int, cols: int) -> torch.Tensor:
    """All (batch, channel, dense column) products of `mul_csr_dense` / `mul_coo_dense` in one launch (include/phb200.h: phb200_csr_dense_batched)."""
    batch_size, nnz, channels = values.shape
    dense_cols = dense.shape[-1]
    assert column_indices.shape[-1] == nnz and row_pointers.shape[-1] == rows + 1
    col = column_indices.to(torch.int32).expand(batch_size, nnz).contiguous()
    ptr = row_pointers.to(torch.int32).expand(batch_size, rows + 1).contiguous()
    values, dense = values.contiguous(), dense.contiguous()
    out = torch.empty((batch_size, rows, channels, dense_cols), dtype=values.dtype, device=values.device)
    E.check(E.lib().phb200_csr_dense_batched(E.ptr(col), E.ptr(ptr), E.ptr(values), E.ptr(dense), E.ptr(out), batch_size, rows, cols, nnz, channels,
                                              dense_cols, E.dtype_code(values.dtype), E.stream_ptr(values.device)))
    return out


@E.on_operand_device
def mul_coo_dense(indices, values, shape, dense) -> torch.Tensor:
    """
    indices (batch, nnz, 2), values (batch, nnz, channels), dense (batch, cols, channels, dense_cols)  ->  (batch, rows, channels, dense_cols)

    The reference scatters with atomics (`Backend.mul_coo_dense`, phiml/backend/_backend.py:1303-1328: gather, multiply, `scatter_add` — order
    unspecified).  Here the entries are ordered by row ONCE per index array (stable sort: entry order kept inside a row, cached) and the product
    runs on the CSR kernels — one launch for the whole batch: every row is summed sequentially in entry order, deterministic, duplicates included.
    """
    values, dense = torch.as_tensor(values), torch.as_tensor(dense)
    E.require_cuda(values, dense)
    dt = _float_type(values, dense)
    values, dense = values.to(dt), dense.to(dt)
    indices = torch.as_tensor(indices, device=values.device)
    batch_size, nnz, channels = values.shape
    rows, cols = int(shape[0]), int(shape[1])
    assert dense.shape[0] == batch_size and dense.shape[1] == cols and dense.shape[2] == channels
    rowptr, perm, col = _coo_row_order(indices, rows)
    if perm.shape[0] != batch_size:
        rowptr, perm, col = rowptr.expand(batch_size, -1), perm.expand(batch_size, -1), col.expand(batch_size, -1)
    if batch_size == 1 and channels == 1:
        m = Matrix.csr(rowptr[0], col[0], values[0, :, 0][perm[0]].contiguous(), (rows, cols))
        return m.matvec(dense[0, :, 0, :].t().contiguous()).t()[None, :, None, :].contiguous()
    ordered = torch.gather(values, 1, perm[:, :, None].expand(batch_size, nnz, channels))
    return _csr_dense_batched(col, rowptr, ordered, dense, rows, cols)


_COO_ORDER: 'dict' = {}


def _coo_row_order(indices: torch.Tensor, rows: int):
    """
    (rowptr int32 (batch, rows+1), permutation int64 (batch, nnz), ordered columns int32 (batch, nnz)) of a (batch, nnz, 2) coordinate array:
    entries ordered by row, entry order kept inside a row.  Cached on the caller's array (storage + version; the entry dies with the array).
    """
    key = (indices.data_ptr(), indices._version, tuple(indices.shape), str(indices.dtype), rows, str(indices.device))
    hit = _COO_ORDER.get(key)
    if hit is not None and hit[0]() is indices:
        return hit[1]
    import weakref
    row = indices[:, :, 0].to(torch.int64)
    row_sorted, perm = torch.sort(row, dim=1, stable=True)
    bounds = torch.arange(rows + 1, device=indices.device, dtype=torch.int64)[None, :].expand(indices.shape[0], rows + 1).contiguous()
    rowptr = torch.searchsorted(row_sorted.contiguous(), bounds).to(torch.int32)      # entries with a row index below r
    col = torch.gather(indices[:, :, 1], 1, perm).to(torch.int32).contiguous()
    value = (rowptr, perm, col)
    if len(_COO_ORDER) >= 8:
        _COO_ORDER.pop(next(iter(_COO_ORDER)))
    try:
        _COO_ORDER[key] = (weakref.ref(indices), value)
    except TypeError:
        pass
    return value


# ----------------------------------------------------------------------------------------------------------------------
# triangular solves
# ----------------------------------------------------------------------------------------------------------------------

class _Plan:
    def __init__(self, h):
        self._h = h

    def __del__(self):
        try:
            if self._h is not None and E._lib is not None:
                E._lib.phb200_sptrsv_plan_destroy(self._h)
        except Exception:
            pass
        self._h = None


@E.on_operand_device
def solve_triangular_sparse(matrix, rhs: torch.Tensor, lower: bool, unit_diagonal: bool) -> torch.Tensor:
    """X = T^-1 B for a sparse triangular native T and B of shape (batch, N); the level schedule is cached on the matrix."""
    m = plain_csr(as_matrix(matrix, detect_stencil=False))
    rhs = torch.as_tensor(rhs)
    E.require_cuda(rhs)
    rhs = rhs.to(m.dtype).contiguous()
    assert rhs.dim() == 2 and rhs.shape[1] == m.shape[0]
    plans = m.__dict__.setdefault('_plans', {})
    if lower not in plans:
        h = ctypes.c_void_p()
        E.check(E.lib().phb200_sptrsv_plan(ctypes.byref(h), m.handle, 1 if lower else 0, E.stream_ptr(rhs.device)))
        plans[lower] = _Plan(h)
    x = torch.empty_like(rhs)
    E.check(E.lib().phb200_sptrsv(m.handle, plans[lower]._h, 1 if lower else 0, 1 if unit_diagonal else 0, E.ptr(rhs), E.ptr(x),
                                  rhs.shape[0], rhs.shape[1], E.stream_ptr(rhs.device)))
    return x


@E.on_operand_device
def solve_triangular(matrix, rhs, lower: bool, unit_diagonal: bool):
    if isinstance(matrix, Matrix) or (isinstance(matrix, torch.Tensor) and matrix.layout != torch.strided):
        return solve_triangular_sparse(matrix, rhs, lower, unit_diagonal)
    raise NotImplementedError("dense triangular solves are not on the phb200 path")
