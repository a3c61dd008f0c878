"""
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

NumPy/SciPy restatement of the reference Krylov loops, written directly on ``(batch, N)`` arrays.
Each function names the reference lines it follows (paths relative to /root/reference):

* ``StopOnL2``            -> phiml/backend/_linalg.py:23-40   (stop_on_l2, incl. the (batch,batch) broadcast quirk)
* ``cg``                  -> phiml/backend/_linalg.py:52-90   (Shewchuk PCG, delta = r.M^-1 r)
* ``cg_adaptive``         -> phiml/backend/_linalg.py:93-128  (Hestenes-Stiefel variant)
* ``bicgstab_l``          -> phiml/backend/_linalg.py:131-224 (BiCGstab(L), L>=2; tau row aliasing reproduced)
* ``bicgstab_1``          -> phiml/backend/_linalg.py:227-274 (classic BiCGstab with shadow residual = ones)
* ``torch_fastpath_cg``   -> phiml/backend/torch/_torch_backend.py:932-946,1353-1382
* ``torch_fastpath_cg_adaptive`` -> phiml/backend/torch/_torch_backend.py:948-963,1385-1413
* ``run_loop``            -> phiml/backend/_backend.py:697-735 (eager while_loop, both branches)
* ``matvec``              -> phiml/backend/_numpy_backend.py:282-283 (per-batch ``A.dot(b[i])``)
* ``linear_solve``        -> phiml/backend/_backend.py:1460-1516 (method-string dispatch)
* ``mul_csr_dense``       -> phiml/backend/_numpy_backend.py:510-519 (one SciPy product per (batch, channel) pair = csr_matvecs order)
* ``mul_coo_dense``       -> phiml/backend/_backend.py:1303-1328 (gather, multiply, scatter-add; NumPy scatter = np.add.at, entry order)

The arithmetic dtype is the dtype of ``y`` (float32 or float64); tolerances are cast to it like the
reference does through ``backend.as_tensor`` of a float tensor.
"""
from dataclasses import dataclass
from typing import Callable, List, Optional, Union

import numpy as np


# ----------------------------------------------------------------------------------------------------------------------
# records and small helpers
# ----------------------------------------------------------------------------------------------------------------------

@dataclass
class OracleResult:
    """Mirror of SolveResult (phiml/backend/_backend.py:24-33), non-trajectory layout."""
    method: str
    x: np.ndarray                     # (batch, N)
    residual: np.ndarray              # (batch, N)
    iterations: np.ndarray            # (batch,) int32
    function_evaluations: np.ndarray  # (batch,) int32
    converged: np.ndarray             # (batch,) bool
    diverged: np.ndarray              # (batch,) bool
    message: List[str]


class IdentityPre:
    """NoPreconditioner, phiml/backend/_linalg.py:437-452."""
    def apply(self, v): return v
    def apply_inv_l(self, v): return v
    def apply_inv_u(self, v): return v
    def __repr__(self): return "none"


def safe_div(num, den):
    """x/y with 0 where y==0 — phiml/backend/_numpy_backend.py:181-184."""
    with np.errstate(divide='ignore', invalid='ignore'):
        q = num / den
    return np.where(den == 0, 0, q).astype(q.dtype, copy=False)


def matvec(lin, v: np.ndarray) -> np.ndarray:
    """A applied to every row of v.  Matrix: one SciPy product per batch entry (_numpy_backend.py:282-283)."""
    if callable(lin):
        return lin(v)
    return np.stack([lin.dot(v[i]) for i in range(v.shape[0])])


def mul_csr_dense(column_indices, row_pointers, values, shape, dense) -> np.ndarray:
    """(batch, nnz), (batch, rows+1), (batch, nnz, channels), (batch, cols, channels, dense_cols) -> (batch, rows, channels, dense_cols).
    Written out entry by entry: every output starts at 0 and receives `value * dense` of its row's entries in entry order (products and sums
    rounded separately in the arithmetic dtype) — what SciPy's csr_matvecs does for `mat * dense[b, :, c, :]` (_numpy_backend.py:510-519)."""
    batch, nnz, channels = values.shape
    rows, cols = shape
    out = np.zeros((batch, rows, channels, dense.shape[-1]), values.dtype)
    for b in range(batch):
        for r in range(rows):
            for k in range(row_pointers[b, r], row_pointers[b, r + 1]):
                out[b, r] += values[b, k][:, None] * dense[b, column_indices[b, k]]
    return out


def mul_coo_dense(indices, values, shape, dense) -> np.ndarray:
    """(batch, nnz, 2), (batch, nnz, channels), (batch, cols, channels, dense_cols) -> (batch, rows, channels, dense_cols); entries added in entry
    order, duplicates included (Backend.mul_coo_dense _backend.py:1303-1328 with NumPy's scatter, np.add.at, _numpy_backend.py:420)."""
    batch, nnz, channels = values.shape
    out = np.zeros((batch, shape[0], channels, dense.shape[-1]), values.dtype)
    for b in range(batch):
        for k in range(nnz):
            out[b, indices[b, k, 0]] += values[b, k][:, None] * dense[b, indices[b, k, 1]]
    return out


def rowdot(a, b):
    """Batched dot product kept as (batch, 1), like ``b.sum(a*b, -1, keepdims=True)``."""
    return np.sum(a * b, axis=-1, keepdims=True)


class StopOnL2:
    """
    Progress test of the reference (``stop_on_l2``).  ``tol_sq`` has shape (batch,) while the squared residual
    has shape (batch, 1); the comparison therefore broadcasts to (batch, batch) and ``all(axis=1)`` is taken,
    so every system has to reach the *smallest* tolerance in the batch.  Divergence is relative to the first
    squared residual seen, only after 8 iterations, threshold 1e5; non-finite values always count as diverged.
    """

    def __init__(self, rhs_norm_sq, rtol, atol, max_iter: np.ndarray, div_threshold=1e5):
        self.max_iter = np.asarray(max_iter)[-1, :]
        self.tol_sq = np.maximum(rtol ** 2 * np.sum(rhs_norm_sq, -1), atol ** 2)   # (batch,)
        self.first = None
        self.div_threshold = div_threshold

    def __call__(self, iterations, rsq):
        rsq = np.abs(rsq)                                                         # (batch, 1)
        with np.errstate(invalid='ignore', divide='ignore', over='ignore'):
            converged = np.all(rsq <= self.tol_sq, axis=1)
            nonfinite = np.any(~np.isfinite(rsq), axis=1)
            if self.first is None:
                self.first = rsq
                diverged = nonfinite
            else:
                diverged = (np.any(rsq / self.first > self.div_threshold, axis=1) & (iterations >= 8)) | nonfinite
        go_on = ~converged & ~diverged & (iterations < self.max_iter)
        return go_on, converged, diverged


def loop_limit(max_iter: np.ndarray):
    """_linalg.py:43-49: the largest limit for one row of limits, the list of checkpoints (equal for all systems) for several."""
    if max_iter.shape[0] == 1:
        return int(np.max(max_iter))
    assert np.all(max_iter == max_iter[:, :1]), "When recording a trajectory, max_iter must be equal for all batch entries"
    return max_iter[:, 0].tolist()


def run_loop(body: Callable, state: tuple, limit) -> tuple:
    """Backend.while_loop (_backend.py:697-735).  One limit: eager loop that tests ``any(continue)`` *before* each body evaluation.  A list of
    checkpoints: the body runs BEFORE the first test, copies are recorded at the checkpoints reached (a repeated checkpoint is recorded once), the
    loop breaks after the pass that left nothing to continue, and the missing records are filled with the LAST RECORDED one (:726-732)."""
    if isinstance(limit, list):
        trj = [state] if 0 in limit else []
        for i in range(1, max(limit) + 1):
            state = body(*state)
            if i in limit:
                trj.append(state)
            if not np.any(state[0]):
                break
        trj.extend([trj[-1]] * (len(limit) - len(trj)))
        return tuple(np.stack([np.asarray(t[k]) for t in trj]) for k in range(len(state)))
    for _ in range(limit):
        if not np.any(state[0]):
            break
        state = body(*state)
    return state


def _prep(y, x0):
    y = np.asarray(y)
    assert y.dtype in (np.float32, np.float64)
    x = np.array(x0, dtype=y.dtype, copy=True)
    return y, x, y.shape[0]


# ----------------------------------------------------------------------------------------------------------------------
# CG (generic, preconditioned)
# ----------------------------------------------------------------------------------------------------------------------

def cg(lin, y, x0, rtol, atol, max_iter, pre=None, name="numpy", trace: Optional[list] = None) -> OracleResult:
    """
    Preconditioned CG.  The monitored quantity is delta = r . M^-1 r (not |r|^2), measured against
    rtol^2 * delta_0 with delta_0 taken from the *initial residual*.
    ``trace`` (optional list) receives per-iteration dicts of the scalars alpha/delta for kernel debugging.
    """
    pre = pre if pre is not None else IdentityPre()
    y, x, nb = _prep(y, x0)
    r = y - matvec(lin, x)
    d = pre.apply(r)
    delta0 = rowdot(r, d)
    check = StopOnL2(np.abs(delta0), rtol, atol, max_iter)
    its = np.zeros(nb, np.int32)
    fev = np.ones(nb, np.int32)
    go, conv, div = check(its, delta0)

    def body(go, x, d, delta, r, its, fev, _c, _d):
        g1 = go.astype(np.int32)
        its = its + g1
        q = matvec(lin, d)
        fev = fev + g1
        dq = rowdot(d, q)
        alpha = safe_div(delta, dq) * g1.astype(y.dtype)[:, None]
        x = x + alpha * d
        r = r - alpha * q
        s = pre.apply(r)
        delta_new = rowdot(r, s)
        d = s + safe_div(delta_new, delta) * d
        if trace is not None:
            trace.append(dict(alpha=alpha[:, 0].copy(), delta=delta_new[:, 0].copy(), dq=dq[:, 0].copy()))
        go, conv, div = check(its, delta_new)
        return go, x, d, delta_new, r, its, fev, conv, div

    _, x, _, _, r, its, fev, conv, div = run_loop(body, (go, x, d, delta0, r, its, fev, conv, div), loop_limit(max_iter))
    return OracleResult(f"Φ-ML CG ({name}) with preconditioner '{pre}'", x, r, its, fev, conv, div, [""] * nb)


# ----------------------------------------------------------------------------------------------------------------------
# CG-adaptive (Hestenes-Stiefel form; the 'auto' method)
# ----------------------------------------------------------------------------------------------------------------------

def cg_adaptive(lin, y, x0, rtol, atol, max_iter, pre=None, name="numpy") -> OracleResult:
    """Ignores preconditioners (falls back to ``cg``), tolerance relative to |y|^2."""
    if pre:
        return cg(lin, y, x0, rtol, atol, max_iter, pre, name)
    y, x, nb = _prep(y, x0)
    r = y - matvec(lin, x)
    d = r
    q = matvec(lin, d)
    its = np.zeros(nb, np.int32)
    fev = np.ones(nb, np.int32)
    rsq = rowdot(r, r)
    check = StopOnL2(rowdot(y, y)[:, 0], rtol, atol, max_iter)
    go, conv, div = check(its, rsq)

    def body(go, x, d, q, r, its, fev, _c, _d):
        g1 = go.astype(np.int32)
        its = its + g1
        dq = rowdot(d, q)
        alpha = safe_div(rowdot(d, r), dq) * g1.astype(y.dtype)[:, None]
        x = x + alpha * d
        r = r - alpha * q
        rsq = rowdot(r, r)
        d = r - safe_div(rowdot(r, q) * d, dq)
        q = matvec(lin, d)
        fev = fev + g1
        go, conv, div = check(its, rsq)
        return go, x, d, q, r, its, fev, conv, div

    _, x, _, _, r, its, fev, conv, div = run_loop(body, (go, x, d, q, r, its, fev, conv, div), loop_limit(max_iter))
    return OracleResult(f"Φ-ML CG-adaptive ({name}) ", x, r, its, fev, conv, div, [""] * nb)


# ----------------------------------------------------------------------------------------------------------------------
# BiCGstab(L), L >= 2
# ----------------------------------------------------------------------------------------------------------------------

def bicgstab_l(lin, y, x0, rtol, atol, max_iter, pre=None, poly_order=2, name="numpy") -> OracleResult:
    """
    Sleijpen/Fokkema BiCGstab(L) as the reference writes it.  Faithfully kept peculiarities:
    * finished systems are NOT frozen (no ``continue`` factor on the step), only their counters stop;
    * ``tau`` is built as ``[[z]*(L+1)]*(L+1)`` in the reference, i.e. all rows are the SAME list, so
      ``tau[i][j]`` really is ``tau[j]`` (last writer wins).  For L=2 this is harmless, for L>=3 it is not.
    * plain divisions (NaN/inf propagate and end the solve through the non-finite test).
    """
    if poly_order == 0:
        raise NotImplementedError("Generic non-stabilized biCG not supported. Use 'scipy-biCG' instead")
    if poly_order == 1:
        return bicgstab_1(lin, y, x0, rtol, atol, max_iter, pre)
    pre = pre if pre is not None else IdentityPre()
    L = poly_order
    y, x, nb = _prep(y, x0)
    r = y - matvec(lin, x)
    r_shadow = r
    its = np.zeros(nb, np.int32)
    fev = np.ones(nb, np.int32)
    check = StopOnL2(rowdot(y, y)[:, 0], rtol, atol, max_iter)
    go, conv, div = check(its, rowdot(r, r))
    one = np.ones((nb, 1), y.dtype)
    zero = np.zeros((nb, 1), y.dtype)
    rho0, rho1, omega, alpha = one, one, one, zero
    u = np.zeros_like(x)
    rh = [np.zeros_like(x) for _ in range(L + 1)]
    uh = [np.zeros_like(x) for _ in range(L + 1)]

    def body(go, x, r, its, fev, _c, _d, rho0, rho1, omega, alpha, u, uh, rh):
        uh, rh = list(uh), list(rh)
        tau = [zero] * (L + 1)          # aliased rows of the reference: tau[i][j] == tau[j]
        sigma = [zero] * (L + 1)
        gam = [zero] * (L + 1)
        gam_p = [zero] * (L + 1)
        gam_pp = [zero] * (L + 1)
        g1 = go.astype(np.int32)
        its = its + g1
        uh[0] = u
        rh[0] = r
        with np.errstate(divide='ignore', invalid='ignore', over='ignore'):
            rho0 = -omega * rho0
            for j in range(L):                      # Bi-CG part
                rho1 = rowdot(rh[j], r_shadow)
                beta = alpha * rho1 / rho0
                rho0 = rho1
                for i in range(j + 1):
                    uh[i] = rh[i] - beta * uh[i]
                uh[j + 1] = matvec(lin, pre.apply(uh[j]))
                fev = fev + g1
                alpha = rho0 / rowdot(uh[j + 1], r_shadow)
                for i in range(j + 1):
                    rh[i] = rh[i] - alpha * uh[i + 1]
                rh[j + 1] = matvec(lin, pre.apply(rh[j]))
                fev = fev + g1
                x = x + alpha * uh[0]
            for j in range(1, L + 1):               # modified Gram-Schmidt
                for i in range(1, j):
                    tau[j] = rowdot(rh[j], rh[i]) / sigma[i]
                    rh[j] = rh[j] - tau[j] * rh[i]
                sigma[j] = rowdot(rh[j], rh[j])
                gam_p[j] = rowdot(rh[0], rh[j]) / sigma[j]
            omega = gam[L] = gam_p[L]               # MR part
            for j in range(L - 1, 0, -1):
                acc = np.zeros_like(zero)
                for i in range(j + 1, L + 1):
                    acc = acc + tau[i] * gam[i]
                gam[j] = gam_p[j] - acc
            for j in range(1, L):
                acc = np.zeros_like(zero)
                for i in range(j + 1, L):
                    acc = acc + tau[i] * gam[i + 1]
                gam_pp[j] = gam[j + 1] + acc
            x = x + gam[1] * rh[0]                   # update
            rh[0] = rh[0] - gam_p[L] * rh[L]
            uh[0] = uh[0] - gam[L] * uh[L]
            for j in range(1, L):
                uh[0] = uh[0] - gam[j] * uh[j]
                x = x + gam_pp[j] * rh[j]
                rh[0] = rh[0] - gam_p[j] * rh[j]
            u = uh[0]
            r = rh[0]
            rsq = rowdot(r, r)
        go, conv, div = check(its, rsq)
        return go, x, r, its, fev, conv, div, rho0, rho1, omega, alpha, u, uh, rh

    out = run_loop(body, (go, x, r, its, fev, conv, div, rho0, rho1, omega, alpha, u, uh, rh), loop_limit(max_iter))
    _, x, r, its, fev, conv, div = out[:7]
    return OracleResult(f"Φ-ML biCG-stab({L}) ({name}) with preconditioner '{pre}'", x, r, its, fev, conv, div, [""] * nb)


# ----------------------------------------------------------------------------------------------------------------------
# BiCGstab(1)
# ----------------------------------------------------------------------------------------------------------------------

def bicgstab_1(lin, y, x0, rtol, atol, max_iter, pre=None) -> OracleResult:
    """Classic BiCGstab; the shadow residual is a vector of ones (not r0); steps are not frozen after convergence."""
    pre = pre if pre is not None else IdentityPre()
    y, x, nb = _prep(y, x0)
    r = y - matvec(lin, x)
    shadow = np.ones_like(x)
    its = np.zeros(nb, np.int32)
    fev = np.ones(nb, np.int32)
    check = StopOnL2(rowdot(y, y)[:, 0], rtol, atol, max_iter)
    go, conv, div = check(its, rowdot(r, r))
    one = np.ones((nb, 1), y.dtype)
    rho_prev, rho, omega, alpha = one, one, one, np.zeros((nb, 1), y.dtype)
    u = np.zeros_like(x)
    p = np.zeros_like(x)

    def body(go, x, r, its, fev, _c, _d, rho_prev, rho, omega, alpha, u, p):
        g1 = go.astype(np.int32)
        its = its + g1
        with np.errstate(divide='ignore', invalid='ignore', over='ignore'):
            rho = rowdot(shadow, r)
            beta = rho / rho_prev * alpha / omega
            rho_prev = rho
            p = r + beta * (p - omega * u)
            ph = pre.apply(p)
            u = matvec(lin, ph)
            fev = fev + g1
            alpha = rho / rowdot(shadow, u)
            h = x + alpha * ph
            s = r - alpha * u
            s_l = pre.apply_inv_l(s)
            z = pre.apply(s)
            t = matvec(lin, z)
            fev = fev + g1
            t_l = pre.apply_inv_l(t)
            omega = rowdot(t_l, s_l) / rowdot(t_l, t_l)
            x = h + omega * z
            r = s - omega * t
            rsq = rowdot(r, r)
        go, conv, div = check(its, rsq)
        return go, x, r, its, fev, conv, div, rho_prev, rho, omega, alpha, u, p

    out = run_loop(body, (go, x, r, its, fev, conv, div, rho_prev, rho, omega, alpha, u, p), loop_limit(max_iter))
    _, x, r, its, fev, conv, div = out[:7]
    return OracleResult(f"Φ-ML biCG-stab with preconditioner '{pre}'", x, r, its, fev, conv, div, [""] * nb)


# ----------------------------------------------------------------------------------------------------------------------
# torch-only fast paths (un-preconditioned, explicit matrix)
# ----------------------------------------------------------------------------------------------------------------------

def _torch_fastpath(lin, y, x0, rtol, atol, max_iter, adaptive: bool) -> OracleResult:
    """
    ``torch_sparse_cg`` / ``torch_sparse_cg_adaptive``: tolerance from |y|^2, true-residual refresh every 20th
    loop pass (adds 1 to function_evaluations of ALL systems), divergence threshold 100, initial ``diverged``
    from non-finite x0, loop runs ``while not all(finished)`` without an outer max-iteration bound.
    """
    y, x, nb = _prep(y, x0)
    tol_sq = np.maximum(rtol ** 2 * np.sum(y ** 2, -1), atol ** 2)       # (batch,)
    mi = np.asarray(max_iter)[0]
    r = y - matvec(lin, x)
    d = r
    passes = 0
    its = np.zeros(nb, np.int32)
    fev = np.ones(nb, np.int32)
    rsq = rsq0 = rowdot(r, r)
    div = np.any(~np.isfinite(x), axis=1)
    conv = np.all(rsq <= tol_sq, axis=1)
    done = conv | div | (its >= mi)
    live = (~done).astype(np.int32)
    while not np.all(done):
        passes += 1
        its = its + live
        q = matvec(lin, d)
        fev = fev + live
        dq = rowdot(d, q)
        num = rowdot(d, r) if adaptive else rsq
        alpha = safe_div(num, dq) * live.astype(y.dtype)[:, None]
        x = x + alpha * d
        if passes % 20 == 0:
            r = y - matvec(lin, x)
            fev = fev + 1
        else:
            r = r - alpha * q
        rsq_old = rsq
        rsq = rowdot(r, r)
        if adaptive:
            d = r - safe_div(rowdot(r, q) * d, dq)
        else:
            d = r + safe_div(rsq, rsq_old) * d
        with np.errstate(divide='ignore', invalid='ignore', over='ignore'):
            div = np.any(rsq / rsq0 > 100, axis=1) & (its >= 8)
        conv = np.all(rsq <= tol_sq, axis=1)
        done = conv | div | (its >= mi)
        live = (~done).astype(np.int32)
    return OracleResult("Φ-ML CG (PyTorch*)", x, r, its, fev, conv, div, [""] * nb)


def torch_fastpath_cg(lin, y, x0, rtol, atol, max_iter):
    return _torch_fastpath(lin, y, x0, rtol, atol, max_iter, adaptive=False)


def torch_fastpath_cg_adaptive(lin, y, x0, rtol, atol, max_iter):
    return _torch_fastpath(lin, y, x0, rtol, atol, max_iter, adaptive=True)


# ----------------------------------------------------------------------------------------------------------------------
# dispatch
# ----------------------------------------------------------------------------------------------------------------------

def linear_solve(method: str, lin, y, x0, rtol, atol, max_iter, pre=None, flavour: str = "numpy") -> OracleResult:
    """
    Method-string dispatch of Backend.linear_solve.  ``flavour='torch'`` selects the torch fast paths where the
    reference torch backend would (no preconditioner, explicit matrix, single checkpoint); ``'numpy'`` always
    runs the generic loops.
    """
    rtol = np.asarray(rtol, dtype=np.asarray(y).dtype)
    atol = np.asarray(atol, dtype=np.asarray(y).dtype)
    max_iter = np.asarray(max_iter)
    fast = flavour == "torch" and not pre and not callable(lin) and max_iter.shape[0] == 1
    if method in ('auto', 'CG-adaptive'):
        if fast:
            return torch_fastpath_cg_adaptive(lin, y, x0, rtol, atol, max_iter)
        return cg_adaptive(lin, y, x0, rtol, atol, max_iter, pre, flavour)
    if method == 'CG':
        if fast:
            return torch_fastpath_cg(lin, y, x0, rtol, atol, max_iter)
        return cg(lin, y, x0, rtol, atol, max_iter, pre, flavour)
    if method in ('biCG', 'biCG-stab(0)'):
        return bicgstab_l(lin, y, x0, rtol, atol, max_iter, pre, 0, flavour)
    if method == 'biCG-stab':
        return bicgstab_l(lin, y, x0, rtol, atol, max_iter, pre, 1, flavour)
    if method.startswith('biCG-stab('):
        return bicgstab_l(lin, y, x0, rtol, atol, max_iter, pre, int(method[len('biCG-stab('):-1]), flavour)
    raise NotImplementedError(f"Method '{method}' not supported for linear solve.")
