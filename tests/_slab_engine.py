"""
TEST DOUBLE (tests only): the five steps of the slab engine on torch CPU tensors, restating the fused two-kernel CG iteration
of phiml_b200/csrc (star_tma.cuh + CgResid) on an extended local array.  Lets the gloo tests drive phiml_b200.multi_gpu.SlabCG
(partitioning, halo exchange, all-reduce sequence, polling) without a GPU.
"""
import numpy as np
import torch


class CpuSlabEngine:
    def __init__(self, ext_grid, xb, xe, offsets, coeffs, dtype, jacobi):
        self.g = tuple(ext_grid)
        self.xb, self.xe = xb, xe
        self.dtype = dtype
        self.coef = {tuple(o): float(c) for o, c in zip(offsets, coeffs)}
        self.cinv = 1.0 / self.coef[(0, 0, 0)] if jacobi else 1.0

    def _apply(self, v):
        """A v on the owned planes of an extended (batch, nx+2, ny, nz) array; ZERO boundaries in y, z; x halos are data."""
        b = v.shape[0]
        a = v.view(b, *self.g)
        out = torch.zeros_like(a)
        sl = slice(self.xb, self.xe)
        pad = torch.nn.functional.pad(a, (1, 1, 1, 1))            # zero pad y and z
        for (ox, oy, oz), c in self.coef.items():
            out[:, sl] += c * pad[:, self.xb + ox:self.xe + ox, 1 + oy:1 + oy + self.g[1], 1 + oz:1 + oz + self.g[2]]
        return out.view(b, -1)

    def _own(self, v):
        b = v.shape[0]
        return v.view(b, *self.g)[:, self.xb:self.xe].reshape(b, -1)

    def start(self, y_ext, x_ext, rtol, atol, max_iter):
        nb = y_ext.shape[0]
        self.nb = nb
        self.y, self.x = y_ext.clone(), x_ext.clone()
        self.rtol = torch.as_tensor(rtol, dtype=self.dtype).reshape(-1).expand(nb).clone()
        self.atol = torch.as_tensor(atol, dtype=self.dtype).reshape(-1).expand(nb).clone()
        self.max_iter = int(max_iter)
        self.q = self._apply(self.x)
        self.r = torch.zeros_like(self.y)
        self.r.view(nb, *self.g)[:, self.xb:self.xe] = (self.y - self.q).view(nb, *self.g)[:, self.xb:self.xe]
        self.d0 = torch.zeros_like(self.y)
        self.d0.view(nb, *self.g)[:, self.xb:self.xe] = self.cinv * self.r.view(nb, *self.g)[:, self.xb:self.xe]
        self.d1 = torch.zeros_like(self.y)
        self.passes = 0
        ro, yo = self._own(self.r), self._own(self.y)
        self.totals = torch.zeros(nb * 4, dtype=torch.float64)
        t = self.totals.view(nb, 4)
        t[:, 0] = (ro * (self.cinv * ro)).sum(1, dtype=torch.float64)
        t[:, 1] = (yo * yo).sum(1, dtype=torch.float64)
        t[:, 2] = (ro * ro).sum(1, dtype=torch.float64)

    def _check(self, first):
        rsq = self.delta.abs()
        conv = rsq <= self.tol_min
        nonfin = ~torch.isfinite(rsq)
        if first:
            self.rsq0 = rsq.clone()
            div = nonfin
            upd = torch.ones_like(conv)
        else:
            div = ((rsq / self.rsq0 > 1e5) & (self.its >= 8)) | nonfin
            upd = self.go.clone()
        go = ~conv & ~div & (self.its < self.max_iter)
        self.conv = torch.where(upd, conv, self.conv) if not first else conv
        self.div = torch.where(upd, div, self.div) if not first else div
        self.go = torch.where(upd, go, self.go) if not first else go

    def step(self, code):
        nb = self.nb
        t = self.totals.view(nb, 4)
        if code == 10:
            self.delta = t[:, 0].to(self.dtype).clone()          # .to() may alias the totals buffer
            tol = torch.maximum(self.rtol ** 2 * self.delta.abs(), self.atol ** 2)
            self.tol_min = tol.min()
            self.its = torch.zeros(nb, dtype=torch.int32)
            self.fev = torch.ones(nb, dtype=torch.int32)
            self.alpha = torch.zeros(nb, dtype=self.dtype)
            self.beta = torch.zeros(nb, dtype=self.dtype)
            self._check(True)
            return
        if not bool(self.go.any()) and code in (0, 1, 2):
            return
        d_old, d_new = (self.d1, self.d0) if self.passes % 2 else (self.d0, self.d1)
        act = self.go[:, None]
        if code == 0:
            new = self.cinv * self.r + self.beta[:, None] * d_old            # all planes incl. halos (halo r is exchanged data)
            d_new.copy_(torch.where(act, new, d_new))
            xo = self.x.view(nb, *self.g)
            xo[:, self.xb:self.xe] += torch.where(act, self.alpha[:, None] * d_old, torch.zeros_like(d_old)).view(nb, *self.g)[:, self.xb:self.xe]
            self.q = torch.where(act, self._apply(d_new), self.q)
            t[:, 0] = (self._own(d_new) * self._own(self.q)).sum(1, dtype=torch.float64)
        elif code == 1:
            dq = t[:, 0].to(self.dtype)
            self.its += self.go.to(torch.int32)
            self.fev += self.go.to(torch.int32)
            self.alpha = torch.where(self.go, torch.where(dq == 0, torch.zeros_like(dq), self.delta / dq), self.alpha)
            ro = self.r.view(nb, *self.g)
            ro[:, self.xb:self.xe] -= torch.where(act, self.alpha[:, None] * self.q, torch.zeros_like(self.q)).view(nb, *self.g)[:, self.xb:self.xe]
            r_own = self._own(self.r)
            t[:, 0] = (r_own * (self.cinv * r_own)).sum(1, dtype=torch.float64)
        elif code == 2:
            dn = t[:, 0].to(self.dtype)
            beta = torch.where(self.delta == 0, torch.zeros_like(dn), dn / self.delta)
            self.beta = torch.where(self.go, beta, self.beta)
            self.delta = torch.where(self.go, dn, self.delta)
            self._check(False)
            self.passes += 1
        elif code == 20:
            d_cur = self.d1 if self.passes % 2 else self.d0
            xo = self.x.view(nb, *self.g)
            xo[:, self.xb:self.xe] += (self.alpha[:, None] * d_cur).view(nb, *self.g)[:, self.xb:self.xe]
            self.alpha = torch.zeros_like(self.alpha)

    def any_continue(self):
        return bool(self.go.any())

    def result(self):
        return self.its, self.fev, self.conv, self.div


class CpuCallbackEngine:
    """
    TEST DOUBLE of phiml_b200.multi_gpu.CudaCallbackEngine: runs the ORACLE's Krylov loops (oracle/krylov.py) on this rank's
    slab with the two callbacks of the distributed mode — ``halo(vec)`` before every product, ``allreduce(totals)`` for every
    dot product — so that the gloo tests exercise SlabSolver's partitioning, halo exchange and reduction plumbing for any method.
    """

    def __init__(self, ext_grid, xb, xe, offsets, coeffs, dtype, method, jacobi):
        self.base = CpuSlabEngine(ext_grid, xb, xe, offsets, coeffs, dtype, jacobi)
        self.method, self.jacobi, self.dtype = method, jacobi, dtype

    def solve(self, y_ext, x_ext, rtol, atol, max_iter, allreduce, halo, fixed_iterations=None):
        from oracle import krylov, precond
        base = self.base
        nb = y_ext.shape[0]
        g = base.g

        def embed(v_own):
            ext = torch.zeros((nb, g[0] * g[1] * g[2]), dtype=self.dtype)
            ext.view(nb, *g)[:, base.xb:base.xe] = torch.as_tensor(v_own).view(nb, base.xe - base.xb, g[1], g[2])
            return ext

        def lin(v_own):
            ext = embed(v_own)
            halo(ext)
            return base._own(base._apply(ext)).numpy()

        def rowdot(a, b):
            t = torch.as_tensor(np.sum(a * b, axis=-1, keepdims=True), dtype=torch.float64).reshape(-1).clone()
            allreduce(t)
            return t.numpy().reshape(nb, 1).astype(a.dtype)

        class Jac:
            def apply(self_, v): return v * np.asarray(base.cinv, dtype=v.dtype)
            apply_inv_l = apply
            def apply_inv_u(self_, v): return v
            def __repr__(self_): return "jacobi"

        saved = krylov.rowdot
        krylov.rowdot = rowdot
        try:
            np_dt = np.float64 if self.dtype == torch.float64 else np.float32
            res = krylov.linear_solve(self.method, lin, base._own(y_ext).numpy(), base._own(x_ext).numpy(), np.asarray(rtol, np_dt), np.asarray(atol, np_dt),
                                      np.full((1, nb), max_iter), Jac() if self.jacobi else None, flavour='numpy')
        finally:
            krylov.rowdot = saved
        return (embed(res.x), embed(res.residual), torch.as_tensor(res.iterations), torch.as_tensor(res.function_evaluations),
                torch.as_tensor(res.converged), torch.as_tensor(res.diverged))
