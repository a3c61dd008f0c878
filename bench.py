#!/usr/bin/env python
"""
bench.py — BASELINE.json metric: CG iterations/s (and HBM GB/s as a fraction of the measured peak) of the
Jacobi-preconditioned CG solve of the 3-D Poisson system 256^3 (7-point Laplacian, ZERO padding) in fp32.

A *step* is one CG iteration (one pass of the hot loop: SpMV + fused dots, x/r update + preconditioner + dot,
direction update, on-device convergence test) of ONE 256^3 system.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--format stencil|csr] [--n 256] [--extras 0|1]

N = 1: the system lives on one GPU (two fused kernels per iteration).
N > 1 (under torchrun, one rank per GPU): the SAME system is row-partitioned into x-slabs over the N GPUs (strong scaling):
per iteration two kernels per GPU which exchange the residual's boundary planes by peer-memory stores over NVLink and all-reduce
the two dot products through peer-memory mailboxes inside the kernels (phiml_b200.multi_gpu.SlabCG) — a collective in every step.

Besides the headline the JSON line carries the other BASELINE configs as blocks (``configs``): ``slab_1024`` (north-star scaling
case, 1024^3 Jacobi-CG on N GPUs), ``sharded_cfg3`` (4096 x 128^2 systems, batch sharded over N GPUs), ``slab_cfg5`` (1024^3
nonsymmetric BiCGstab(2) on x-slabs) and, at N = 1, ``cfg4_ilu`` (ILU(0)-CG with the structured triangular solves).

Rank 0 prints ONE JSON line.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

METRIC = "cg_iterations_per_second"
UNIT = "iterations/s"


def parse():
    p = argparse.ArgumentParser()
    p.add_argument('--gpus', type=int, default=1)
    p.add_argument('--steps', type=int, default=500)
    p.add_argument('--warmup', type=int, default=20)
    p.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    p.add_argument('--format', default='stencil', choices=['stencil', 'csr'])
    p.add_argument('--n', type=int, default=256)
    p.add_argument('--extras', type=int, default=1, help='1: also time the other BASELINE configs (blocks under "configs")')
    p.add_argument('--big', type=int, default=1024, help='grid size of the slab_1024 / slab_cfg5 blocks')
    p.add_argument('--extras-budget', type=float, default=240.0, help='seconds the extra blocks may take before the line is printed without them')
    p.add_argument('--no-cpu-baseline', action='store_true')
    p.add_argument('--no-e2e', action='store_true')
    return p.parse_args()


def measured_peak():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    try:
        with open(path) as f:
            return float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
    except Exception:
        return 6650.0, 'fallback (B200_PROFILING.md)'


def ncu_traffic(fmt, n):
    """DRAM bytes per launch of the two kernels of the iteration from the committed ncu --set full capture (256^3 only)."""
    if n != 256:
        return {}
    try:
        with open(os.path.join(ROOT, 'profiles', 'ncu_traffic.json')) as f:
            return json.load(f).get(fmt, {})
    except Exception:
        return {}


def workload_config(n, n_gpus):
    """Identical in both arms (ours / reference)."""
    return {"workload": f"3D Poisson {n}^3 7-point Laplacian (ZERO padding), Jacobi-preconditioned CG fp32, rel_tol 1e-5",
            "rows": n ** 3, "nnz": 7 * n ** 3 - 6 * n ** 2, "systems": 1,
            "parallelism": "one GPU" if n_gpus == 1 else f"one system row-partitioned into x-slabs over {n_gpus} GPUs",
            "l2": "working set (5 vectors x %.0f MB) exceeds the 126 MB L2 at 1 GPU; no flush needed" % (n ** 3 * 4 / 1e6)}


# ----------------------------------------------------------------------------------------------------------------------
# CPU side: the reference's own NumPy/SciPy path (baseline/_ref) or, without it, the oracle port.  1 core either way.
# ----------------------------------------------------------------------------------------------------------------------

def cpu_poisson_csr(n, dtype=np.float32):
    """Canonical CSR of -laplace on n^3 through Kronecker sums (identical arrays to oracle.assemble.poisson_csr, far less memory)."""
    import scipy.sparse as sp
    one = sp.diags([-np.ones(n - 1), 2 * np.ones(n), -np.ones(n - 1)], [-1, 0, 1], format='csr', dtype=np.float64)
    eye = sp.identity(n, format='csr', dtype=np.float64)
    a = sp.kron(sp.kron(one, eye), eye) + sp.kron(sp.kron(eye, one), eye) + sp.kron(sp.kron(eye, eye), one)
    a = a.tocsr()
    a.sort_indices()
    a = a.astype(dtype)
    a.indices = a.indices.astype(np.int32)
    a.indptr = a.indptr.astype(np.int32)
    return a


class TimedTrace(list):
    def append(self, item):
        item['t'] = time.perf_counter()
        super().append(item)


def cpu_port_iterations_per_second(n, warmup, steps):
    """Oracle port: warmup+steps Jacobi-CG iterations, the timed region covers exactly `steps` of them."""
    from oracle import krylov, precond
    a = cpu_poisson_csr(n)
    y = np.random.default_rng(0).standard_normal((1, n ** 3)).astype(np.float32)
    trace = TimedTrace()
    w = max(1, warmup)            # the trace is stamped at the END of every iteration: the timed region starts at the end of iteration w
    total = w + steps
    krylov.cg(a, y, np.zeros_like(y), np.float32([1e-5]), np.float32([0]), np.array([[total]]), precond.Jacobi(a), trace=trace)
    ts = [t['t'] for t in trace]
    done = len(ts)
    w = min(w, max(1, done - 1))
    its = done - w                # == steps unless the solve converged earlier
    dt = ts[-1] - ts[w - 1]
    return its / dt, its, dt


def cpu_reference_iterations_per_second(n, warmup, steps):
    """The UNMODIFIED reference (baseline/_ref): NUMPY.linear_solve('CG', scipy csr, ...) = phiml/backend/_linalg.py:cg on SciPy's
    csr_matvec, driven with a Jacobi object (the reference has no 'jacobi' by name).  Two calls, W and W+K iterations:
    the difference is K iterations of the loop (set-up and first residual cancel)."""
    from _phiml import import_phiml
    if import_phiml() is None:
        return None
    from phiml.backend import NUMPY
    from phiml.backend._backend import Preconditioner

    class RefJacobi(Preconditioner):
        def __init__(self, csr): self.inv = (1 / csr.diagonal()).astype(csr.dtype)
        def apply(self, vec): return vec * self.inv
        def apply_inv_l(self, vec): return vec * self.inv
        def apply_inv_u(self, vec): return vec
        def __repr__(self): return "jacobi"

    a = cpu_poisson_csr(n)
    y = np.random.default_rng(0).standard_normal((1, n ** 3)).astype(np.float32)
    pre = RefJacobi(a)
    w = max(1, warmup)

    def run(iters):
        t0 = time.perf_counter()
        with NUMPY:
            r = NUMPY.linear_solve('CG', a, y, np.zeros_like(y), np.float32([1e-5]), np.float32([0]), np.array([[iters]]), pre, None)
        return time.perf_counter() - t0, int(r.iterations[0])
    t_w, i_w = run(w)
    t_all, i_all = run(w + steps)
    its = i_all - i_w
    dt = t_all - t_w
    return its / dt, its, dt


def cpu_baseline(n, warmup, steps):
    r = None
    try:
        r = cpu_reference_iterations_per_second(n, warmup, steps)
    except Exception as exc:           # the reference copy is optional; the port always exists
        sys.stderr.write(f"[bench] reference arm unavailable ({exc!r}); timing the oracle port\n")
    if r is not None:
        v, its, dt = r
        return v, its, dt, "reference", (f"{its} Jacobi-CG iterations of the {n}^3 fp32 system by the unmodified reference (baseline/_ref: phiml.backend.NUMPY.linear_solve "
                                         f"-> _linalg.cg on SciPy csr_matvec; NumPy ufuncs and csr_matvec are single-threaded, so 1 core is all the path can use), {dt:.1f} s")
    v, its, dt = cpu_port_iterations_per_second(n, warmup, steps)
    return v, its, dt, "port", (f"{its} Jacobi-CG iterations of the {n}^3 fp32 system with the oracle port of phiml/backend/_linalg.py:52-90 (NumPy ufuncs + SciPy csr_matvec, "
                                f"single-threaded like the reference's NumPy backend), {dt:.1f} s")


# ----------------------------------------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------------------------------------

class ClockSampler:
    FIELDS = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.FIELDS}', '--format=csv,noheader,nounits', '-lms', '50', '-i', str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def __exit__(self, *exc):
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for line in self.lines:
            parts = [p.strip() for p in line.split(',')]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[2:6]):
                if v.lower().startswith('active'):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------------------------------

class Ctx:
    """Per-process context of the GPU arm."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get('WORLD_SIZE', '1'))
        self.rank = int(os.environ.get('RANK', '0'))
        self.local_rank = int(os.environ.get('LOCAL_RANK', '0'))
        assert self.world == args.gpus or self.world == 1, f"--gpus {args.gpus} but WORLD_SIZE={self.world}"
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device('cuda', self.local_rank)
        if self.world > 1:
            dist.init_process_group('nccl', device_id=self.dev)
        self.peak, self.peak_src = measured_peak()
        self.gen = torch.Generator(device=self.dev)
        self.gen.manual_seed(1234 + self.rank)

    def fence(self):
        """barrier + synchronize on both sides of a timed region"""
        self.torch.cuda.synchronize(self.dev)
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize(self.dev)

    def max_over_ranks(self, v):
        if self.world == 1:
            return float(v)
        t = self.torch.tensor([float(v)], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t[0])

    def timed(self, fn, reps=1):
        """CUDA events on the launch stream around fn(), barrier + sync on both sides, max over ranks; list of ms (one per rep)."""
        torch = self.torch
        out = []
        for _ in range(reps):
            self.fence()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize(self.dev)
            ms = e0.elapsed_time(e1)
            self.fence()
            out.append(self.max_over_ranks(ms))
        return out


def laplace_stencil():
    from phiml_b200 import stencils
    sh = stencils.neg_laplace_pairs(3)
    return [s for s, _ in sh], [c for _, c in sh]


def time_jacobi_cg(ctx, n, warmup, steps, want_profile=False):
    """K Jacobi-CG iterations of ONE n^3 Poisson system on ctx.world GPUs with rtol = atol = 0 (the convergence test runs but never
    fires).  Returns a dict; ms = median over repetitions of exactly `steps` iterations each."""
    import phiml_b200 as pb
    from phiml_b200 import _engine as E
    from phiml_b200.multi_gpu import SlabCG
    torch = ctx.torch
    offsets, coeffs = laplace_stencil()
    N, w = n ** 3, 4
    plane = n * n
    out = {}
    prof = None
    if ctx.world == 1:
        stencil = pb.Matrix.stencil_matrix((n, n, n), offsets, coeffs, torch.float32, ctx.dev)
        pre = pb.JacobiPreconditioner(stencil)
        y = torch.randn((1, N), dtype=torch.float32, device=ctx.dev, generator=ctx.gen)
        x0 = torch.zeros_like(y)
        zero = torch.zeros(1, dtype=torch.float32, device=ctx.dev)
        mi = torch.full((1,), 2 ** 30, dtype=torch.int32, device=ctx.dev)
        solver = pb.DeviceSolver(stencil, pre, E.CG, 0, 1)
        solver.start(y, x0, zero, zero, mi)
        solver.advance(warmup)
        est = ctx.timed(lambda: solver.advance(steps))[0]
        reps = int(min(25, max(1, np.ceil(60.0 / max(est, 1e-3)))))
        launches0 = pb.launch_count()
        ms_all = ctx.timed(lambda: solver.advance(steps), reps)
        launches = (pb.launch_count() - launches0) // reps
        done = int(solver.result()[0][0])
        assert done == warmup + steps * (reps + 1), f"device executed {done} iterations, expected {warmup + steps * (reps + 1)}"
        assert solver.any_continue()
        if want_profile:
            solver.profile(True)
            solver.advance(60)
            prof = solver.profile_read()
            solver.profile(False)
        del solver
    else:
        cg = SlabCG((n, n, n), offsets, coeffs, torch.float32, jacobi=True, device=ctx.dev)
        y = torch.randn((1, cg.nxl * plane), dtype=torch.float32, device=ctx.dev, generator=ctx.gen)
        cg.begin(y, torch.zeros_like(y), [0.0], [0.0], 2 ** 30)
        assert cg.peer_halos_active and cg.peer_reduce_active, "multi-GPU bench needs the peer-memory halo stores and the in-kernel all-reduce"
        cg.iterate(warmup)
        est = ctx.timed(lambda: cg.iterate(steps))[0]
        reps = int(min(25, max(1, np.ceil(60.0 / max(est, 1e-3)))))
        launches0 = pb.launch_count()
        ms_all = ctx.timed(lambda: cg.iterate(steps), reps)
        launches = (pb.launch_count() - launches0) // reps
        if want_profile:
            cg.engine.solver.profile(True)
            cg.iterate(60)
            prof = cg.engine.solver.profile_read()
            cg.engine.solver.profile(False)
        x, r, its, fev, conv, div = cg.finish()
        done = int(its[0])
        expect = warmup + steps * (reps + 1) + (60 if want_profile else 0)
        assert done == expect, f"device executed {done} iterations, expected {expect}"
        out["collectives_per_iteration"] = {"all_reduce_in_kernel": 2, "halo_planes_stored_to_peers": 2 if ctx.world > 2 else 1}
        out["nvlink_bytes_per_iteration_per_gpu"] = {"halo_stores": 2 * plane * w, "mailbox_stores": 2 * ctx.world * 64}
        del cg
    ms = float(np.median(ms_all))
    its_s = steps / (ms * 1e-3)
    bytes_design = 9 * N * w
    out.update({"n": n, "gpus": ctx.world, "iterations_per_s": its_s, "ms_per_iteration": ms / steps, "repeats": len(ms_all), "ms_min_max": [min(ms_all) / steps, max(ms_all) / steps],
                "bytes_moved_by_design": bytes_design, "hbm_gbs_per_gpu": bytes_design / ctx.world * its_s / 1e9,
                "hbm_frac_of_peak_per_gpu": bytes_design / ctx.world * its_s / 1e9 / ctx.peak,
                "survey_model_bytes": 13 * N * w, "survey_model_frac_of_peak_per_gpu": 13 * N * w / ctx.world * its_s / 1e9 / ctx.peak})
    torch.cuda.empty_cache()
    return out, ms, launches, prof


def time_cfg5(ctx, n, iters):
    """BASELINE config 5: nonsymmetric advection-diffusion, BiCGstab(2) fp32, ONE n^3 system on x-slabs (any world size)."""
    import phiml_b200 as pb
    from phiml_b200 import _engine as E
    from phiml_b200 import stencils
    from phiml_b200.multi_gpu import SlabSolver
    torch = ctx.torch
    A = stencils.ADVDIFF
    sh = stencils.advection_diffusion_pairs(3, A['velocity'], A['c'], A['nu'], np.float32)
    offsets, coeffs = [s for s, _ in sh], [float(c) for _, c in sh]
    N, plane = n ** 3, n * n
    if ctx.world == 1:
        st = pb.Matrix.stencil_matrix((n, n, n), offsets, coeffs, torch.float32, ctx.dev)
        y = torch.randn((1, N), dtype=torch.float32, device=ctx.dev, generator=ctx.gen)
        zero = torch.zeros(1, dtype=torch.float32, device=ctx.dev)
        mi = torch.full((1,), 2 ** 30, dtype=torch.int32, device=ctx.dev)
        solver = pb.DeviceSolver(st, None, E.BICGSTAB, 2, 1)
        solver.start(y, torch.zeros_like(y), zero, zero, mi)
        solver.advance(2)
        ms = min(ctx.timed(lambda: solver.advance(iters), 2))
        per_it = ms / iters
        extra = {"collectives_per_iteration": None}
        del solver
    else:
        sol = SlabSolver((n, n, n), offsets, coeffs, torch.float32, method='biCG-stab(2)', jacobi=False, device=ctx.dev)
        y = torch.randn((1, sol.nxl * plane), dtype=torch.float32, device=ctx.dev, generator=ctx.gen)
        x0 = torch.zeros_like(y)
        def run(k):
            # CUDA events around the loop alone (set-up: allocation + IPC mapping stays outside); divided by the iterations the loop really made
            # (far beyond convergence BiCGstab breaks down and the device loop stops by itself)
            its = sol.solve(y, x0, [0.0], [0.0], 2 ** 30, fixed_iterations=k)[2]
            return ctx.max_over_ranks(sol.engine.last_loop_ms) / max(1, int(its.max()))
        run(2)
        c0 = dict(sol.counts)
        t0 = time.perf_counter()
        per1 = min(run(iters) for _ in range(2))
        call_ms = (time.perf_counter() - t0) * 1e3 / 2
        c1 = dict(sol.counts)
        per2 = run(2 * iters)
        c2 = dict(sol.counts)
        per_it = min(per1, per2)
        ms1 = per1 * iters
        if getattr(sol, 'native', False):    # nothing is called on the host: BiCGstab(2) has 9 reductions and 4 products per iteration (csrc/solver.cu pass_bicg_l)
            coll = {"all_reduce_in_kernel": 9, "halo_pushes_to_peers": 4, "host_callbacks": 0}
        else:
            coll = {"all_reduce": ((c2['allreduce'] - c1['allreduce']) - (c1['allreduce'] - c0['allreduce'])) / (2 * iters),
                    "halo_exchanges": ((c2['halo'] - c1['halo']) - (c1['halo'] - c0['halo'])) / (2 * iters)}
        extra = {"collectives_per_iteration": coll,
                 "nvlink_bytes_per_iteration_per_gpu": 4 * 2 * plane * 4, "setup_ms": call_ms - ms1,
                 "timing": "CUDA events around the fixed-iteration loop of every rank, max over ranks"}
        del sol
    model = 53 * N * 4
    out = {"workload": f"3D advection-diffusion {n}^3 nonsymmetric 7-point, BiCGstab(2) fp32, x-slabs", "n": n, "gpus": ctx.world, "iterations_per_s": 1e3 / per_it,
           "ms_per_iteration": per_it, "survey_model_bytes": model, "survey_model_gbs_per_gpu": model / ctx.world / (per_it * 1e-3) / 1e9,
           "survey_model_frac_of_peak_per_gpu": model / ctx.world / (per_it * 1e-3) / 1e9 / ctx.peak}
    out.update(extra)
    torch.cuda.empty_cache()
    return out


def time_cfg3(ctx, batch=4096, n=128):
    """BASELINE config 3: `batch` independent 2-D Poisson systems sharing one 128x128 matrix, CG fp32 to rel_tol 1e-5, the batch
    sharded over the GPUs (one all-reduce(min) of the tolerance, no collective in the loop).  Strong scaling of the WHOLE batch."""
    import phiml_b200 as pb
    from phiml_b200 import solve as S
    from phiml_b200 import stencils
    from phiml_b200.multi_gpu import solve_sharded, shard_bounds
    torch, dist = ctx.torch, ctx.dist
    sh = stencils.neg_laplace_pairs(2)
    st = pb.Matrix.stencil_matrix((n, n), [s for s, _ in sh], [c for _, c in sh], torch.float32, ctx.dev)
    lo, hi = shard_bounds(batch, ctx.world, ctx.rank)
    g = torch.Generator(device='cpu')
    g.manual_seed(0)
    y = torch.randn((batch, n * n), dtype=torch.float32, generator=g)[lo:hi].to(ctx.dev)     # same data for any world size
    x0 = torch.zeros_like(y)
    nb = hi - lo
    res = [None]

    def solve():
        res[0] = solve_sharded('CG', st, y, x0, [1e-5] * nb, [0.0] * nb, np.full((1, nb), 5000), None, generic_rules=True)
    solve()
    ms = min(ctx.timed(solve, 3))
    its = res[0].iterations.to(torch.int64)
    tot, mx, mn = its.sum().reshape(1), its.max().reshape(1), its.min().reshape(1)
    conv = res[0].converged.all().to(torch.int64).reshape(1)
    if ctx.world > 1:
        dist.all_reduce(tot); dist.all_reduce(mx, op=dist.ReduceOp.MAX); dist.all_reduce(mn, op=dist.ReduceOp.MIN); dist.all_reduce(conv, op=dist.ReduceOp.MIN)
    out = {"workload": f"{batch} x 2D Poisson {n}^2 (shared matrix), CG fp32 rel_tol 1e-5, generic rules, batch sharded", "gpus": ctx.world, "ms_per_batched_solve": ms,
           "iterations_min_max": [int(mn), int(mx)], "converged": bool(conv), "system_iterations_per_s": float(tot) / (ms * 1e-3), "resident_kernel": bool(S.LAST_RUN['resident']),
           "survey_model_gbs_per_gpu": 13 * batch * n * n * 4 * float(mx) / ctx.world / (ms * 1e-3) / 1e9, "collectives": "one all-reduce(min) of the tolerance at set-up; none in the loop"}
    torch.cuda.empty_cache()
    return out


def time_cfg4(ctx, n, iters):
    """BASELINE config 4 (one GPU): ILU(0)-preconditioned CG fp32 on the n^3 Poisson system; triangular solves as structured
    wavefronts (csrc/star_trsv.cuh)."""
    import phiml_b200 as pb
    from phiml_b200 import _engine as E
    torch = ctx.torch
    offsets, coeffs = laplace_stencil()
    N = n ** 3
    t0 = time.perf_counter()
    st = pb.Matrix.stencil_matrix((n, n, n), offsets, coeffs, torch.float32, ctx.dev)
    csr = st.to_csr()
    m = pb.as_matrix(torch.sparse_csr_tensor(csr.rowptr, csr.col, csr.values, csr.shape))
    ilu = pb.IncompleteLU(m)
    torch.cuda.synchronize(ctx.dev)
    setup_s = time.perf_counter() - t0
    y = torch.randn((1, N), dtype=torch.float32, device=ctx.dev, generator=ctx.gen)
    zero = torch.zeros(1, dtype=torch.float32, device=ctx.dev)
    mi = torch.full((1,), 2 ** 30, dtype=torch.int32, device=ctx.dev)
    solver = pb.DeviceSolver(m.stencil, ilu, E.CG, 0, 1)
    solver.start(y, torch.zeros_like(y), zero, zero, mi)
    solver.advance(3)
    ms = min(ctx.timed(lambda: solver.advance(iters), 2))
    per_it = ms / iters
    t = torch.empty_like(y)
    apply_ms = min(ctx.timed(lambda: ilu.apply(y), 3))
    model = 23 * N * 4
    out = {"workload": f"3D Poisson {n}^3, ILU(0)-preconditioned CG fp32 ({ilu!r})", "n": n, "gpus": 1, "iterations_per_s": 1e3 / per_it, "ms_per_iteration": per_it,
           "star_wavefront": bool(ilu.star_wavefront), "ilu_apply_ms": apply_ms, "setup_s_assembly_plus_factorisation": setup_s,
           "survey_model_bytes": model, "survey_model_frac_of_peak": model / (per_it * 1e-3) / 1e9 / ctx.peak}
    del solver, ilu, m, csr, st, t
    torch.cuda.empty_cache()
    return out


def time_coded(ctx, n, kind):
    """SURVEY 8f N2 (second half): the n^3 7-point operator with ZERO_GRADIENT ('neumann') or PERIODIC boundaries (+ 0.01 I), as the CSR native
    matrix_from_function would hand over, recognised on the device as a pattern-coded stencil (csrc/coded.cuh).  Product and Jacobi-CG on the
    coded kernel and on the CSR kernel of the SAME native (operands rotated over > 4 x L2)."""
    import phiml_b200 as pb
    from phiml_b200 import _engine as E
    torch = ctx.torch
    dev, N = ctx.dev, n ** 3
    idx = torch.arange(N, device=dev, dtype=torch.int64)
    comps = (idx // (n * n), (idx // n) % n, idx % n)
    cols, vals = [], []
    deg = torch.zeros(N, device=dev, dtype=torch.float32)
    for comp, stride in zip(comps, (n * n, n, 1)):
        for sgn in (-1, 1):
            j = comp + sgn
            inside = torch.ones_like(j, dtype=torch.bool) if kind == 'periodic' else (j >= 0) & (j < n)
            jj = idx + ((j % n) - comp) * stride
            cols.append(torch.where(inside, jj, torch.full_like(jj, 2 ** 40)))
            vals.append(torch.where(inside, -torch.ones(N, device=dev), torch.zeros(N, device=dev)))
            deg += inside.float()
    cols.append(idx)
    vals.append(deg + 0.01)
    C, V = torch.stack(cols, 1), torch.stack(vals, 1)
    del cols, vals
    order = torch.argsort(C, dim=1)
    C, V = torch.gather(C, 1, order), torch.gather(V, 1, order)
    keep = C < 2 ** 40
    rowptr = torch.zeros(N + 1, dtype=torch.int32, device=dev)
    rowptr[1:] = torch.cumsum(keep.sum(1), 0).to(torch.int32)
    col, val = C[keep].to(torch.int32), V[keep]
    del C, V, keep, order
    t0 = time.perf_counter()
    m = pb.as_matrix(torch.sparse_csr_tensor(rowptr, col, val, (N, N)))
    torch.cuda.synchronize(dev)
    rec_s = time.perf_counter() - t0
    assert m.stencil is None and m.coded is not None, "the operator must be recognised as a pattern-coded stencil"
    nnz = int(val.numel())
    pairs = max(2, int(np.ceil(4 * 126e6 / (2 * N * 4))))
    xs = [torch.randn((1, N), dtype=torch.float32, device=dev, generator=ctx.gen) for _ in range(pairs)]
    ys = [torch.empty_like(xs[0]) for _ in range(pairs)]
    out = {"workload": f"3D 7-point operator {n}^3 with {'PERIODIC' if kind == 'periodic' else 'ZERO_GRADIENT'} boundaries (+0.01 I), fp32, CSR native recognised as a coded stencil",
           "n": n, "nnz": nnz, "patterns": m.coded.patterns, "offsets": len(m.coded.offsets), "recognition_s": rec_s}
    lib, sp = E.lib(), E.stream_ptr(dev)
    for name, mat, nbytes in (("coded", m.coded, N + 2 * N * 4), ("csr", m, 8 * nnz + 4 * (N + 1) + 2 * N * 4)):
        def call(i):
            E.check(lib.phb200_spmv(mat.handle, E.ptr(xs[i % pairs]), E.ptr(ys[i % pairs]), 1, N, N, sp))
        for i in range(pairs + 1):
            call(i)
        ms = min(ctx.timed(lambda: [call(i) for i in range(40)], 2)) / 40
        y = xs[0]
        zero = torch.zeros(1, dtype=torch.float32, device=dev)
        mi = torch.full((1,), 2 ** 30, dtype=torch.int32, device=dev)
        solver = pb.DeviceSolver(mat, pb.JacobiPreconditioner(m), E.CG, 0, 1)
        solver.start(y, torch.zeros_like(y), zero, zero, mi)
        solver.advance(10)
        it_ms = min(ctx.timed(lambda: solver.advance(60), 2)) / 60
        out[name] = {"spmv_ms": ms, "algorithmic_bytes": nbytes, "gbs": nbytes / ms / 1e6, "frac_of_peak": nbytes / ms / 1e6 / ctx.peak, "jacobi_cg_iterations_per_s": 1e3 / it_ms}
        del solver
    ya, yb = torch.empty_like(xs[0]), torch.empty_like(xs[0])
    E.check(lib.phb200_spmv(m.coded.handle, E.ptr(xs[0]), E.ptr(ya), 1, N, N, sp))
    E.check(lib.phb200_spmv(m.handle, E.ptr(xs[0]), E.ptr(yb), 1, N, N, sp))
    out["products_bit_identical"] = bool(torch.equal(ya, yb))
    del m, xs, ys, ya, yb, rowptr, col, val
    pb.clear_cache()
    torch.cuda.empty_cache()
    return out


def guarded(name, fn, sink, ctx):
    """Extra blocks must never take the headline down: an error is recorded in the block instead (all ranks stay in step)."""
    try:
        sink[name] = fn()
    except Exception as exc:
        sink[name] = {"error": repr(exc)[:300]}
        if ctx.world > 1:
            raise


def run_ours(args):
    ctx = Ctx(args)
    torch, dist = ctx.torch, ctx.dist
    import phiml_b200 as pb
    from phiml_b200 import _engine as E
    world, rank, dev = ctx.world, ctx.rank, ctx.dev
    n, N, w = args.n, args.n ** 3, 4
    peak, peak_src = ctx.peak, ctx.peak_src
    offsets, coeffs = laplace_stencil()

    if world == 1 and args.format == 'csr':
        return run_csr(args, ctx)

    # ---- headline: K Jacobi-CG iterations of ONE n^3 system on `world` GPUs
    with ClockSampler(ctx.local_rank) as clocks:
        head, ms, launches, prof = time_jacobi_cg(ctx, n, args.warmup, args.steps, want_profile=True)
    value = args.steps / (ms * 1e-3)

    # ---- per-kernel durations (CUDA events on the launch stream around each kernel of 60 extra iterations)
    names = ["k_star_cg_tma: d=M^-1 r+beta d, x+=alpha d, q=A d, d.q" + (" + in-kernel all-reduce over NVLink" if world > 1 else ""),
             "k_phase<CgResid>: r-=alpha q, M^-1 r, r.s, beta, convergence test" + (" + peer halo stores + in-kernel all-reduce" if world > 1 else "")]
    model_bytes = [6 * N * w // world, 3 * N * w // world]
    kernels = []
    traffic = ncu_traffic('stencil', n) if world == 1 else {}
    tkeys = ['k_star', 'k_phase']
    for k, (tot_ms, cnt) in enumerate((prof or [])[:2]):
        avg = tot_ms / max(cnt, 1)
        tr = traffic.get(tkeys[k])
        kernels.append({"kernel": names[k], "avg_ms": avg, "bytes": model_bytes[k], "gbs": model_bytes[k] / (avg * 1e-3) / 1e9 if avg > 0 else None,
                        "traffic": (tr['dram_read'] + tr['dram_write']) if tr else None})
    roofline = None
    if kernels:
        dom = max(range(len(kernels)), key=lambda k: kernels[k]['avg_ms'])
        kd = kernels[dom]
        roofline = {"bound": "hbm", "kernel": kd['kernel'], "achieved": kd['gbs'], "peak": peak, "unit": "GB/s", "frac": kd['gbs'] / peak, "traffic": kd['traffic'],
                    "traffic_source": "ncu --set full capture (cold L2), profiles/ncu_traffic.json <- profiles/r02_ncu_full_cg.txt" if kd['traffic'] else None, "peak_source": peak_src,
                    "per_gpu": world > 1, "share_of_step": kd['avg_ms'] / sum(k_['avg_ms'] for k_ in kernels), "kernels": kernels}
    iteration = {"bytes_moved_by_design": 9 * N * w, "achieved_gbs_per_gpu": head["hbm_gbs_per_gpu"], "frac_of_peak_per_gpu": head["hbm_frac_of_peak_per_gpu"],
                 "survey_model_bytes": 13 * N * w, "survey_model_frac_of_peak_per_gpu": head["survey_model_frac_of_peak_per_gpu"]}

    # ---- stand-alone products (BASELINE metric: "SpMV HBM GB/s (% of peak)"), one GPU, operands rotated over > 4 x L2
    spmv = None
    stencil = csr = native = None
    if world == 1:
        stencil = pb.Matrix.stencil_matrix((n, n, n), offsets, coeffs, torch.float32, dev)
        if stencil.nnz() < 2 ** 31:
            csr = stencil.to_csr()              # the native the torch backend would hold (device-side assembly, bit-exact)
            native = torch.sparse_csr_tensor(csr.rowptr, csr.col, csr.values, csr.shape)
        spmv = time_products(ctx, stencil, csr, N)

    # ---- end to end through the Backend-level call with HOST buffers
    e2e = None
    if not args.no_e2e:
        e2e = e2e_single(ctx, native, n) if world == 1 else e2e_slab(ctx, n)

    # ---- the same solve through PhiML's OWN front end (math.solve_linear with the plug-in installed; baseline/_ref), cold and warm
    if e2e is not None and world == 1 and rank == 0:
        e2e["front_end"] = front_end_block(n)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, its, dt, kind, sample = cpu_baseline(n, 3, 20)
        cpu = {"value": v, "unit": UNIT, "cores": 1, "kind": kind, "sample": sample}

    configs = {}
    out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
           "data": "synthetic", "config": workload_config(n, world), "clocks": clocks.summary(), "e2e": e2e,
           "gpu_launches": launches, "roofline": roofline, "iteration": iteration, "headline_detail": head, "spmv": spmv, "configs": configs,
           "cpu_baseline": cpu, "impl": "phb200",
           "impl_detail": "stencil-specialised TMA kernels (matrix recognised from the CSR native)" + ("" if world == 1 else "; SlabCG: peer-memory halo stores + in-kernel all-reduce, no NCCL in the loop")}
    emitted = threading.Lock()

    def emit():
        if emitted.acquire(blocking=False) and rank == 0:
            print(json.dumps(out), flush=True)

    # ---- the other BASELINE configs.  They must never cost the headline: a watchdog prints the line with what has been collected
    #      so far and ends the process if a block overruns its budget (or a rank is stuck in a collective).
    if args.extras:
        def overrun():
            configs["watchdog"] = "extra blocks exceeded their time budget; line printed without the missing ones"
            emit()
            os._exit(0)
        dog = threading.Timer(args.extras_budget + (0 if rank == 0 else 5), overrun)
        dog.daemon = True
        dog.start()
        try:
            if world == 1:
                del stencil, csr, native
                pb.clear_cache()
                torch.cuda.empty_cache()
            big = args.big
            guarded("slab_1024", lambda: time_jacobi_cg(ctx, big, 5, 30)[0], configs, ctx)
            guarded("sharded_cfg3", lambda: time_cfg3(ctx), configs, ctx)
            guarded("slab_cfg5", lambda: time_cfg5(ctx, big, 6 if world == 1 else 8), configs, ctx)
            if world == 1:
                guarded("cfg4_ilu_256", lambda: time_cfg4(ctx, 256, 30), configs, ctx)
                guarded("cfg4_ilu_512", lambda: time_cfg4(ctx, 512, 20), configs, ctx)
                guarded("coded_neumann_256", lambda: time_coded(ctx, 256, 'neumann'), configs, ctx)
                guarded("coded_periodic_256", lambda: time_coded(ctx, 256, 'periodic'), configs, ctx)
        except Exception as exc:              # multi-rank: a failed block may leave the ranks out of step — stop here, keep the headline
            configs["aborted"] = repr(exc)[:300]
            emit()
            os._exit(0)
        dog.cancel()
    emit()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def time_products(ctx, stencil, csr_matrix, N):
    import phiml_b200 as pb
    from phiml_b200 import _engine as E
    torch = ctx.torch
    dev, w, peak = ctx.dev, 4, ctx.peak
    pairs = max(2, int(np.ceil(4 * 126e6 / (2 * N * w))))
    xs = [torch.randn((1, N), dtype=torch.float32, device=dev, generator=ctx.gen) for _ in range(pairs)]
    ys = [torch.empty_like(xs[0]) for _ in range(pairs)]
    lib, sp = E.lib(), E.stream_ptr(dev)
    ptrs = [(E.ptr(a), E.ptr(b)) for a, b in zip(xs, ys)]

    def time_matvec(handle, reps=50):
        # phb200_spmv called directly (the Python wrapper costs more host time than the 25 us stencil kernel runs)
        def call(i):
            px, py = ptrs[i % pairs]
            E.check(lib.phb200_spmv(handle, px, py, 1, N, N, sp))
        for i in range(pairs + 1):
            call(i)
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(reps):
            call(i)
        e1.record()
        torch.cuda.synchronize(dev)
        return e0.elapsed_time(e1) / reps
    spmv = {"l2": "operands rotated over %d (x, y) pairs = %.0f MB, > 4 x L2" % (pairs, pairs * 2 * N * w / 1e6)}
    ms_st = time_matvec(stencil.handle)
    b_st = 2 * N * w
    spmv["stencil"] = {"ms": ms_st, "algorithmic_bytes": b_st, "gbs": b_st / ms_st / 1e6, "frac_of_peak": b_st / ms_st / 1e6 / peak}
    if csr_matrix is not None:
        nnz = int(csr_matrix.values.numel())
        ms_csr = time_matvec(csr_matrix.handle)
        b_csr = 8 * nnz + 4 * (N + 1) + 2 * N * w
        spmv["csr"] = {"ms": ms_csr, "algorithmic_bytes": b_csr, "gbs": b_csr / ms_csr / 1e6, "frac_of_peak": b_csr / ms_csr / 1e6 / peak}
    return spmv


def e2e_single(ctx, native, n):
    """B200Backend.linear_solve with pinned HOST y / x0 and x copied back, solve to rel_tol 1e-5.  ``warm``: the matrix native is
    resident on the device (what the plug-in's residency cache gives repeated solve_linear calls); ``cold``: the CSR native
    itself (0.94 GB at 256^3) is uploaded from pinned host memory inside the timed region, as a first call would."""
    import phiml_b200 as pb
    torch = ctx.torch
    dev, N, w = ctx.dev, n ** 3, 4
    if native is None:
        return None
    y_host = torch.empty((1, N), dtype=torch.float32).pin_memory()
    y_host.copy_(torch.as_tensor(np.random.default_rng(0).standard_normal((1, N)).astype(np.float32)))
    x0_host = torch.zeros((1, N), dtype=torch.float32).pin_memory()
    x_host = torch.empty((1, N), dtype=torch.float32).pin_memory()
    backend = pb.B200Backend(dev)
    parts = {"h2d": 0.0, "setup": 0.0, "solve": 0.0, "d2h": 0.0}

    def tick():
        torch.cuda.synchronize(dev)
        return time.perf_counter()

    def one(nat, upload=None):
        pb.clear_cache()
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        if upload is not None:
            crow, col, val = (t.to(dev, non_blocking=True) for t in upload)
            nat = torch.sparse_csr_tensor(crow, col, val, (N, N))
        yd = y_host.to(dev, non_blocking=True)
        x0d = x0_host.to(dev, non_blocking=True)
        t1 = tick()
        m = pb.as_matrix(nat)
        p = backend.compute_preconditioner('jacobi', m)
        t2 = tick()
        res = backend.linear_solve('CG', m, yd, x0d, [1e-5], [0.0], np.array([[5000]]), p, None)
        t3 = tick()
        x_host.copy_(res.x, non_blocking=True)
        its = int(res.iterations.cpu()[0])
        conv = bool(res.converged.cpu()[0])
        torch.cuda.synchronize(dev)
        t4 = time.perf_counter()
        assert conv, "e2e solve did not converge"
        return its, t4 - t0, (t1 - t0, t2 - t1, t3 - t2, t4 - t3)

    one(native)
    its_sum, s_sum = 0, 0.0
    for rep in range(2):
        its, dt, pt = one(native)
        its_sum += its
        s_sum += dt
        for k_, v_ in zip(parts, pt):
            parts[k_] += v_ / 2
    e2e = {"value": its_sum / s_sum, "unit": UNIT, "h2d_bytes_per_step": 2 * N * w, "d2h_bytes_per_step": N * w + 6,
           "iterations_per_solve": its_sum // 2, "seconds_per_solve": s_sum / 2, "seconds_breakdown": parts,
           "call": "B200Backend.linear_solve('CG', csr_native, y, x0, rtol, atol, max_iter, jacobi) incl. stencil recognition + Jacobi set-up; warm = matrix native resident on the device"}
    # cold: the CSR native travels too
    host_csr = tuple(t.cpu().pin_memory() for t in (native.crow_indices(), native.col_indices(), native.values()))
    its, dt, pt = one(None, upload=host_csr)
    e2e["cold"] = {"value": its / dt, "unit": UNIT, "seconds_per_solve": dt, "h2d_bytes": int(sum(t.numel() * t.element_size() for t in host_csr)) + 2 * N * w,
                   "seconds_breakdown": dict(zip(parts, pt)), "note": "first call: CSR native uploaded from pinned host memory inside the timed region"}
    return e2e


def front_end_block(n):
    """math.solve_linear(jit_compile_linear(-laplace), y, Solve('CG', 1e-5, preconditioner='jacobi')) on the unmodified reference's front end
    with the phb200 plug-in (scripts/phiml_front_end.py in a child process: trace -> device assembly -> residency -> device loop; host y in,
    host x out).  cold = first call (trace + assembly + preconditioner), warm = repeated calls."""
    script = os.path.join(ROOT, 'scripts', 'phiml_front_end.py')
    if not os.path.isdir(os.path.join(ROOT, 'baseline', '_ref', 'phiml')):
        return {"unavailable": "no copy of the reference (baseline/_ref)"}
    try:
        r = subprocess.run([sys.executable, script, '--n', str(n), '--pre', 'jacobi', '--reps', '4'], capture_output=True, text=True, timeout=240)
        line = [l for l in r.stdout.splitlines() if l.startswith('{')][-1]
        d = json.loads(line)
        cold, warm = d["calls"][0], d["calls"][1:]
        ws = float(np.median([c["seconds"] for c in warm]))
        return {"call": "phiml.math.solve_linear through the plug-in (host y in, host x out)", "iterations": cold["iterations"], "method": cold["method"],
                "cold_seconds": cold["seconds"], "warm_seconds": ws, "warm_value": cold["iterations"] / ws, "unit": UNIT,
                "warm_backend_call_seconds": float(np.median([c["backend_call_seconds"] for c in warm])), "true_residual": d.get("true_residual"),
                "device_assemblies": d.get("assembly", {}).get("device_assemblies"), "native_cache": d.get("native_cache")}
    except Exception as exc:
        return {"error": repr(exc)[:300]}


def e2e_slab(ctx, n):
    """N GPUs: SlabCG.solve with every rank's slab of y / x0 coming from pinned host memory and its slab of x copied back."""
    from phiml_b200.multi_gpu import SlabCG
    torch = ctx.torch
    offsets, coeffs = laplace_stencil()
    cg = SlabCG((n, n, n), offsets, coeffs, torch.float32, jacobi=True, device=ctx.dev)
    plane, w = n * n, 4
    y_full = np.random.default_rng(0).standard_normal((1, n ** 3)).astype(np.float32)
    y_host = torch.as_tensor(y_full[:, cg.x_lo * plane:cg.x_hi * plane]).contiguous().pin_memory()
    x0_host = torch.zeros_like(y_host).pin_memory()
    x_host = torch.empty_like(y_host).pin_memory()
    its_sum, s_sum = 0, 0.0
    for rep in range(3):
        ctx.fence()
        t0 = time.perf_counter()
        yd = y_host.to(ctx.dev, non_blocking=True)
        x0d = x0_host.to(ctx.dev, non_blocking=True)
        x, r, its, fev, conv, div = cg.solve(yd, x0d, [1e-5], [0.0], 5000)
        x_host.copy_(x, non_blocking=True)
        k = int(its.cpu()[0])
        ok = bool(conv.cpu()[0])
        torch.cuda.synchronize(ctx.dev)
        dt = ctx.max_over_ranks(time.perf_counter() - t0)
        assert ok, "e2e slab solve did not converge"
        if rep > 0:
            its_sum += k
            s_sum += dt
    return {"value": its_sum / s_sum, "unit": UNIT, "h2d_bytes_per_step": 2 * n ** 3 * w, "d2h_bytes_per_step": n ** 3 * w + 6 * ctx.world,
            "iterations_per_solve": its_sum // 2, "seconds_per_solve": s_sum / 2,
            "call": "SlabCG.solve(y_slab, x0_slab, rtol 1e-5) on every rank: pinned host slabs in, x slab out; set-up (extended arrays, peer-memory mappings via CUDA IPC) inside the timed region"}


def run_csr(args, ctx):
    """--format csr (one GPU): the same system through the general CSR kernels (no stencil recognition)."""
    import phiml_b200 as pb
    from phiml_b200 import _engine as E
    torch = ctx.torch
    dev, n, N, w = ctx.dev, args.n, args.n ** 3, 4
    offsets, coeffs = laplace_stencil()
    stencil = pb.Matrix.stencil_matrix((n, n, n), offsets, coeffs, torch.float32, dev)
    csr = stencil.to_csr()
    native = torch.sparse_csr_tensor(csr.rowptr, csr.col, csr.values, csr.shape)
    matrix = pb.as_matrix(native, detect_stencil=False)
    pre = pb.JacobiPreconditioner(matrix)
    y = torch.randn((1, N), dtype=torch.float32, device=dev, generator=ctx.gen)
    zero = torch.zeros(1, dtype=torch.float32, device=dev)
    mi = torch.full((1,), 2 ** 30, dtype=torch.int32, device=dev)
    solver = pb.DeviceSolver(matrix, pre, E.CG, 0, 1)
    solver.start(y, torch.zeros_like(y), zero, zero, mi)
    solver.advance(args.warmup)
    launches0 = pb.launch_count()
    with ClockSampler(ctx.local_rank) as clocks:
        ms_all = ctx.timed(lambda: solver.advance(args.steps), 5)
    launches = (pb.launch_count() - launches0) // 5
    ms = float(np.median(ms_all))
    solver.profile(True)
    solver.advance(60)
    prof = solver.profile_read()
    solver.profile(False)
    nnz = int(csr.values.numel())
    names = ["k_spmv_csr: q=A d, d.q, alpha", "k_phase<CgUpdate>: x,r update, M^-1 r, r.s, beta, convergence test", "k_phase<CgDirection>: d = M^-1 r + beta d"]
    model_bytes = [8 * nnz + 4 * (N + 1) + 2 * N * w, 7 * N * w, 4 * N * w]
    kernels = []
    for k, (tot_ms, cnt) in enumerate(prof[:3]):
        avg = tot_ms / max(cnt, 1)
        kernels.append({"kernel": names[k], "avg_ms": avg, "bytes": model_bytes[k], "gbs": model_bytes[k] / (avg * 1e-3) / 1e9})
    kd = max(kernels, key=lambda k_: k_['avg_ms'])
    roofline = {"bound": "hbm", "kernel": kd['kernel'], "achieved": kd['gbs'], "peak": ctx.peak, "unit": "GB/s", "frac": kd['gbs'] / ctx.peak, "traffic": None,
                "peak_source": ctx.peak_src, "share_of_step": kd['avg_ms'] / sum(k_['avg_ms'] for k_ in kernels), "kernels": kernels}
    its_s = args.steps / (ms * 1e-3)
    survey = 8 * nnz + 4 * (N + 1) + 13 * N * w
    out = {"metric": METRIC, "value": its_s, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
           "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(n, 1), "clocks": clocks.summary(), "e2e": None,
           "gpu_launches": launches, "roofline": roofline,
           "iteration": {"bytes_moved_by_design": sum(model_bytes), "frac_of_peak": sum(model_bytes) * its_s / 1e9 / ctx.peak, "survey_model_bytes": survey, "survey_model_frac_of_peak": survey * its_s / 1e9 / ctx.peak},
           "spmv": time_products(ctx, stencil, matrix, N), "impl": "phb200", "impl_detail": "general CSR int32 kernels (stencil recognition disabled)"}
    print(json.dumps(out))


# ----------------------------------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU implementation of the path on the host cores
# ----------------------------------------------------------------------------------------------------------------------

def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    steps = min(args.steps, 30)
    warm = max(1, min(args.warmup, 3))
    t0 = time.perf_counter()
    v, its, dt, kind, sample = cpu_baseline(args.n, warm, steps)
    out = {"metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": its, "warmup": warm, "ms_per_step": dt / its * 1e3,
           "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": workload_config(args.n, args.gpus), "impl": "reference",
           "impl_detail": "SciPy CSR int32 (csr_matvec) + NumPy ufuncs on one host core; the reference has no multi-GPU (or GPU-resident sparse solver) path, so the same CPU run stands for every N",
           "cpu_baseline": {"value": v, "unit": UNIT, "cores": 1, "kind": kind, "sample": sample + f"; whole run {time.perf_counter() - t0:.0f} s"},
           "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(out))


if __name__ == '__main__':
    a = parse()
    if a.impl == 'reference':
        run_reference(a)
    else:
        run_ours(a)
