#!/usr/bin/env python
"""
bench.py — BASELINE.json metric: CG iterations/s (and HBM GB/s as a fraction of the measured peak) of the
Jacobi-preconditioned CG solve of the 3-D Poisson system 256^3 (7-point Laplacian, ZERO padding) in fp32.

A *step* is one CG iteration (one pass of the hot loop: SpMV + fused dots, x/r update + preconditioner + dot,
direction update, on-device convergence test) over one system.  With N ranks every rank owns an independent system
(weak scaling, no data-path collective).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--format stencil|csr] [--n 256]

Under torchrun every rank runs this file; rank 0 prints ONE JSON line.
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
    p.add_argument('--steps', type=int, default=400)
    p.add_argument('--warmup', type=int, default=20)
    p.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    p.add_argument('--format', default='stencil', choices=['stencil', 'csr'])
    p.add_argument('--n', type=int, default=256)
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


def workload_config(n, fmt, n_gpus):
    return {"workload": f"3D Poisson {n}^3 7-point Laplacian (ZERO padding), Jacobi-preconditioned CG fp32, rel_tol 1e-5",
            "matrix": "stencil-specialised kernels (auto-detected from the CSR native)" if fmt == 'stencil' else "general CSR int32",
            "rows": n ** 3, "nnz": 7 * n ** 3 - 6 * n ** 2, "systems_per_gpu": 1, "parallelism": f"independent systems x{n_gpus}",
            "l2": "working set (5 vectors x %.0f MB) exceeds the 126 MB L2; no flush needed" % (n ** 3 * 4 / 1e6)}


# ----------------------------------------------------------------------------------------------------------------------
# CPU baseline: the oracle (NumPy/SciPy restatement of the reference's generic cg + SciPy CSR matvec), 1 core
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


def cpu_iterations_per_second(n, warmup, steps):
    """Runs the oracle's Jacobi-CG for warmup+steps iterations on the same system; returns (it/s, description)."""
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
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.FIELDS}', '--format=csv,noheader,nounits', '-lms', '100', '-i', str(self.index)],
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

def run_ours(args):
    import torch
    import torch.distributed as dist
    import phiml_b200 as pb
    from phiml_b200 import _engine as E
    from phiml_b200 import stencils

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    assert world == args.gpus or world == 1, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)

    n, N = args.n, args.n ** 3
    w = 4
    shifts = stencils.neg_laplace_pairs(3)
    stencil = pb.Matrix.stencil_matrix((n, n, n), [s for s, _ in shifts], [c for _, c in shifts], torch.float32, dev)
    too_big_for_csr = stencil.nnz() >= 2 ** 31      # int32 CSR (the reference's native) cannot hold it: stencil handle only
    if too_big_for_csr:
        assert args.format == 'stencil'
        args.no_e2e = True
        csr, native, matrix, solver_matrix = None, None, stencil, stencil
    else:
        csr = stencil.to_csr()              # the native the torch backend would hold (device-side assembly, bit-exact)
        native = torch.sparse_csr_tensor(csr.rowptr, csr.col, csr.values, csr.shape)
    if too_big_for_csr:
        pass
    elif args.format == 'stencil':
        matrix = pb.as_matrix(native)       # recognised as a stencil on the device
        assert matrix.stencil is not None, "stencil recognition failed"
        solver_matrix = matrix.stencil
    else:
        matrix = pb.as_matrix(native, detect_stencil=False)
        solver_matrix = matrix
    pre = pb.JacobiPreconditioner(matrix)
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234 + rank)
    y = torch.randn((1, N), dtype=torch.float32, device=dev, generator=gen)
    x0 = torch.zeros_like(y)
    zero = torch.zeros(1, dtype=torch.float32, device=dev)
    mi = torch.full((1,), 2 ** 30, dtype=torch.int32, device=dev)

    # ---- device-resident throughput: K iterations with the convergence test executed but never satisfied (rtol=atol=0)
    solver = pb.DeviceSolver(solver_matrix, pre, E.CG, 0, 1)
    solver.start(y, x0, zero, zero, mi)
    solver.advance(args.warmup)
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = pb.launch_count()
    with ClockSampler(local_rank) as clocks:
        ev0.record()
        solver.advance(args.steps)
        ev1.record()
        torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    launches = pb.launch_count() - launches0
    ms = ev0.elapsed_time(ev1)
    its_done = int(solver.result()[0][0])
    assert its_done == args.warmup + args.steps, f"device executed {its_done} iterations, expected {args.warmup + args.steps}"
    assert solver.any_continue()
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t[0])
    value = world * args.steps / (ms * 1e-3)

    # ---- per-kernel durations (CUDA events on the launch stream) for the roofline line
    solver.profile(True)
    solver.advance(60)
    prof = solver.profile_read()
    solver.profile(False)
    if args.format == 'stencil':
        # fused two-kernel iteration on the recognised stencil: bytes the kernels must move
        names = ["k_star: d=M^-1 r+beta d, x+=alpha d, q=A d, d.q (fused direction + SpMV + dot)",
                 "k_phase<CgResid>: r-=alpha q, M^-1 r, r.s, beta, convergence test"]
        model_bytes = [6 * N * w, 3 * N * w]
    else:
        names = ["k_spmv_csr_stream: q=A d, d.q, alpha", "k_phase<CgUpdate>: x,r update, M^-1 r, r.s, beta, convergence test", "k_phase<CgDirection>: d = M^-1 r + beta d"]
        model_bytes = [8 * csr.values.numel() + 4 * (N + 1) + 2 * N * w, 7 * N * w, 4 * N * w]
    peak, peak_src = measured_peak()
    kernels = []
    traffic = ncu_traffic(args.format, n)
    tkeys = ['k_star', 'k_phase'] if args.format == 'stencil' else []
    for k, (tot_ms, cnt) in enumerate(prof[:len(names)]):
        avg = tot_ms / max(cnt, 1)
        tr = traffic.get(tkeys[k]) if k < len(tkeys) else None
        kernels.append({"kernel": names[k], "avg_ms": avg, "bytes": model_bytes[k], "gbs": model_bytes[k] / (avg * 1e-3) / 1e9 if avg > 0 else None,
                        "traffic": (tr['dram_read'] + tr['dram_write']) if tr else None})
    dom = max(range(len(kernels)), key=lambda k: kernels[k]['avg_ms']) if kernels else None
    roofline = None
    if dom is not None:
        kd = kernels[dom]
        roofline = {"bound": "hbm", "kernel": kd['kernel'], "achieved": kd['gbs'], "peak": peak, "unit": "GB/s", "frac": kd['gbs'] / peak,
                    "traffic": kd['traffic'], "traffic_source": "ncu --set full capture (cold L2), profiles/ncu_traffic.json" if kd['traffic'] else None, "peak_source": peak_src, "share_of_step": kd['avg_ms'] / sum(k_['avg_ms'] for k_ in kernels),
                    "kernels": kernels}
    its_per_s = args.steps / (ms * 1e-3)
    survey_bytes = 13 * N * w if args.format == 'stencil' else 8 * csr.values.numel() + 4 * (N + 1) + 13 * N * w   # SURVEY.md §8(d) canonical model
    iter_bytes = sum(model_bytes)
    iteration = {"bytes_moved_by_design": iter_bytes, "achieved_gbs": iter_bytes * its_per_s / 1e9, "frac_of_peak": iter_bytes * its_per_s / 1e9 / peak,
                 "survey_model_bytes": survey_bytes, "survey_model_gbs": survey_bytes * its_per_s / 1e9, "survey_model_frac_of_peak": survey_bytes * its_per_s / 1e9 / peak}

    # ---- stand-alone products (BASELINE metric: "SpMV HBM GB/s (% of peak)"): y = A x through phb200_spmv, device-resident,
    #      operands rotated over a working set of > 4 x L2
    spmv = None
    if rank == 0:
        def time_matvec(mat, use_stencil, reps=50):
            # phb200_spmv called directly (the Python wrapper costs more host time than the 25 us stencil kernel runs)
            m_ = mat.stencil if (use_stencil and mat.stencil is not None) else mat
            # rotate over enough (x, y) pairs that consecutive calls cannot be served from the 126 MB L2
            pairs = max(2, int(np.ceil(4 * 126e6 / (2 * N * w))))
            xs = [torch.randn((1, N), dtype=torch.float32, device=dev, generator=gen) for _ in range(pairs)]
            ys = [torch.empty_like(xs[0]) for _ in range(pairs)]
            lib, sp = E.lib(), E.stream_ptr(dev)
            ptrs = [(E.ptr(a), E.ptr(b)) for a, b in zip(xs, ys)]

            def call(i):
                px, py = ptrs[i % pairs]
                E.check(lib.phb200_spmv(m_.handle, px, py, 1, N, N, sp))
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
        spmv = {"l2": "operands rotated over %d (x, y) pairs = %.0f MB, > 4 x L2" % (max(2, int(np.ceil(4 * 126e6 / (2 * N * w)))), max(2, int(np.ceil(4 * 126e6 / (2 * N * w)))) * 2 * N * w / 1e6)}
        ms_st = time_matvec(stencil, True)
        b_st = 2 * N * w
        spmv["stencil"] = {"ms": ms_st, "algorithmic_bytes": b_st, "gbs": b_st / ms_st / 1e6, "frac_of_peak": b_st / ms_st / 1e6 / peak}
        if not too_big_for_csr:
            nnz = int(csr.values.numel())
            ms_csr = time_matvec(matrix, False)
            b_csr = 8 * nnz + 4 * (N + 1) + 2 * N * w
            spmv["csr"] = {"ms": ms_csr, "algorithmic_bytes": b_csr, "gbs": b_csr / ms_csr / 1e6, "frac_of_peak": b_csr / ms_csr / 1e6 / peak}

    # ---- end to end through the Backend-level call: pinned host y/x0 -> device, solve to rel_tol 1e-5, x -> host
    e2e = None
    if not args.no_e2e:
        y_host = torch.empty((1, N), dtype=torch.float32).pin_memory()
        y_host.copy_(torch.as_tensor(np.random.default_rng(0).standard_normal((1, N)).astype(np.float32)))
        x0_host = torch.zeros((1, N), dtype=torch.float32).pin_memory()
        x_host = torch.empty((1, N), dtype=torch.float32).pin_memory()
        backend = pb.B200Backend(dev)
        e2e_its, e2e_s = 0, 0.0
        parts = {"h2d": 0.0, "setup": 0.0, "solve": 0.0, "d2h": 0.0}

        def tick():
            torch.cuda.synchronize(dev)
            return time.perf_counter()

        for rep in range(3):
            pb.clear_cache()
            torch.cuda.synchronize(dev)
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            yd = y_host.to(dev, non_blocking=True)
            x0d = x0_host.to(dev, non_blocking=True)
            t1 = tick()
            m = pb.as_matrix(native, detect_stencil=(args.format == 'stencil'))
            p = backend.compute_preconditioner('jacobi', m)
            t2 = tick()
            res = backend.linear_solve('CG', m, yd, x0d, [1e-5], [0.0], np.array([[5000]]), p, None)
            t3 = tick()
            x_host.copy_(res.x, non_blocking=True)
            its = int(res.iterations.cpu()[0])
            conv = bool(res.converged.cpu()[0])
            torch.cuda.synchronize(dev)
            t4 = time.perf_counter()
            dt = t4 - t0
            assert conv, "e2e solve did not converge"
            if rep > 0:
                e2e_its += its
                e2e_s += dt
                for k_, v_ in zip(parts, (t1 - t0, t2 - t1, t3 - t2, t4 - t3)):
                    parts[k_] += v_ / 2
        if world > 1:
            t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_s = float(t[0])
        e2e = {"value": world * e2e_its / e2e_s, "unit": UNIT, "h2d_bytes_per_step": 2 * N * w, "d2h_bytes_per_step": N * w + 6,
               "iterations_per_solve": e2e_its // 2, "seconds_per_solve": e2e_s / 2, "seconds_breakdown": parts,
               "call": "B200Backend.linear_solve('CG', csr_native, y, x0, rtol, atol, max_iter, jacobi) incl. stencil recognition + Jacobi set-up; matrix native resident on the device"}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, its, dt = cpu_iterations_per_second(n, 3, 30)
        cpu = {"value": v, "unit": UNIT, "cores": 1, "kind": "port",
               "sample": f"{its} Jacobi-CG iterations of the same {n}^3 fp32 system with the oracle (NumPy ufuncs + SciPy csr_matvec, single-threaded like the reference's NumPy backend), {dt:.1f} s"}

    if rank == 0:
        out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
               "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
               "data": "synthetic", "config": workload_config(n, args.format, world), "clocks": clocks.summary(), "e2e": e2e,
               "gpu_launches": launches, "roofline": roofline, "iteration": iteration, "spmv": spmv, "cpu_baseline": cpu, "impl": "phb200"}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


# ----------------------------------------------------------------------------------------------------------------------
# reference arm: the reference's CPU algorithm (oracle port: NumPy/SciPy, like its NumPy backend) on the host cores
# ----------------------------------------------------------------------------------------------------------------------

def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    steps = min(args.steps, 40)
    warm = max(1, min(args.warmup, 5))
    t0 = time.perf_counter()
    v, its, dt = cpu_iterations_per_second(args.n, warm, steps)
    out = {"metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": its, "warmup": warm, "ms_per_step": dt / its * 1e3,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": workload_config(args.n, 'csr', args.gpus), "impl": "reference",
           "cpu_baseline": {"value": v, "unit": UNIT, "cores": 1, "kind": "port",
                            "sample": f"{its} timed Jacobi-CG iterations ({warm} warm-up) of the {args.n}^3 fp32 system, oracle port of phiml/backend/_linalg.py:52-90 on SciPy CSR; "
                                      f"NumPy ufuncs and SciPy csr_matvec are single-threaded, so 1 core is all the path can use; whole run {time.perf_counter() - t0:.0f} s"},
           "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(out))


if __name__ == '__main__':
    a = parse()
    if a.impl == 'reference':
        run_reference(a)
    else:
        run_ours(a)
