"""
BASELINE config 5 driver: nonsymmetric 7-point advection-diffusion system, BiCGstab(2) fp32 (or any method), ONE system
row-partitioned in x-slabs over the ranks (phiml_b200.multi_gpu.SlabSolver: NCCL halo exchange before every product,
all-reduce for every dot product).  Run under torchrun, one rank per GPU:

    torchrun --nproc-per-node N scripts/run_slab_any.py --size 256 --check        # parity against the single-GPU device loop on rank 0
    torchrun --nproc-per-node N scripts/run_slab_any.py --size 1024 --iters 20    # strong-scaling throughput (fixed iterations)

Prints one JSON line from rank 0.
"""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch, torch.distributed as dist
import phiml_b200 as pb
from phiml_b200.multi_gpu import SlabSolver, slab_bounds
from phiml_b200 import stencils
ADVDIFF = stencils.ADVDIFF

p = argparse.ArgumentParser()
p.add_argument('--size', dest='n', type=int, default=256)
p.add_argument('--iters', type=int, default=6)
p.add_argument('--warmup', type=int, default=3)
p.add_argument('--method', default='biCG-stab(2)')
p.add_argument('--op', default='advdiff', choices=['advdiff', 'poisson'])
p.add_argument('--jacobi', action='store_true')
p.add_argument('--check', action='store_true')
a = p.parse_args()
world = int(os.environ.get('WORLD_SIZE', '1')); rank = int(os.environ.get('RANK', '0')); lr = int(os.environ.get('LOCAL_RANK', '0'))
torch.cuda.set_device(lr); dev = torch.device('cuda', lr)
if world > 1:
    dist.init_process_group('nccl', device_id=dev)
n = a.n; grid = (n, n, n); plane = n * n
shifts = stencils.advection_diffusion_pairs(3, ADVDIFF['velocity'], ADVDIFF['c'], ADVDIFF['nu'], np.float32) if a.op == 'advdiff' else stencils.neg_laplace_pairs(3)
offsets = [s for s, _ in shifts]; coeffs = [float(c) for _, c in shifts]
sol = SlabSolver(grid, offsets, coeffs, torch.float32, method=a.method, jacobi=a.jacobi, device=dev)
out = {"n": n, "world": world, "method": a.method, "op": a.op, "jacobi": a.jacobi}
L = int(a.method[-2]) if a.method.startswith('biCG-stab(') else 1
if a.check:
    y_full = torch.as_tensor(np.random.default_rng(0).standard_normal((1, n ** 3)).astype(np.float32), device=dev) if rank == 0 else None
    y_loc = torch.empty((1, sol.nxl * plane), dtype=torch.float32, device=dev)
    if world > 1:
        if rank == 0:
            for r in range(1, world):
                lo, hi = slab_bounds(n, world, r)
                dist.send(y_full[:, lo * plane:hi * plane].contiguous(), r)
            y_loc.copy_(y_full[:, :sol.nxl * plane])
        else:
            dist.recv(y_loc, 0)
    else:
        y_loc.copy_(y_full)
    x, r_, its, fev, conv, div = sol.solve(y_loc, torch.zeros_like(y_loc), [1e-5], [0.0], 5000)
    out.update(iterations=int(its[0]), function_evaluations=int(fev[0]), converged=bool(conv[0]), allreduces=sol.counts['allreduce'], halo_exchanges=sol.counts['halo'])
    if world > 1:
        if rank == 0:
            xs = [x] + [torch.empty((1, (slab_bounds(n, world, r)[1] - slab_bounds(n, world, r)[0]) * plane), dtype=torch.float32, device=dev) for r in range(1, world)]
            for r in range(1, world):
                dist.recv(xs[r], r)
            x_all = torch.cat(xs, dim=1)
        else:
            dist.send(x.contiguous(), 0)
    else:
        x_all = x
    if rank == 0:
        st = pb.Matrix.stencil_matrix(grid, offsets, coeffs, torch.float32, dev)
        pre = pb.JacobiPreconditioner(st) if a.jacobi else None
        from phiml_b200 import solve as S
        S.RESIDENT['mode'] = 'never'
        if a.method.startswith('biCG'):
            ref = pb.linear_solve(a.method, st, y_full, torch.zeros_like(y_full), [1e-5], [0.0], np.array([[5000]]), pre, None)
        elif a.method in ('CG-adaptive', 'auto') and not a.jacobi:
            ref = pb.generic_conjugate_gradient_adaptive(st, y_full, torch.zeros_like(y_full), [1e-5], [0.0], np.array([[5000]]))
        else:
            ref = pb.generic_conjugate_gradient(st, y_full, torch.zeros_like(y_full), [1e-5], [0.0], np.array([[5000]]), pre)
        out.update(ref_iterations=int(ref.iterations[0]), rel_err_vs_single_gpu=float((x_all - ref.x).abs().max() / ref.x.abs().max()),
                   true_residual=float((y_full - st.matvec(x_all)).norm() / y_full.norm()), ref_true_residual=float((y_full - st.matvec(ref.x)).norm() / y_full.norm()))
else:
    gen = torch.Generator(device=dev); gen.manual_seed(1000 + rank)
    y_loc = torch.randn((1, sol.nxl * plane), dtype=torch.float32, device=dev, generator=gen)
    x0 = torch.zeros_like(y_loc)

    def timed(iters):
        torch.cuda.synchronize(dev)
        if world > 1: dist.barrier()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        its = sol.solve(y_loc, x0, [0.0], [0.0], 2 ** 30, fixed_iterations=iters)[2]
        e1.record()
        torch.cuda.synchronize(dev)
        # CUDA events around the loop alone (set-up: allocation, IPC mapping, first residual stays outside), scaled to the requested count by the
        # iterations the loop really made (far beyond convergence BiCGstab breaks down and the device loop stops by itself)
        ms = sol.engine.last_loop_ms * iters / max(1, int(its.max()))
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t[0])
        return ms
    timed(a.warmup)
    timed(a.iters)                      # allocator and NCCL warm-up at the measured sizes
    c0 = dict(sol.counts)
    ms1 = timed(a.iters)
    c1 = dict(sol.counts)
    ms2 = timed(2 * a.iters)
    c2 = dict(sol.counts)
    ms1 = min(ms1, timed(a.iters))
    ms2 = min(ms2, timed(2 * a.iters))
    per_it = min(ms1 / a.iters, ms2 / (2 * a.iters))
    N = n ** 3
    model = (53 if a.method == 'biCG-stab(2)' else 13) * N * 4           # stencil BiCGstab(2) 53 N w, CG 13 N w
    out.update(iters=[a.iters, 2 * a.iters], ms_total=[ms1, ms2], ms_per_iteration=per_it, iterations_per_s=1e3 / per_it,
               survey_model_bytes_per_it=model, agg_gbs_survey_model=model / (per_it * 1e-3) / 1e9,
               allreduces_per_iteration=((c2['allreduce'] - c1['allreduce']) - (c1['allreduce'] - c0['allreduce'])) / a.iters,
               halo_exchanges_per_iteration=((c2['halo'] - c1['halo']) - (c1['halo'] - c0['halo'])) / a.iters,
               halo_bytes_per_rank_per_exchange=2 * plane * 4)
if rank == 0:
    print(json.dumps(out))
if world > 1:
    dist.destroy_process_group()
