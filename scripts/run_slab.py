"""
Multi-GPU slab CG check + throughput (run under torchrun, one rank per GPU):

    torchrun --nproc-per-node N scripts/run_slab.py --size 256 --check          # parity against a single-GPU solve on rank 0
    torchrun --nproc-per-node N scripts/run_slab.py --size 1024 --iters 100     # strong-scaling throughput (fixed iterations)

Prints one JSON line from rank 0.
"""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch, torch.distributed as dist
import phiml_b200 as pb
from phiml_b200.multi_gpu import SlabCG
from phiml_b200 import stencils

p = argparse.ArgumentParser()
p.add_argument('--size', dest='n', type=int, default=256)
p.add_argument('--iters', type=int, default=100)
p.add_argument('--warmup', type=int, default=10)
p.add_argument('--check', action='store_true')
p.add_argument('--nccl-reduce', action='store_true', help='NCCL all-reduce of the dot products instead of the in-kernel peer-memory all-reduce')
p.add_argument('--nccl-halos', action='store_true', help='separate NCCL send/recv halo exchange instead of peer-memory stores from the residual kernel')
a = p.parse_args()
world = int(os.environ.get('WORLD_SIZE', '1')); rank = int(os.environ.get('RANK', '0')); lr = int(os.environ.get('LOCAL_RANK', '0'))
torch.cuda.set_device(lr); dev = torch.device('cuda', lr)
if world > 1:
    dist.init_process_group('nccl', device_id=dev)
n = a.n; grid = (n, n, n); plane = n * n
shifts = stencils.neg_laplace_pairs(3); offsets = [s for s, _ in shifts]; coeffs = [c for _, c in shifts]
cg = SlabCG(grid, offsets, coeffs, torch.float32, jacobi=True, device=dev)
gen = torch.Generator(device=dev); gen.manual_seed(1000 + rank)
out = {"n": n, "world": world, "all_reduce": "nccl" if a.nccl_reduce or world == 1 else "in-kernel, peer-memory mailboxes", "halo_exchange": "nccl send/recv" if a.nccl_halos or world == 1 else "peer-memory stores from the residual kernel (CUDA IPC)"}
if a.check:
    # same right-hand side everywhere: generated per global plane index so that any partition sees identical data
    y_full = None
    if rank == 0:
        y_full = torch.as_tensor(np.random.default_rng(0).standard_normal((1, n ** 3)).astype(np.float32), device=dev)
    y_loc = torch.empty((1, cg.nxl * plane), dtype=torch.float32, device=dev)
    if world > 1:
        if rank == 0:
            for r in range(1, world):
                from phiml_b200.multi_gpu import slab_bounds
                lo, hi = slab_bounds(n, world, r)
                dist.send(y_full[:, lo * plane:hi * plane].contiguous(), r)
            y_loc.copy_(y_full[:, :cg.nxl * plane])
        else:
            dist.recv(y_loc, 0)
    else:
        y_loc.copy_(y_full)
    x, r_, its, fev, conv, div = cg.solve(y_loc, torch.zeros_like(y_loc), [1e-5], [0.0], 5000, peer_halos=(False if a.nccl_halos else None), peer_reduce=(False if a.nccl_reduce else None))
    out.update(iterations=int(its[0]), converged=bool(conv[0]))
    xs = [torch.empty((1, (slab_bounds(n, world, r)[1] - slab_bounds(n, world, r)[0]) * plane), dtype=torch.float32, device=dev) for r in range(world)] if world > 1 and rank == 0 else None
    if world > 1:
        if rank == 0:
            xs[0].copy_(x)
            for r in range(1, world):
                dist.recv(xs[r], r)
            x_all = torch.cat(xs, dim=1)
        else:
            dist.send(x.contiguous(), 0)
    else:
        x_all = x
    if rank == 0:
        st = pb.Matrix.stencil_matrix(grid, offsets, coeffs, torch.float32, dev)
        ref = pb.generic_conjugate_gradient(st, y_full, torch.zeros_like(y_full), [1e-5], [0.0], np.array([[5000]]), pb.JacobiPreconditioner(st))
        out.update(ref_iterations=int(ref.iterations[0]), rel_err=float((x_all - ref.x).abs().max() / ref.x.abs().max()),
                   true_residual=float((y_full - st.matvec(x_all)).norm() / y_full.norm()))
else:
    y_loc = torch.randn((1, cg.nxl * plane), dtype=torch.float32, device=dev, generator=gen)
    x0 = torch.zeros_like(y_loc)
    def timed(iters):
        torch.cuda.synchronize(dev)
        if world > 1: dist.barrier()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        cg.solve(y_loc, x0, [0.0], [0.0], 2 ** 30, fixed_iterations=iters, peer_halos=(False if a.nccl_halos else None), peer_reduce=(False if a.nccl_reduce else None))
        e1.record()
        torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t[0])
        return ms
    timed(a.warmup)
    # a solve call = set-up (extended arrays, initial residual, 2 halo exchanges) + iterations: the slope between two
    # iteration counts is the per-iteration time, the intercept the set-up
    ms1, ms2 = timed(a.iters), timed(2 * a.iters)
    per_it = (ms2 - ms1) / a.iters
    N = n ** 3
    out.update(iters=[a.iters, 2 * a.iters], ms_total=[ms1, ms2], ms_per_iteration=per_it, setup_ms=ms1 - per_it * a.iters, iterations_per_s=1e3 / per_it,
               bytes_design_per_it=9 * N * 4, agg_gbs_design=9 * N * 4 / (per_it * 1e-3) / 1e9, agg_gbs_survey_model=13 * N * 4 / (per_it * 1e-3) / 1e9,
               halo_bytes_per_rank_per_it=2 * plane * 4)
if rank == 0:
    print(json.dumps(out))
if world > 1:
    dist.destroy_process_group()
