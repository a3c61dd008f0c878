"""Times DeviceSolver.start / run separately for the small-system configs (cfg1, cfg3), resident vs streaming kernels."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
import phiml_b200 as pb
from phiml_b200 import _engine as E, solve as S
from phiml_b200 import stencils

dev = torch.device('cuda', 0)


def stencil(grid, dtype):
    sh = stencils.neg_laplace_pairs(len(grid))
    return pb.Matrix.stencil_matrix(grid, [s for s, _ in sh], [c for _, c in sh], dtype, dev)


def run(name, grid, dtype, B, method, rtol, atol, modes=('never', 'always')):
    st = stencil(grid, dtype)
    n = int(np.prod(grid))
    npdt = np.float32 if dtype == torch.float32 else np.float64
    y = torch.as_tensor(np.random.default_rng(0).standard_normal((B, n)).astype(npdt), device=dev)
    x0 = torch.zeros_like(y)
    rt = torch.full((B,), rtol, dtype=dtype, device=dev); at = torch.full((B,), atol, dtype=dtype, device=dev)
    mi = torch.full((B,), 5000, dtype=torch.int32, device=dev)
    for mode in modes:
        S.RESIDENT['mode'] = mode
        best = None
        for rep in range(3):
            solver = pb.DeviceSolver(st, None, method, 0, B)
            e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            torch.cuda.synchronize(); t0 = time.perf_counter()
            e0.record(); solver.start(y, x0, rt, at, mi); e1.record(); solver.run(5000); e2.record()
            torch.cuda.synchronize(); wall = time.perf_counter() - t0
            its = solver.result()[0].cpu().numpy()
            rec = {"config": name, "mode": mode, "resident_used": solver.resident_used, "start_ms": e0.elapsed_time(e1), "run_ms": e1.elapsed_time(e2), "wall_ms": wall * 1e3,
                   "iterations_min_max": [int(its.min()), int(its.max())], "us_per_loop_iteration": e1.elapsed_time(e2) * 1e3 / max(1, int(its.max())),
                   "system_iterations_per_s": float(its.sum()) / (e1.elapsed_time(e2) * 1e-3)}
            if best is None or rec['run_ms'] < best['run_ms']:
                best = rec
        print(json.dumps(best), flush=True)
    S.RESIDENT['mode'] = 'auto'


which = sys.argv[1:] or ['cfg1', 'cfg3']
if 'cfg1' in which:
    run('cfg1 128^2 fp64 CG generic', (128, 128), torch.float64, 1, E.CG, 1e-12, 1e-12)
    run('cfg1 128^2 fp64 CG torch rules', (128, 128), torch.float64, 1, E.CG_TORCH, 1e-12, 1e-12)
    run('128^2 fp32 CG generic B=1', (128, 128), torch.float32, 1, E.CG, 1e-5, 0.0)
if 'cfg3' in which:
    run('cfg3 4096 x 128^2 fp32 CG generic', (128, 128), torch.float32, 4096, E.CG, 1e-5, 0.0)
    run('cfg3 4096 x 128^2 fp32 CG torch rules', (128, 128), torch.float32, 4096, E.CG_TORCH, 1e-5, 0.0)
if 'auto' in which:     # PhiML's default method ('auto' -> CG-adaptive, torch rules)
    run('128^2 fp32 auto (CG-adaptive torch rules) B=1', (128, 128), torch.float32, 1, E.CG_ADAPTIVE_TORCH, 1e-5, 0.0)
    run('128^2 fp64 auto (CG-adaptive torch rules) B=1', (128, 128), torch.float64, 1, E.CG_ADAPTIVE_TORCH, 1e-10, 0.0)
    run('256 x 128^2 fp32 auto (CG-adaptive torch rules)', (128, 128), torch.float32, 256, E.CG_ADAPTIVE_TORCH, 1e-5, 0.0)
if 'small3d' in which:
    run('32^3 fp32 CG generic B=64', (32, 32, 32), torch.float32, 64, E.CG, 1e-5, 0.0)
    run('24^3 fp32 CG generic B=512', (24, 24, 24), torch.float32, 512, E.CG, 1e-5, 0.0)
