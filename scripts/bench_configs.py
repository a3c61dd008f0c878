"""
Timings of the other BASELINE.json configs on one GPU (not the contract bench; see bench.py for config 2).
    python scripts/bench_configs.py [cfg1] [cfg3] [cfg4] [cfg5]  [--batch 4096] [--n3 256]
Prints one JSON line per config.
"""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
import phiml_b200 as pb
from phiml_b200 import stencils
ADVDIFF = stencils.ADVDIFF

p = argparse.ArgumentParser()
p.add_argument('configs', nargs='*', default=['cfg1', 'cfg3', 'cfg4', 'cfg5'])
p.add_argument('--batch', type=int, default=4096)
p.add_argument('--n4', type=int, default=128)
p.add_argument('--n5', type=int, default=256)
a = p.parse_args()
dev = torch.device('cuda:0')


def stencil(grid, shifts, dtype):
    return pb.Matrix.stencil_matrix(grid, [s for s, _ in shifts], [float(c) for _, c in shifts], dtype, dev)


def timed(fn, reps=2):
    best = None
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        out = fn()
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return out, best


for cfg in a.configs:
    if cfg == 'cfg1':   # 2-D 128x128 CG fp64 (reference: 488 iterations at default tolerances, 0.14-0.19 s on the CPU)
        st = stencil((128, 128), stencils.neg_laplace_pairs(2), torch.float64)
        y = torch.as_tensor(np.random.default_rng(0).standard_normal((1, 128 * 128)), device=dev)
        res, dt = timed(lambda: pb.generic_conjugate_gradient(st, y, torch.zeros_like(y), [1e-12], [1e-12], np.array([[5000]])))
        print(json.dumps({"config": "cfg1 2D 128^2 CG fp64 default tol", "iterations": res.iterations.tolist(), "reference_iterations": 488, "seconds": dt,
                          "iterations_per_s": float(res.iterations.max()) / dt}))
    elif cfg == 'cfg3':  # batch x 128^2 fp32 CG, shared matrix (reference B=4096: 384 s, per-system iterations 236..299)
        B = a.batch
        st = stencil((128, 128), stencils.neg_laplace_pairs(2), torch.float32)
        y = torch.as_tensor(np.random.default_rng(0).standard_normal((B, 128 * 128)).astype(np.float32), device=dev)
        res, dt = timed(lambda: pb.generic_conjugate_gradient(st, y, torch.zeros_like(y), [1e-5] * B, [0.0] * B, np.full((1, B), 5000)))
        its = res.iterations.cpu().numpy()
        N = 128 * 128
        print(json.dumps({"config": f"cfg3 batch {B} x 128^2 CG fp32", "iterations_min_max": [int(its.min()), int(its.max())], "reference_min_max_B4096": [236, 299],
                          "converged": bool(res.converged.all()), "seconds": dt, "loop_iterations_per_s": float(its.max()) / dt, "system_iterations_per_s": float(its.sum()) / dt,
                          "gbs_survey_model_13SNw": 13 * B * N * 4 * float(its.max()) / dt / 1e9}))
    elif cfg == 'cfg4':  # ILU(0)-CG fp32 (reference at 128^3: 97 iterations, 40.5 s in the loop)
        n = a.n4
        st = stencil((n, n, n), stencils.neg_laplace_pairs(3), torch.float32)
        t0 = time.perf_counter(); csr = st.to_csr(); nat = torch.sparse_csr_tensor(csr.rowptr, csr.col, csr.values, csr.shape); m = pb.as_matrix(nat)
        torch.cuda.synchronize(); t1 = time.perf_counter()
        ilu = pb.IncompleteLU(m); torch.cuda.synchronize(); t2 = time.perf_counter()
        y = torch.as_tensor(np.random.default_rng(0).standard_normal((1, n ** 3)).astype(np.float32), device=dev)
        res, dt = timed(lambda: pb.generic_conjugate_gradient(m, y, torch.zeros_like(y), [1e-5], [0.0], np.array([[5000]]), ilu))
        print(json.dumps({"config": f"cfg4 {n}^3 ILU(0)-CG fp32", "iterations": res.iterations.tolist(), "reference_iterations_128": 97, "levels": ilu.levels,
                          "assemble_s": t1 - t0, "factor_s": t2 - t1, "solve_s": dt, "iterations_per_s": float(res.iterations.max()) / dt}))
    elif cfg == 'cfg5':  # nonsymmetric advection-diffusion, biCG-stab(2) fp32
        n = a.n5
        st = stencil((n, n, n), stencils.advection_diffusion_pairs(3, ADVDIFF['velocity'], ADVDIFF['c'], ADVDIFF['nu'], np.float32), torch.float32)
        y = torch.as_tensor(np.random.default_rng(0).standard_normal((1, n ** 3)).astype(np.float32), device=dev)
        res, dt = timed(lambda: pb.linear_solve('biCG-stab(2)', st, y, torch.zeros_like(y), [1e-5], [0.0], np.array([[5000]]), None, None))
        r_true = float((y - st.matvec(res.x)).norm() / y.norm())
        print(json.dumps({"config": f"cfg5 {n}^3 advection-diffusion biCG-stab(2) fp32", "iterations": res.iterations.tolist(), "function_evaluations": res.function_evaluations.tolist(),
                          "converged": bool(res.converged.all()), "true_residual": r_true, "seconds": dt, "iterations_per_s": float(res.iterations.max()) / dt}))
