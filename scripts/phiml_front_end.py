"""
The BASELINE configs through PhiML's OWN front end on the GPU (unmodified reference from baseline/_ref + phb200 plug-in):

    python scripts/phiml_front_end.py [--n 256] [--pre jacobi|ilu|none] [--method CG]

trace (host, NumPy) -> device assembly -> math.solve_linear, cold call and warm calls timed with time.perf_counter
around math.solve_linear (what SolveInfo.solve_time brackets is the Backend.linear_solve call only).  One JSON line.
"""
import argparse, json, os, sys, time, resource
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
from _phiml import import_phiml
assert import_phiml() is not None, "needs baseline/_ref (python -c 'import __graft_entry__ as g; g.build()' where /root/reference exists)"
from phiml import math
from phiml.math import spatial, Solve, SolveTape, extrapolation
from phiml_b200 import phiml_plugin, assembly
import phiml_b200 as pb

p = argparse.ArgumentParser()
p.add_argument('--n', type=int, default=256)
p.add_argument('--pre', default='jacobi')
p.add_argument('--method', default='CG')
p.add_argument('--reps', type=int, default=3)
p.add_argument('--op', default='zero', choices=['zero', 'neumann', 'periodic'], help="boundary of the Laplacian: ZERO (stencil kernels), ZERO_GRADIENT (+0.05 I; coded star, fused CG), PERIODIC (+0.05 I; coded stencil)")
a = p.parse_args()
TORCH = phiml_plugin.install()
n = a.n
OPS = {'zero': lambda x: -math.laplace(x, padding=extrapolation.ZERO),
       'neumann': lambda x: 0.05 * x - math.laplace(x, padding=extrapolation.ZERO_GRADIENT),
       'periodic': lambda x: 0.05 * x - math.laplace(x, padding=extrapolation.PERIODIC)}
f = math.jit_compile_linear(OPS[a.op])
out = {"n": n, "operator": a.op, "method": a.method, "preconditioner": a.pre, "calls": []}
with TORCH, math.precision(32):
    y_host = np.random.default_rng(0).standard_normal((n, n, n)).astype(np.float32)
    x0 = math.zeros(spatial(x=n, y=n, z=n))
    for rep in range(a.reps):
        torch.cuda.synchronize()
        l0 = pb.launch_count()
        t0 = time.perf_counter()
        y = math.tensor(y_host, spatial('x,y,z'))                      # host -> device inside the timed region
        solve = Solve(a.method, 1e-5, 0, x0=x0, max_iterations=5000, preconditioner=None if a.pre == 'none' else a.pre)
        with SolveTape() as tape:
            x = math.solve_linear(f, y, solve)
        x_host = x.numpy('x,y,z')                                       # device -> host
        dt = time.perf_counter() - t0
        info = tape[solve]
        its = int(info.iterations.numpy())
        out["calls"].append({"seconds": dt, "iterations": its, "iterations_per_s": its / dt, "backend_call_seconds": float(info.solve_time), "method": info.method,
                             "phb200_launches": pb.launch_count() - l0, "kind": "cold (trace + device assembly + preconditioner + solve)" if rep == 0 else "warm"})
    r = y - f(x)
    out["true_residual"] = float(math.l2_loss(r).numpy()) ** 0.5 / float(math.l2_loss(y).numpy()) ** 0.5
out["assembly"] = dict(assembly.STATS)
from phiml_b200 import sparse as _sp
out["natives"] = [{"kind": m.kind, "stencil": m.stencil is not None, "coded": m.coded is not None, "star_diag": bool(getattr(m.coded, 'star_diag', False)) if m.coded is not None else False} for m in _sp._CACHE.values()]
out["native_cache"] = {"hits": phiml_plugin.NATIVE_CACHE.hits, "misses": phiml_plugin.NATIVE_CACHE.misses}
out["host_peak_rss_gb"] = resource.getrusage(resource.RUSAGE_SELF).ru_maxrss / 1e6
out["device_peak_gb"] = torch.cuda.max_memory_allocated() / 1e9
print(json.dumps(out))
