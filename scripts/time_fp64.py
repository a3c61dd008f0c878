"""Device-loop throughput of the fused Jacobi-CG on 256^3 in fp64 (and fp32 for comparison)."""
import sys, os, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
import phiml_b200 as pb
from phiml_b200 import _engine as E
from phiml_b200 import stencils
dev = torch.device("cuda", 0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
for dt in (torch.float64, torch.float32):
    sh = stencils.neg_laplace_pairs(3)
    st = pb.Matrix.stencil_matrix((n, n, n), [s for s, _ in sh], [float(c) for _, c in sh], dt, dev)
    pre = pb.JacobiPreconditioner(st)
    y = torch.randn(1, n ** 3, device=dev, dtype=dt)
    z = torch.zeros(1, device=dev, dtype=dt); mi = torch.full((1,), 2 ** 30, dtype=torch.int32, device=dev)
    solver = pb.DeviceSolver(st, pre, E.CG, 0, 1)
    solver.start(y, torch.zeros_like(y), z, z, mi); solver.advance(20); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); solver.advance(300); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 300
    w = 8 if dt == torch.float64 else 4
    print(json.dumps({"n": n, "dtype": str(dt), "us_per_iteration": ms * 1e3, "it_per_s": 1e3 / ms, "gbs_on_9Nw": 9 * n ** 3 * w / ms / 1e6, "stages": os.environ.get("X", "")}))
