"""Timing probe (one GPU): the fused Jacobi-CG pair on n^3 (default 256), us per iteration and per kernel, plus a checksum of x after a fixed
number of iterations so that variants (environment knobs) can be compared for identical results.   python scripts/time_cg.py [n] [tag]"""
import hashlib, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import phiml_b200 as pb
from phiml_b200 import _engine as E, stencils

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
tag = sys.argv[2] if len(sys.argv) > 2 else ''
dev = torch.device('cuda', 0)
sh = stencils.neg_laplace_pairs(3)
st = pb.Matrix.stencil_matrix((n, n, n), [s for s, _ in sh], [c for _, c in sh], torch.float32, dev)
pre = pb.JacobiPreconditioner(st)
N = n ** 3
g = torch.Generator(device=dev); g.manual_seed(1)
y = torch.randn((1, N), device=dev, generator=g)
zero = torch.zeros(1, device=dev); mi = torch.full((1,), 2 ** 30, dtype=torch.int32, device=dev)
s = pb.DeviceSolver(st, pre, E.CG, 0, 1)
s.start(y, torch.zeros_like(y), zero, zero, mi)
s.advance(100)
torch.cuda.synchronize()
sha = hashlib.sha256(s.x.cpu().numpy().tobytes()).hexdigest()[:16]
best = []
for _ in range(7):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); s.advance(300); e1.record(); torch.cuda.synchronize()
    best.append(e0.elapsed_time(e1) / 300)
s.profile(True); s.advance(60); prof = s.profile_read(); s.profile(False)
ms = sorted(best)[len(best) // 2]
print(json.dumps({'tag': tag, 'n': n, 'us_per_iteration': ms * 1e3, 'it_per_s': 1e3 / ms, 'min_max_us': [min(best) * 1e3, max(best) * 1e3],
                  'kernels_us': [p[0] / max(p[1], 1) * 1e3 for p in prof[:2]], 'frac_of_peak_9Nw': 9 * N * 4 / (ms * 1e-3) / 1e9 / 6552.3, 'x_sha_after_100': sha,
                  'env': {k: v for k, v in os.environ.items() if k.startswith('PHB200_')}}))
