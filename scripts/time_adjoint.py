"""
Timings of the implicit-gradient step on one GPU: matrix-gradient kernels and the adjoint solve at n^3.
    python scripts/time_adjoint.py [--size 256]
Prints one JSON line per measurement.
"""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
import phiml_b200 as pb
from phiml_b200 import adjoint
from phiml_b200.sparse import plain_csr
from phiml_b200 import stencils
ADVDIFF = stencils.ADVDIFF

p = argparse.ArgumentParser()
p.add_argument('--size', dest='n', type=int, default=256)
a = p.parse_args()
dev = torch.device('cuda:0')
n = a.n; N = n ** 3
peak = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json'))).get('hbm_gbs', 6552.3) if os.path.exists(os.path.join(ROOT, 'MEASURED_PEAKS.json')) else 6552.3
shifts = stencils.advection_diffusion_pairs(3, ADVDIFF['velocity'], ADVDIFF['c'], ADVDIFF['nu'], np.float32)
st = pb.Matrix.stencil_matrix((n, n, n), [s for s, _ in shifts], [float(c) for _, c in shifts], torch.float32, dev)
csr = st.to_csr()
nat = torch.sparse_csr_tensor(csr.rowptr, csr.col, csr.values, csr.shape)
nnz = int(csr.values.numel())
gen = torch.Generator(device=dev); gen.manual_seed(0)
dy = torch.randn((1, N), device=dev, generator=gen); x = torch.randn((1, N), device=dev, generator=gen)


def timed(fn, reps=20):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


ms = timed(lambda: adjoint.matrix_gradient(nat, dy, x))
bytes_csr = 4 * nnz + 4 * (N + 1) + 4 * nnz + 2 * N * 4          # col + rowptr + out + dy + x
print(json.dumps({"kernel": "k_matrix_gradient_csr", "n": n, "nnz": nnz, "ms": ms, "algorithmic_bytes": bytes_csr, "gbs": bytes_csr / ms / 1e6, "frac_of_peak": bytes_csr / ms / 1e6 / peak}))
rows = torch.repeat_interleave(torch.arange(N, device=dev), (csr.rowptr[1:] - csr.rowptr[:-1]).to(torch.int64))
coo = torch.sparse_coo_tensor(torch.stack([rows, csr.col.to(torch.int64)]), csr.values, csr.shape)
ms = timed(lambda: adjoint.matrix_gradient(coo, dy, x))
bytes_coo = 16 * nnz + 4 * nnz + 2 * N * 4
print(json.dumps({"kernel": "k_matrix_gradient_coo", "n": n, "nnz": nnz, "ms": ms, "algorithmic_bytes": bytes_coo, "gbs": bytes_coo / ms / 1e6, "frac_of_peak": bytes_coo / ms / 1e6 / peak}))
del coo, rows
# torch formulation of the reference (_optimize.py:805-806): two gathers, product, negation
ri = torch.repeat_interleave(torch.arange(N, device=dev), (csr.rowptr[1:] - csr.rowptr[:-1]).to(torch.int64)); ci = csr.col.to(torch.int64)
ms_t = timed(lambda: -(dy[0][ri] * x[0][ci]), reps=5)
print(json.dumps({"kernel": "torch gathers (reference formulation, indices already expanded)", "ms": ms_t}))
del ri, ci

# adjoint solve: A^T dy = dx with the forward method; set-up (transposition of the CSC view, stencil recognition) timed separately
y = torch.randn((1, N), device=dev, generator=gen)
rt, at, mi = np.array([1e-5]), np.array([0.0]), np.array([[5000]])
fwd = pb.linear_solve('biCG-stab(2)', nat, y, torch.zeros_like(y), rt, at, mi, None, None)
torch.cuda.synchronize(); t0 = time.perf_counter()
t = adjoint.transpose(nat); m = plain_csr(pb.as_matrix(t)); torch.cuda.synchronize(); t1 = time.perf_counter()
res, dm = adjoint.implicit_gradient_solve('biCG-stab(2)', nat, fwd.x, fwd.x, rt, at, mi, None, matrix_adjoint=True); torch.cuda.synchronize(); t2 = time.perf_counter()
res, dm = adjoint.implicit_gradient_solve('biCG-stab(2)', nat, fwd.x, fwd.x, rt, at, mi, None, matrix_adjoint=True); torch.cuda.synchronize(); t3 = time.perf_counter()
r_true = float((fwd.x - m.matvec(res.x)).norm() / fwd.x.norm())
print(json.dumps({"adjoint": "biCG-stab(2) on A^T + dm", "n": n, "stencil_recognised_on_transpose": m.stencil is not None, "transpose_setup_s": t1 - t0, "first_call_s": t2 - t1, "second_call_s": t3 - t2,
                  "iterations": res.iterations.tolist(), "forward_iterations": fwd.iterations.tolist(), "true_residual": r_true}))
