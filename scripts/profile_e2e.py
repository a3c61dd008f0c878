"""Host-side profile of the end-to-end call of bench.py (where do the milliseconds outside the kernels go?)."""
import cProfile, pstats, io, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
import phiml_b200 as pb
from phiml_b200 import stencils

dev = torch.device('cuda:0')
n = 256; N = n ** 3
shifts = stencils.neg_laplace_pairs(3)
st = pb.Matrix.stencil_matrix((n, n, n), [s for s, _ in shifts], [float(c) for _, c in shifts], torch.float32, dev)
csr = st.to_csr()
native = torch.sparse_csr_tensor(csr.rowptr, csr.col, csr.values, csr.shape)
y_host = torch.as_tensor(np.random.default_rng(0).standard_normal((1, N)).astype(np.float32)).pin_memory()
x0_host = torch.zeros((1, N), dtype=torch.float32).pin_memory()
backend = pb.B200Backend(dev)


def once():
    pb.clear_cache()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    yd = y_host.to(dev, non_blocking=True); x0d = x0_host.to(dev, non_blocking=True)
    m = pb.as_matrix(native); p = backend.compute_preconditioner('jacobi', m)
    res = backend.linear_solve('CG', m, yd, x0d, [1e-5], [0.0], np.array([[5000]]), p, None)
    its = int(res.iterations.cpu()[0]); torch.cuda.synchronize()
    return its, time.perf_counter() - t0


once(); once()
pr = cProfile.Profile(); pr.enable()
for _ in range(3): print(once())
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats('cumulative').print_stats(28); print(s.getvalue()[:6000])
