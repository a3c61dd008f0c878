"""GPU: time breakdown of an end-to-end solve (host RHS -> device -> solve -> host)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import numpy as np, torch
import phiml_b200 as pb
from phiml_b200 import _engine as E
from oracle.assemble import laplace_shifts
dev = torch.device('cuda:0')
n = 256; N = n ** 3
shifts = laplace_shifts(3)
stencil = pb.Matrix.stencil_matrix((n, n, n), [s for s, _ in shifts], [c for _, c in shifts], torch.float32, dev)
csr = stencil.to_csr()
native = torch.sparse_csr_tensor(csr.rowptr, csr.col, csr.values, csr.shape)
y_host = torch.as_tensor(np.random.default_rng(0).standard_normal((1, N)).astype(np.float32)).pin_memory()
def T():
    torch.cuda.synchronize(); return time.perf_counter()
for rep in range(3):
    pb.clear_cache()
    t0 = T()
    yd = y_host.to(dev, non_blocking=True); x0d = torch.zeros_like(yd)
    t1 = T()
    m = pb.as_matrix(native)
    t2 = T()
    p = pb.JacobiPreconditioner(m)
    t3 = T()
    rtol = torch.full((1,), 1e-5, device=dev); atol = torch.zeros(1, device=dev); mi = torch.full((1,), 5000, dtype=torch.int32, device=dev)
    s = pb.DeviceSolver(m.stencil, p, E.CG, 0, 1)
    t4 = T()
    s.start(yd, x0d, rtol, atol, mi)
    t5 = T()
    s.run(5000)
    t6 = T()
    its = s.result()[0].cpu()
    t7 = T()
    passes, launches = s.stats()
    print(f"rep {rep}: h2d {t1-t0:.4f} as_matrix {t2-t1:.4f} jacobi {t3-t2:.4f} create {t4-t3:.4f} start {t5-t4:.4f} run {t6-t5:.4f} result {t7-t6:.4f} its {its.tolist()} passes {passes} launches {launches}")
    del s
    t8 = T()
    print(f"   destroy {t8-t7:.4f}")
