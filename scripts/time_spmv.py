"""Stand-alone stencil / CSR product through phb200_spmv at n^3 (for ncu captures of k_star_spmv_tma / k_spmv_csr_stream)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
import phiml_b200 as pb
from phiml_b200 import _engine as E
from phiml_b200 import stencils

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
dev = torch.device('cuda:0'); N = n ** 3
shifts = stencils.neg_laplace_pairs(3)
st = pb.Matrix.stencil_matrix((n, n, n), [s for s, _ in shifts], [float(c) for _, c in shifts], torch.float32, dev)
x = torch.randn((1, N), device=dev); y = torch.empty_like(x)
for m in (st,) + ((st.to_csr(),) if 'csr' in sys.argv else ()):
    for _ in range(reps):
        E.check(E.lib().phb200_spmv(m.handle, E.ptr(x), E.ptr(y), 1, N, N, E.stream_ptr(dev)))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        E.check(E.lib().phb200_spmv(m.handle, E.ptr(x), E.ptr(y), 1, N, N, E.stream_ptr(dev)))
    e1.record(); torch.cuda.synchronize()
    print(m.kind, e0.elapsed_time(e1) / 50 * 1e3, 'us')
