"""In-kernel cycle counters of the x-marching triangular-solve pipelines (PHB200_TRSV_DBG): per warp total / boundary wait / ring wait / back-pressure."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import phiml_b200 as pb
from phiml_b200 import stencils
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dev = torch.device('cuda', 0)
sh = stencils.neg_laplace_pairs(3)
st = pb.Matrix.stencil_matrix((n, n, n), [s for s, _ in sh], [c for _, c in sh], torch.float32, dev)
csr = st.to_csr()
m = pb.as_matrix(torch.sparse_csr_tensor(csr.rowptr, csr.col, csr.values, csr.shape))
ilu = pb.IncompleteLU(m)
v = torch.randn(1, n ** 3, device=dev)
ilu.apply_inv_l(v) if False else None
R = 8
nj, nch = (n + R - 1) // R, (n + 127) // 128
dbg = torch.zeros(nj * nch * R * 4, dtype=torch.int64, device=dev)
for _ in range(3):
    ilu.apply(v)
torch.cuda.synchronize()
os.environ['PHB200_TRSV_DBG'] = str(dbg.data_ptr())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); ilu.apply(v); e1.record(); torch.cuda.synchronize()
d = dbg.cpu().numpy().reshape(nj * nch, R, 4).astype(np.float64) / 1.9e3     # us at ~1.9 GHz  (values of the UPPER solve, the last launch)
print(f"n={n} apply {e0.elapsed_time(e1)*1e3:.0f} us (L+U); upper solve, per warp, us: total / boundary / ring / back-pressure")
for cta in (0, 1, nj * nch // 2, nj * nch - 1):
    for w in (0, 1, R - 1):
        print(f"  cta(start order) {cta:4d} warp {w}: " + " / ".join(f"{x:8.1f}" for x in d[cta, w]))
print("mean over all warps:", " / ".join(f"{x:8.1f}" for x in d.reshape(-1, 4).mean(0)), " max total", d[..., 0].max())
