import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
import phiml_b200 as pb
from phiml_b200 import stencils
dev = torch.device('cuda', 0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
sh = stencils.neg_laplace_pairs(3)
st = pb.Matrix.stencil_matrix((n, n, n), [s for s, _ in sh], [c for _, c in sh], torch.float32, dev)
csr = st.to_csr()
m = pb.as_matrix(torch.sparse_csr_tensor(csr.rowptr, csr.col, csr.values, csr.shape))
ilu = pb.IncompleteLU(m)
v = torch.randn(1, n ** 3, device=dev)
ilu.apply(v); ilu.apply(v)
ni = nj = (n + 7) // 8; nch = (n + 31) // 32
dbg = torch.zeros(ni * nj * nch * 6, dtype=torch.int64, device=dev)
os.environ['PHB200_TRSV_DBG'] = str(dbg.data_ptr())
out = ilu.apply(v)          # stamps of the UPPER solve overwrite those of the lower one
torch.cuda.synchronize()
t = dbg.cpu().numpy().reshape(ni, nj, nch, 6).astype(np.int64)
t0 = t[..., 0].min()
mid = t[ni // 2, nj // 2]
print("tile(mid) per chunk: start, +flagwait, +fence, +boundary/cp.async, +8 tasks, +publish (ns)")
for c in range(nch):
    r = mid[c]
    print(c, r[0] - t0, r[1] - r[0], r[2] - r[1], r[3] - r[2], r[4] - r[3], r[5] - r[4])
d = np.diff(t, axis=-1)
print("mean over all tiles/chunks (ns): flagwait %.0f fence %.0f boundary %.0f tasks %.0f publish %.0f" % tuple(d.reshape(-1, 5).mean(0)))
print("median: flagwait %.0f fence %.0f boundary %.0f tasks %.0f publish %.0f" % tuple(np.median(d.reshape(-1, 5), axis=0)))
print("total span (us):", (t[..., 5].max() - t0) / 1e3)
# critical path estimate: time from the -x neighbour's publish of chunk c to this tile's start of tasks
lag = t[1:, :, :, 3] - t[:-1, :, :, 5]
print("x hand-over: publish(prev tile) -> tasks start here, median ns:", np.median(lag), " p10", np.percentile(lag, 10))
