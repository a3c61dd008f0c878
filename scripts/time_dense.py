"""Timing probe (one GPU): `matrix @ x` outside the solver — Backend.mul_csr_dense / mul_coo_dense with (batch, channel) pairs — as ONE launch of
phb200_csr_dense_batched against the per-pair loop the reference (and round 1 of this repo) runs.   python scripts/time_dense.py [n=64]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import phiml_b200 as pb
from phiml_b200 import _engine as E, stencils

dev = torch.device('cuda', 0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
sh = stencils.neg_laplace_pairs(3)
st = pb.Matrix.stencil_matrix((n, n, n), [s for s, _ in sh], [c for _, c in sh], torch.float32, dev)
csr = st.to_csr()
rowptr, col = csr.rowptr, csr.col
N, nnz = n ** 3, int(col.numel())


def timed(fn, reps=5):
    best = 1e30
    for _ in range(reps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


out = {'n': n, 'rows': N, 'nnz': nnz, 'cases': []}
g = torch.Generator(device=dev); g.manual_seed(0)
for batch, ch, k in [(8, 2, 4), (16, 1, 1), (4, 4, 1), (1, 1, 8)]:
    vals = torch.randn((batch, nnz, ch), device=dev, generator=g)
    dense = torch.randn((batch, N, ch, k), device=dev, generator=g)
    ci, rp = col[None].expand(batch, nnz).contiguous(), rowptr[None].expand(batch, N + 1).contiguous()
    one = lambda: pb.mul_csr_dense(ci, rp, vals, (N, N), dense)

    def pairs():
        res = torch.empty((batch, N, ch, k), device=dev)
        for b in range(batch):
            for c in range(ch):
                res[b, :, c, :] = pb.mul_csr_dense(ci[b:b + 1], rp[b:b + 1], vals[b:b + 1, :, c:c + 1].contiguous(), (N, N), dense[b:b + 1, :, c:c + 1, :].contiguous())[0, :, 0, :]
        return res
    a, b_ = one(), pairs()
    l0 = E.launch_count(); one(); l1 = E.launch_count()
    ms_one, ms_pairs = timed(one), timed(pairs)
    bytes_ = batch * (nnz * (4 + 4 * ch) + (N + 1) * 4 + 2 * N * ch * k * 4)
    out['cases'].append({'batch': batch, 'channels': ch, 'dense_cols': k, 'one_launch_ms': ms_one, 'launches': l1 - l0, 'pair_loop_ms': ms_pairs, 'bit_identical': bool(torch.equal(a, b_)),
                         'algorithmic_bytes': bytes_, 'gbs': bytes_ / ms_one / 1e6})
print(json.dumps(out))
