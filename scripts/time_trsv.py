"""Times IncompleteLU.apply (lower + upper solve) for the star wavefront and CSR level forms.  usage: time_trsv.py n [n ...]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
import phiml_b200 as pb
from phiml_b200 import stencils
dev = torch.device('cuda', 0)
for n in [int(v) for v in sys.argv[1:]] or [128, 256]:
    sh = stencils.neg_laplace_pairs(3)
    st = pb.Matrix.stencil_matrix((n, n, n), [s for s, _ in sh], [c for _, c in sh], torch.float32, dev)
    csr = st.to_csr()
    m = pb.as_matrix(torch.sparse_csr_tensor(csr.rowptr, csr.col, csr.values, csr.shape))
    t0 = time.perf_counter(); ilu = pb.IncompleteLU(m); torch.cuda.synchronize(); t_setup = time.perf_counter() - t0
    v = torch.randn(1, n ** 3, device=dev)
    out = ilu.apply(v)
    reps = 10
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps):
        out = ilu.apply(v)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    N = n ** 3
    print(json.dumps({"n": n, "star_wavefront": ilu.star_wavefront, "setup_s": t_setup, "apply_ms (L+U)": ms, "GB/s on 6Nw": 6 * N * 4 / ms / 1e6,
                      "env": {k: v for k, v in os.environ.items() if k.startswith('PHB200')}}), flush=True)
