"""Timing probes (one GPU): per-iteration time of the fused Jacobi-CG pair on grids (nx, 256, 256) — what one slab rank does without any
communication — and the stand-alone CSR / stencil products.   python scripts/time_misc.py [slab] [csr]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import phiml_b200 as pb
from phiml_b200 import _engine as E, stencils

dev = torch.device('cuda', 0)
what = sys.argv[1:] or ['slab', 'csr']
sh = stencils.neg_laplace_pairs(3)
off, co = [s for s, _ in sh], [c for _, c in sh]


def timed(fn, reps=5):
    best = 1e30
    for _ in range(reps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


out = {}
if 'slab' in what:
    for nx in (256, 128, 64, 32):
        st = pb.Matrix.stencil_matrix((nx, 256, 256), off, co, torch.float32, dev)
        pre = pb.JacobiPreconditioner(st)
        N = nx * 65536
        y = torch.randn((1, N), device=dev)
        zero = torch.zeros(1, device=dev); mi = torch.full((1,), 2 ** 30, dtype=torch.int32, device=dev)
        s = pb.DeviceSolver(st, pre, E.CG, 0, 1)
        s.start(y, torch.zeros_like(y), zero, zero, mi)
        s.advance(20)
        ms = timed(lambda: s.advance(200)) / 200
        s.profile(True); s.advance(50); prof = s.profile_read(); s.profile(False)
        out[f'jacobi_cg_{nx}x256x256'] = {'us_per_iteration': ms * 1e3, 'kernels_us': [p[0] / max(p[1], 1) * 1e3 for p in prof[:2]], 'frac_of_peak_9Nw': 9 * N * 4 / (ms * 1e-3) / 1e9 / 6552.3}
if 'csr' in what:
    n = 256; N = n ** 3
    st = pb.Matrix.stencil_matrix((n, n, n), off, co, torch.float32, dev)
    csr = st.to_csr()
    xs = [torch.randn((1, N), device=dev) for _ in range(4)]
    ys = [torch.empty_like(xs[0]) for _ in range(4)]
    lib, sp = E.lib(), E.stream_ptr(dev)
    def run(h):
        def f():
            for i in range(20):
                E.check(lib.phb200_spmv(h, E.ptr(xs[i % 4]), E.ptr(ys[i % 4]), 1, N, N, sp))
        return f
    b = 8 * csr.values.numel() + 4 * (N + 1) + 2 * N * 4
    f = run(csr.handle); f()
    ms = timed(f) / 20
    out['spmv_csr_256'] = {'us': ms * 1e3, 'gbs': b / ms / 1e6, 'frac_of_peak': b / ms / 1e6 / 6552.3, 'bulk': os.environ.get('PHB200_CSR_BULK', '1')}
    ref = st.matvec(xs[0]); got = csr.matvec(xs[0], use_stencil=False)
    out['spmv_csr_256']['bit_identical_to_stencil'] = bool(torch.equal(ref, got))
    f = run(st.handle); f()
    ms = timed(f) / 20
    out['spmv_stencil_256'] = {'us': ms * 1e3, 'frac_of_peak': 2 * N * 4 / ms / 1e6 / 6552.3}
    pre = pb.JacobiPreconditioner(csr)
    y = torch.randn((1, N), device=dev)
    zero = torch.zeros(1, device=dev); mi = torch.full((1,), 2 ** 30, dtype=torch.int32, device=dev)
    s = pb.DeviceSolver(csr, pre, E.CG, 0, 1)
    s.start(y, torch.zeros_like(y), zero, zero, mi)
    s.advance(10)
    ms = timed(lambda: s.advance(50)) / 50
    s.profile(True); s.advance(30); prof = s.profile_read(); s.profile(False)
    out['jacobi_cg_csr_256'] = {'it_per_s': 1e3 / ms, 'kernels_us': [p[0] / max(p[1], 1) * 1e3 for p in prof[:3]]}
    # batch of 64 right-hand sides on a 2-D matrix (tiles of 8 share a staged chunk)
    sh2 = stencils.neg_laplace_pairs(2)
    st2 = pb.Matrix.stencil_matrix((1024, 1024), [s_ for s_, _ in sh2], [c for _, c in sh2], torch.float32, dev)
    c2 = st2.to_csr()
    X = torch.randn((64, 1024 * 1024), device=dev)
    ms = timed(lambda: c2.matvec(X, use_stencil=False))
    out['spmv_csr_1024sq_batch64'] = {'us': ms * 1e3, 'equal_to_stencil': bool(torch.equal(c2.matvec(X, use_stencil=False), st2.matvec(X)))}
print(json.dumps(out))
