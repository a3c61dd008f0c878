"""Product and Jacobi-CG timings of pattern-coded stencils against the CSR kernels on the same matrix (n^3 Neumann / periodic Laplacian + 0.01 I, fp32).
usage: time_coded.py [n] [neumann|periodic]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import phiml_b200 as pb
from phiml_b200 import _engine as E

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
kind = sys.argv[2] if len(sys.argv) > 2 else 'neumann'
dev = torch.device('cuda', 0)
N = n ** 3


def assemble():
    """CSR of the 7-point operator with ZERO_GRADIENT or PERIODIC boundaries, built with torch ops on the device (set-up only)."""
    idx = torch.arange(N, device=dev, dtype=torch.int64)
    iz, iy, ix = idx % n, (idx // n) % n, idx // (n * n)
    cols, vals = [], []
    deg = torch.zeros(N, device=dev, dtype=torch.float32)
    for comp, stride in ((ix, n * n), (iy, n), (iz, 1)):
        for s in (-1, 1):
            j = comp + s
            if kind == 'periodic':
                inside = torch.ones_like(j, dtype=torch.bool)
                jj = (j % n - comp) * stride + idx
            else:
                inside = (j >= 0) & (j < n)
                jj = idx + s * stride
            cols.append(torch.where(inside, jj, torch.full_like(jj, -1)))
            vals.append(torch.where(inside, -torch.ones(N, device=dev), torch.zeros(N, device=dev)))
            deg += inside.float()
    cols.append(idx); vals.append(deg + 0.01)
    C = torch.stack(cols, 1); V = torch.stack(vals, 1)
    C = torch.where(C < 0, torch.full_like(C, 2 ** 40), C)
    order = torch.argsort(C, dim=1)
    C = torch.gather(C, 1, order); V = torch.gather(V, 1, order)
    keep = C < 2 ** 40
    rowptr = torch.zeros(N + 1, dtype=torch.int32, device=dev)
    rowptr[1:] = torch.cumsum(keep.sum(1), 0).to(torch.int32)
    return rowptr, C[keep].to(torch.int32), V[keep]


rowptr, col, val = assemble()
t0 = time.perf_counter()
m = pb.as_matrix(torch.sparse_csr_tensor(rowptr, col, val, (N, N)))
torch.cuda.synchronize()
t_rec = time.perf_counter() - t0
assert m.coded is not None
out = {'n': n, 'kind': kind, 'nnz': int(val.numel()), 'patterns': m.coded.patterns, 'offsets': m.coded.offsets, 'recognition_s': t_rec}
xs = [torch.randn((1, N), device=dev) for _ in range(4)]
ys = [torch.empty((1, N), device=dev) for _ in range(4)]
for name, h, nbytes in (('coded', m.coded.handle, 9 * N), ('csr', m.handle, 8 * int(val.numel()) + 4 * (N + 1) + 8 * N)):
    for r in range(8):
        E.check(E.lib().phb200_spmv(h, E.ptr(xs[r % 4]), E.ptr(ys[r % 4]), 1, N, N, E.stream_ptr(dev)))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for r in range(40):
        E.check(E.lib().phb200_spmv(h, E.ptr(xs[r % 4]), E.ptr(ys[r % 4]), 1, N, N, E.stream_ptr(dev)))
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 40 * 1e3
    out[name] = {'spmv_us': us, 'algorithmic_bytes': nbytes, 'gbs': nbytes / us / 1e3, 'frac_of_peak': nbytes / us / 1e3 / 6552.3}
ya, yb = torch.empty((1, N), device=dev), torch.empty((1, N), device=dev)
E.check(E.lib().phb200_spmv(m.coded.handle, E.ptr(xs[0]), E.ptr(ya), 1, N, N, E.stream_ptr(dev)))
E.check(E.lib().phb200_spmv(m.handle, E.ptr(xs[0]), E.ptr(yb), 1, N, N, E.stream_ptr(dev)))
out['bit_identical'] = bool(torch.equal(ya, yb))
y = torch.randn((1, N), device=dev)
zero = torch.zeros(1, device=dev); mi = torch.full((1,), 2 ** 30, dtype=torch.int32, device=dev)
for name, mat in (('coded', m.coded), ('csr', m)):
    for tag, pre in (('jacobi_cg', pb.JacobiPreconditioner(m)), ('cg', None)):
        s = pb.DeviceSolver(mat, pre, E.CG, 0, 1)
        s.start(y, torch.zeros_like(y), zero, zero, mi)
        s.advance(20); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); s.advance(100); e1.record(); torch.cuda.synchronize()
        out[name][tag + '_it_per_s'] = 100 / (e0.elapsed_time(e1) * 1e-3)
        s.profile(True); s.advance(40); prof = s.profile_read(); s.profile(False)
        out[name][tag + '_kernels_us'] = [round(p[0] / max(p[1], 1) * 1e3, 1) for p in prof[:4]]
print(json.dumps(out))
