"""Which product routes reproduce the reference on the non-canonical natives (tests/golden/spmv_edges.npz)?   python scripts/probe_spmv_edges.py [name ...]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
import phiml_b200 as pb
from _cases import SPMV_EDGE_CASES, spmv_edge_case
from _util import golden
DEV = 'cuda:0'
g = golden('spmv_edges')
for name in (sys.argv[1:] or SPMV_EDGE_CASES):
    indptr, indices, data, shape, x = spmv_edge_case(name)
    ref = g[f"spmv_edge/{name}/csr"]
    xd = torch.as_tensor(x, device=DEV)
    nat = torch.sparse_csr_tensor(torch.as_tensor(indptr, device=DEV), torch.as_tensor(indices, device=DEV), torch.as_tensor(data, device=DEV), shape)
    m = pb.as_matrix(nat)
    twin = m.product_twin
    out = {'kind': m.kind, 'twin': twin.kind, 'stencil': m.stencil is not None}
    for label, fn in [('csr', lambda: m.matvec(xd, use_stencil=False)), ('twin', lambda: m.matvec(xd, use_stencil=True))]:
        y = fn().cpu().numpy()
        out[label] = bool(np.array_equal(y, ref, equal_nan=True))
    print(name, out)
