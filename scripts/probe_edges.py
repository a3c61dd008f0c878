"""Prints device vs reference counters / flags for the edge cases (tests/golden/edges.npz).   python scripts/probe_edges.py [substring ...]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
import phiml_b200 as pb
from _cases import EDGE_CASES, edge_case_inputs
from _util import golden, oracle_matrix

DEV = 'cuda:0'
g = golden('edges')
want = sys.argv[1:]
for name, c in EDGE_CASES.items():
    if want and not any(w in name for w in want):
        continue
    csr = oracle_matrix(c['op'], c['grid'], c['dtype'])
    y, x0, rtol, atol, max_iter = edge_case_inputs(name, csr.shape[0])
    for use_stencil in (True, False):
        m = pb.as_matrix(torch.sparse_csr_tensor(torch.as_tensor(csr.indptr, device=DEV), torch.as_tensor(csr.indices, device=DEV), torch.as_tensor(csr.data, device=DEV), csr.shape))
        if use_stencil and m.stencil is None:
            continue
        if not use_stencil:
            m.stencil = None
        yd, x0d = torch.as_tensor(y, device=DEV), torch.as_tensor(x0, device=DEV)
        pre = pb.compute_preconditioner(c['pre'], m)
        if c['backend'] == 'numpy' and c['method'] == 'CG':
            res = pb.generic_conjugate_gradient(m, yd, x0d, rtol, atol, max_iter, pre)
        elif c['backend'] == 'numpy' and c['method'] == 'CG-adaptive':
            res = pb.generic_conjugate_gradient_adaptive(m, yd, x0d, rtol, atol, max_iter)
        else:
            res = pb.linear_solve(c['method'], m, yd, x0d, rtol, atol, max_iter, pre, None)
        key = f"edge/{name}"
        dev = [res.iterations.tolist(), res.function_evaluations.tolist(), res.converged.tolist(), res.diverged.tolist()]
        ref = [g[key + '/iterations'].tolist(), g[key + '/function_evaluations'].tolist(), g[key + '/converged'].tolist(), g[key + '/diverged'].tolist()]
        x, rx = res.x.cpu().numpy(), g[key + '/x']
        fin = [np.isfinite(x).all(axis=1).tolist(), np.isfinite(rx).all(axis=1).tolist()]
        flag = 'OK  ' if dev == ref and fin[0] == fin[1] else 'DIFF'
        print(flag, name, 'stencil' if use_stencil else 'csr', '\n     dev', dev, fin[0], '\n     ref', ref, fin[1])
