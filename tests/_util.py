"""Helpers shared by the CPU and GPU tests: golden access, oracle matrices / preconditioners for a case."""
import hashlib
import os

import numpy as np

from _cases import ADVDIFF, SOLVE_CASES, rhs_for, sample_index  # noqa: F401

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
_CACHE = {}


def golden(group: str):
    if group not in _CACHE:
        _CACHE[group] = np.load(os.path.join(GOLDEN_DIR, group + '.npz'))
    return _CACHE[group]


def sha(a) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def shifts_for(op: str, rank: int, dtype):
    from oracle.assemble import laplace_shifts, advdiff_shifts
    if op == 'poisson':
        return laplace_shifts(rank)
    return advdiff_shifts(rank, ADVDIFF['velocity'], ADVDIFF['c'], ADVDIFF['nu'], dtype)


def oracle_matrix(op: str, grid, dtype):
    from oracle.assemble import stencil_csr
    dtype = np.dtype(dtype).type
    return stencil_csr(grid, shifts_for(op, len(grid), dtype), dtype)


def oracle_pre(case, csr):
    from oracle import precond
    p = case.get('pre')
    if p is None:
        return None
    if p == 'jacobi':
        return precond.Jacobi(csr)
    if p == 'ilu':
        sweeps = precond.ilu_sweeps(csr.shape[0], csr.nnz)
        L, U = precond.ilu_fixed_point(csr, sweeps)
        return precond.IncompleteLU(L, U, f"iter={sweeps}")
    raise ValueError(p)


def case_inputs(name):
    case = SOLVE_CASES[name]
    csr = oracle_matrix(case['op'], case['grid'], case['dtype'])
    y = rhs_for(case, csr.shape[0])
    nb = y.shape[0]
    dt = y.dtype.type
    return case, csr, y, np.zeros_like(y), np.full(nb, case['rtol'], dt), np.full(nb, case['atol'], dt), np.full((1, nb), case.get('max_iter', 5000))


ADJOINT_CASES = {   # mirrors tests/golden/make_golden.py: name -> (grid, dtype, method, rtol, batch, seed, grad_for_f)
    'adj_a8x6x8_bicg2_f64': ((8, 6, 8), 'float64', 'biCG-stab(2)', 1e-10, 2, 5, True),
    'adj_a8x6x8_bicg2_f32': ((8, 6, 8), 'float32', 'biCG-stab(2)', 1e-5, 2, 5, True),
    'adj_a12x10_bicg1_f64': ((12, 10), 'float64', 'biCG-stab(1)', 1e-10, 3, 6, True),
    'adj_a12x8x4_bicg2_f64_tight': ((12, 8, 4), 'float64', 'biCG-stab(2)', 1e-12, 1, 7, True),
    'adj_a8x6x8_bicg2_f64_const': ((8, 6, 8), 'float64', 'biCG-stab(2)', 1e-10, 2, 5, False),
    'adj_a16x12_bicg2_f32_const': ((16, 12), 'float32', 'biCG-stab(2)', 1e-5, 2, 8, False),
}


def adjoint_golden(name):
    """Golden record of one backward pass: dict of arrays + the forward / backward matrices as SciPy CSR."""
    import scipy.sparse as sp
    g = golden('adjoint')
    key = f"adjoint/{name}/"
    rec = {k[len(key):]: g[k] for k in g.files if k.startswith(key)}
    grid = ADJOINT_CASES[name][0]
    n = int(np.prod(grid))
    mats = {}
    for tag in ('forward', 'backward'):
        layout = str(rec[f'{tag}_layout'])
        v = rec[f'{tag}_values']
        if layout == 'torch.sparse_coo':
            i = rec[f'{tag}_indices']
            m = sp.coo_matrix((v, (i[0], i[1])), shape=(n, n)).tocsr()
        elif layout == 'torch.sparse_csr':
            m = sp.csr_matrix((v, rec[f'{tag}_col'], rec[f'{tag}_crow']), shape=(n, n))
        else:
            m = sp.csc_matrix((v, rec[f'{tag}_row'], rec[f'{tag}_ccol']), shape=(n, n)).tocsr()
        m.sort_indices()
        mats[tag] = m
    return rec, mats
