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
