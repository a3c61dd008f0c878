"""
CPU suite: the oracle against the IMPORTED reference on randomly drawn cases (where /root/reference exists; the committed fixtures under tests/golden/ pin the
same functions on fixed cases everywhere else).  Random grids (1-3 D), random constant-coefficient operators (symmetric positive definite for CG, shifted
nonsymmetric for BiCGstab), random batches, tolerances and iteration limits; the reference's NumPy backend and the restatement must agree on every
counter and flag exactly and on x to rounding — both compute in NumPy, in the same order.
"""
import os
import sys

import numpy as np
import pytest
import scipy.sparse as sp

REF = '/root/reference'
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, 'phiml')), reason="the reference is not on this machine")


def reference_backend():
    if REF not in sys.path and not any(os.path.isdir(os.path.join(p, 'phiml')) for p in sys.path if p):
        sys.path.insert(0, REF)
    from phiml.backend import NUMPY
    from phiml import math
    return NUMPY, math


def random_operator(rng, symmetric: bool):
    rank = int(rng.integers(1, 4))
    grid = tuple(int(rng.integers(3, 8)) for _ in range(rank))
    n = int(np.prod(grid))
    idx = np.arange(n).reshape(grid)
    rows, cols, vals = [], [], []
    diag = 0.0
    for ax in range(rank):
        lo, hi = rng.uniform(0.3, 1.5, 2)
        if symmetric:
            hi = lo
        a = [slice(None)] * rank
        b = [slice(None)] * rank
        a[ax], b[ax] = slice(0, -1), slice(1, None)
        i, j = idx[tuple(a)].ravel(), idx[tuple(b)].ravel()
        rows += [i, j]
        cols += [j, i]
        vals += [np.full(i.size, -hi), np.full(i.size, -lo)]
        diag += lo + hi
    rows.append(np.arange(n)); cols.append(np.arange(n)); vals.append(np.full(n, diag + rng.uniform(0.05, 0.5)))
    m = sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(n, n)).tocsr()
    m.sort_indices()
    return m


@pytest.mark.parametrize('seed', range(40))
def test_random_solves_agree_with_the_imported_reference(seed):
    from oracle import krylov, precond
    NUMPY, math = reference_backend()
    rng = np.random.default_rng(1000 + seed)
    method = ['CG', 'CG-adaptive', 'biCG-stab(1)', 'biCG-stab(2)', 'biCG-stab(3)'][seed % 5]
    dt = np.float64 if rng.random() < 0.5 else np.float32
    csr = random_operator(rng, symmetric=method.startswith('CG')).astype(dt)
    nb = int(rng.integers(1, 4))
    y = rng.standard_normal((nb, csr.shape[0])).astype(dt)
    if rng.random() < 0.2:
        y[int(rng.integers(0, nb))] = 0
    x0 = (rng.standard_normal(y.shape) * (rng.random() < 0.3)).astype(dt)
    rtol = np.full(nb, 10.0 ** rng.uniform(-10 if dt == np.float64 else -5, -2), dt)
    atol = np.full(nb, 0.0 if rng.random() < 0.7 else 10.0 ** rng.uniform(-8, -3), dt)
    max_iter = np.full((1, nb), int(rng.choice([0, 1, 3, 50, 1000])))
    use_jacobi = method in ('CG', 'biCG-stab(1)') and rng.random() < 0.4
    pre_o = precond.Jacobi(csr) if use_jacobi else None

    class RefJacobi:
        def __init__(self, m): self.inv = (1 / m.diagonal()).astype(m.dtype)
        def apply(self, v): return v * self.inv
        def apply_inv_l(self, v): return v * self.inv
        def apply_inv_u(self, v): return v
        def __repr__(self): return "jacobi"
    pre_r = RefJacobi(csr) if use_jacobi else None
    with math.precision(64 if dt == np.float64 else 32), np.errstate(all='ignore'):
        with NUMPY:
            ref = NUMPY.linear_solve(method, csr, y, x0.copy(), rtol, atol, max_iter, pre_r, None)
        mine = krylov.linear_solve(method, csr, y, x0.copy(), rtol, atol, max_iter, pre_o, flavour='numpy')
    np.testing.assert_array_equal(mine.iterations, np.asarray(ref.iterations))
    np.testing.assert_array_equal(mine.function_evaluations, np.asarray(ref.function_evaluations))
    np.testing.assert_array_equal(mine.converged, np.asarray(ref.converged))
    np.testing.assert_array_equal(mine.diverged, np.asarray(ref.diverged))
    rx = np.asarray(ref.x)
    np.testing.assert_array_equal(np.isfinite(mine.x), np.isfinite(rx))
    ok = np.isfinite(rx)
    scale = max(float(np.abs(np.where(ok, rx, 0)).max()), 1e-30)
    assert float(np.abs(np.where(ok, mine.x - rx, 0)).max()) <= (1e-12 if dt == np.float64 else 1e-5) * scale
    assert mine.method == ref.method
