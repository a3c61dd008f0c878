"""
CPU suite: pins oracle/adjoint.py (implicit-gradient step of math.solve_linear) to the backward passes recorded
from the real reference (tests/golden/adjoint.npz, made by tests/golden/make_golden.py adjoint).
"""
import numpy as np
import pytest

from _util import ADJOINT_CASES, ADVDIFF, adjoint_golden
from oracle import adjoint
from oracle.assemble import advdiff_shifts, stencil_csr


@pytest.mark.parametrize('name', sorted(ADJOINT_CASES))
def test_backward_matrix_is_the_transpose(name):
    grid, dt, *_ = ADJOINT_CASES[name]
    rec, mats = adjoint_golden(name)
    t = adjoint.transpose(mats['forward'])
    b = mats['backward']
    np.testing.assert_array_equal(t.indptr, b.indptr)
    np.testing.assert_array_equal(t.indices, b.indices)
    np.testing.assert_array_equal(t.data, b.data)
    # and the forward matrix is the oracle's own assembly of the operator (k = 1)
    # with grad_for_f the values are traced, so the reference cannot drop the boundary entries: they are stored as explicit zeros
    # at the wrapped positions (600 instead of 556 entries at 12x10)
    own = stencil_csr(grid, advdiff_shifts(len(grid), ADVDIFF['velocity'], ADVDIFF['c'], ADVDIFF['nu'], np.dtype(dt).type), np.dtype(dt).type)
    fwd = mats['forward'].copy()
    fwd.eliminate_zeros()
    assert (fwd.nnz < mats['forward'].nnz) == ADJOINT_CASES[name][6]
    np.testing.assert_array_equal(own.indices, fwd.indices)
    np.testing.assert_allclose(own.data, fwd.data, rtol=2e-7 if dt == 'float32' else 1e-15)


@pytest.mark.parametrize('name', sorted(ADJOINT_CASES))
def test_implicit_gradient_matches_reference(name):
    grid, dt, method, rtol, nb, seed, grad_f = ADJOINT_CASES[name]
    rec, mats = adjoint_golden(name)
    x = rec['x']
    dx = rec['backward_rhs']
    np.testing.assert_array_equal(dx, x)                     # l2 loss: dL/dx = x
    entries = (rec['dm_row'], rec['dm_col']) if grad_f else None
    tol = np.full(nb, rtol, dtype=dt)
    res, dm = adjoint.implicit_gradient_solve(method, mats['forward'], x, dx, tol, np.zeros(nb, dtype=dt), np.full((1, nb), 1000), None, 'torch', entries)
    assert np.max(np.abs(res.iterations.astype(int) - rec['backward_iterations'].astype(int))) <= 1
    scale = np.abs(rec['dy']).max()
    np.testing.assert_allclose(res.x, rec['dy'], rtol=0, atol=(2e-4 if dt == "float32" else 1e-8) * scale)
    if grad_f:
        # the golden dm is built from the reference's own dy: restating it from that dy must agree to rounding
        exact = adjoint.matrix_gradient(rec['dm_row'], rec['dm_col'], rec['dy'], x)
        np.testing.assert_allclose(exact, rec['dm'], rtol=0, atol=(1e-5 if dt == 'float32' else 1e-13) * np.abs(rec['dm']).max())
        np.testing.assert_allclose(dm, rec['dm'], rtol=0, atol=(2e-4 if dt == "float32" else 1e-8) * np.abs(rec['dm']).max())


def test_csr_rows_expand_pointers():
    rec, mats = adjoint_golden('adj_a8x6x8_bicg2_f64_const')
    a = mats['forward']
    rows = adjoint.csr_rows(a)
    assert rows.shape == (a.nnz,) and np.all(np.diff(rows) >= 0)
    np.testing.assert_array_equal(np.bincount(rows, minlength=a.shape[0]), np.diff(a.indptr))


@pytest.mark.parametrize('name', sorted(n for n, c in ADJOINT_CASES.items() if c[6]))
def test_parameter_gradient_follows_from_matrix_gradient(name):
    """dL/dk of A(k) = I + c u.grad - k nu laplace, as the reference's autograd reports it, equals sum_k dm[k] dA[k]/dk with the
    oracle's dm — pins sign, entry order and batch summation of the matrix gradient end to end."""
    from oracle.assemble import laplace_shifts
    grid, dt, method, rtol, nb, seed, _ = ADJOINT_CASES[name]
    rec, mats = adjoint_golden(name)
    lap = stencil_csr(grid, laplace_shifts(len(grid), scale=1.0), np.float64)        # +laplace with ZERO padding
    d_entries = -ADVDIFF['nu'] * np.asarray(lap[rec['dm_row'], rec['dm_col']]).ravel()
    dm = adjoint.matrix_gradient(rec['dm_row'], rec['dm_col'], rec['dy'], rec['x'])
    np.testing.assert_allclose(float(np.sum(dm.astype(np.float64) * d_entries)), float(rec['dk']), rtol=2e-5 if dt == 'float32' else 1e-10)
