"""
GPU suite, implicit-gradient step: adjoint solve on the transposed native + matrix gradient kernel, against the
backward passes recorded from the real reference (tests/golden/adjoint.npz) and the oracle (oracle/adjoint.py).
Tolerances: dm from given dy / x bit-exact (same products, same order); dy within 1e-4 (fp32, two iterative solves stopped at
1e-5) / 1e-8 (fp64, stopped at 1e-10) relative plus a true-residual check of the adjoint system; iterations within +-2.
"""
import numpy as np
import pytest
import torch

from _util import ADJOINT_CASES, ADVDIFF, adjoint_golden

pytestmark = pytest.mark.gpu

import phiml_b200 as pb  # noqa: E402
from phiml_b200 import adjoint  # noqa: E402
from phiml_b200.sparse import plain_csr  # noqa: E402

DEV = 'cuda:0'


def native_from_golden(rec, tag, n):
    layout = str(rec[f'{tag}_layout'])
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a), device=DEV)
    v = t(rec[f'{tag}_values'])
    if layout == 'torch.sparse_coo':
        return torch.sparse_coo_tensor(t(rec[f'{tag}_indices']), v, (n, n))
    if layout == 'torch.sparse_csr':
        return torch.sparse_csr_tensor(t(rec[f'{tag}_crow']), t(rec[f'{tag}_col']), v, (n, n))
    return torch.sparse_csc_tensor(t(rec[f'{tag}_ccol']), t(rec[f'{tag}_row']), v, (n, n))


@pytest.mark.parametrize('dt', ['float32', 'float64'])
@pytest.mark.parametrize('nb', [1, 3])
def test_matrix_gradient_kernels_bit_exact(dt, nb):
    from oracle import adjoint as ora
    from oracle.assemble import advdiff_shifts, stencil_csr
    grid = (9, 7, 6)
    a = stencil_csr(grid, advdiff_shifts(3, ADVDIFF['velocity'], ADVDIFF['c'], ADVDIFF['nu'], np.dtype(dt).type), np.dtype(dt).type)
    n = a.shape[0]
    rng = np.random.default_rng(21)
    dy, x = rng.standard_normal((nb, n)).astype(dt), rng.standard_normal((nb, n)).astype(dt)
    rows = ora.csr_rows(a)
    ref = ora.matrix_gradient(rows, a.indices.astype(np.int64), dy, x)
    dyd, xd = torch.as_tensor(dy, device=DEV), torch.as_tensor(x, device=DEV)
    csr = torch.sparse_csr_tensor(torch.as_tensor(a.indptr, device=DEV), torch.as_tensor(a.indices, device=DEV), torch.as_tensor(a.data, device=DEV), a.shape)
    np.testing.assert_array_equal(adjoint.matrix_gradient(csr, dyd, xd).cpu().numpy(), ref)
    # COO in a shuffled entry order: dm follows the order of the native
    perm = rng.permutation(a.nnz)
    coo = torch.sparse_coo_tensor(torch.as_tensor(np.stack([rows[perm], a.indices[perm].astype(np.int64)]), device=DEV), torch.as_tensor(a.data[perm], device=DEV), a.shape)
    np.testing.assert_array_equal(adjoint.matrix_gradient(coo, dyd, xd).cpu().numpy(), ref[perm])
    # a matrix with empty rows, rows longer than one CTA chunk and more rows than one chunk
    import scipy.sparse as sp
    b = sp.random(700, 500, density=0.02, random_state=rng, format='csr', dtype=np.dtype(dt))
    dense_row = sp.csr_matrix(np.ones((1, 500), dtype=dt))
    b = sp.vstack([b[:300], dense_row, sp.csr_matrix((3, 500), dtype=dt), b[300:]]).tocsr()
    b.sort_indices()
    dy2, x2 = rng.standard_normal((nb, b.shape[0])).astype(dt), rng.standard_normal((nb, b.shape[1])).astype(dt)
    ref2 = ora.matrix_gradient(ora.csr_rows(b), b.indices.astype(np.int64), dy2, x2)
    nat = torch.sparse_csr_tensor(torch.as_tensor(b.indptr, device=DEV), torch.as_tensor(b.indices, device=DEV), torch.as_tensor(b.data, device=DEV), b.shape)
    np.testing.assert_array_equal(adjoint.matrix_gradient(nat, torch.as_tensor(dy2, device=DEV), torch.as_tensor(x2, device=DEV)).cpu().numpy(), ref2)
    # no stored entries
    empty = torch.sparse_csr_tensor(torch.zeros(6, dtype=torch.int32, device=DEV), torch.zeros(0, dtype=torch.int32, device=DEV), torch.zeros(0, dtype=getattr(torch, dt), device=DEV), (5, 5))
    assert adjoint.matrix_gradient(empty, torch.ones(1, 5, dtype=getattr(torch, dt), device=DEV), torch.ones(1, 5, dtype=getattr(torch, dt), device=DEV)).numel() == 0


@pytest.mark.parametrize('name', sorted(ADJOINT_CASES))
def test_implicit_gradient_matches_reference_backward_pass(name):
    grid, dt, method, rtol, nb, seed, grad_f = ADJOINT_CASES[name]
    rec, mats = adjoint_golden(name)
    n = int(np.prod(grid))
    fwd = native_from_golden(rec, 'forward', n)
    x = torch.as_tensor(rec['x'], device=DEV)
    dx = torch.as_tensor(rec['backward_rhs'], device=DEV)
    # the transposed native is the reference's: same arrays under the other layout
    t = adjoint.transpose(fwd)
    assert str(t.layout) == str(rec['backward_layout'])
    if not grad_f:
        assert t.values().data_ptr() == fwd.values().data_ptr() and t.ccol_indices().data_ptr() == fwd.crow_indices().data_ptr()
        np.testing.assert_array_equal(t.row_indices().cpu().numpy(), rec['backward_row'])
        np.testing.assert_array_equal(t.ccol_indices().cpu().numpy(), rec['backward_ccol'])
    else:
        np.testing.assert_array_equal(t._indices().cpu().numpy(), rec['backward_indices'])
    tol = np.full(nb, rtol)
    res, dm = adjoint.implicit_gradient_solve(method, fwd, x, dx, tol, np.zeros(nb), np.full((1, nb), 1000), None, matrix_adjoint=grad_f)
    assert bool(res.converged.all())
    assert np.max(np.abs(res.iterations.cpu().numpy().astype(int) - rec['backward_iterations'].astype(int))) <= 2
    dy = res.x.cpu().numpy()
    scale = np.abs(rec['dy']).max()
    np.testing.assert_allclose(dy, rec['dy'], rtol=0, atol=(1e-4 if dt == 'float32' else 1e-8) * scale)   # fp32: two solves stopped at 1e-5
    # true residual of the adjoint system
    r = mats['forward'].T.dot(dy.T).T - rec['backward_rhs']
    assert np.max(np.linalg.norm(r, axis=1) / np.linalg.norm(rec['backward_rhs'], axis=1)) < (5e-5 if dt == 'float32' else 1e-9)
    if grad_f:
        np.testing.assert_allclose(dm.cpu().numpy(), rec['dm'], rtol=0, atol=(1e-4 if dt == 'float32' else 1e-8) * np.abs(rec['dm']).max())
        # dL/dk through the matrix gradient = the reference's autograd value
        from oracle.assemble import laplace_shifts, stencil_csr
        lap = stencil_csr(grid, laplace_shifts(len(grid), scale=1.0), np.float64)
        d_entries = -ADVDIFF['nu'] * np.asarray(lap[rec['dm_row'], rec['dm_col']]).ravel()
        np.testing.assert_allclose(float(np.sum(dm.cpu().numpy().astype(np.float64) * d_entries)), float(rec['dk']), rtol=1e-4 if dt == 'float32' else 1e-8)


def test_transposed_natives_reach_the_stencil_kernels():
    """CSC natives of the backward pass (arrays of A, read as A^T) and COO natives with the stored boundary zeros of
    grad_for_f are recognised as stencils, so adjoint solves run on the same kernels as forward solves."""
    rec, mats = adjoint_golden('adj_a8x6x8_bicg2_f64_const')
    fwd = native_from_golden(rec, 'forward', 384)
    t = plain_csr(pb.as_matrix(adjoint.transpose(fwd)))
    assert t.stencil is not None and t.stencil.grid == (8, 6, 8)
    x = torch.as_tensor(np.random.default_rng(3).standard_normal((2, 384)), device=DEV)
    np.testing.assert_array_equal(t.matvec(x, use_stencil=True).cpu().numpy(), t.matvec(x, use_stencil=False).cpu().numpy())
    np.testing.assert_allclose(t.matvec(x).cpu().numpy(), mats['backward'].dot(x.cpu().numpy().T).T, rtol=1e-13, atol=1e-13)
    rec, mats = adjoint_golden('adj_a8x6x8_bicg2_f64')
    assert (rec['forward_values'] == 0).any()
    for tag in ('forward', 'backward'):
        m = pb.as_matrix(native_from_golden(rec, tag, 384))
        assert m.stencil is not None, tag
        assert m.nnz() == rec[f'{tag}_values'].size       # the CSR keeps the reference's pattern (ILU is defined on it)
        np.testing.assert_allclose(m.matvec(x).cpu().numpy(), mats[tag].dot(x.cpu().numpy().T).T, rtol=1e-13, atol=1e-13)
        np.testing.assert_allclose(m.matvec(x, use_stencil=False).cpu().numpy(), mats[tag].dot(x.cpu().numpy().T).T, rtol=1e-13, atol=1e-13)


def test_stencil_handle_transposes_by_mirroring_offsets():
    from oracle.assemble import advdiff_shifts, stencil_csr
    grid = (6, 5, 8)
    shifts = advdiff_shifts(3, ADVDIFF['velocity'], ADVDIFF['c'], ADVDIFF['nu'], np.float64)
    st = pb.Matrix.stencil_matrix(grid, [s for s, _ in shifts], [c for _, c in shifts], torch.float64, DEV)
    a = stencil_csr(grid, shifts, np.float64)
    x = np.random.default_rng(4).standard_normal((2, a.shape[0]))
    y = adjoint.transpose(st).matvec(torch.as_tensor(x, device=DEV)).cpu().numpy()
    np.testing.assert_allclose(y, a.T.tocsr().dot(x.T).T, rtol=1e-13, atol=1e-13)


@pytest.mark.parametrize('layout', ['coo', 'csr'])
def test_autograd_function_against_dense_autograd(layout):
    """SolveLinear: gradients w.r.t. the right-hand side and the stored matrix values equal torch's dense autograd."""
    from oracle.assemble import advdiff_shifts, stencil_csr
    grid = (5, 4, 4)
    a = stencil_csr(grid, advdiff_shifts(3, ADVDIFF['velocity'], ADVDIFF['c'], ADVDIFF['nu'], np.float64), np.float64)
    n = a.shape[0]
    rows = np.repeat(np.arange(n), np.diff(a.indptr))
    rng = np.random.default_rng(9)
    y0 = torch.as_tensor(rng.standard_normal((2, n)), device=DEV)
    w = torch.as_tensor(rng.standard_normal((2, n)), device=DEV)
    vals0 = torch.as_tensor(a.data, device=DEV)
    ri, ci = torch.as_tensor(rows, device=DEV), torch.as_tensor(a.indices.astype(np.int64), device=DEV)
    # dense reference
    vd, yd = vals0.clone().requires_grad_(True), y0.clone().requires_grad_(True)
    dense = torch.zeros((n, n), dtype=torch.float64, device=DEV).index_put((ri, ci), vd)
    xd = torch.linalg.solve(dense, yd.T).T
    (0.5 * (xd * xd).sum() + (w * xd).sum()).backward()
    # phb200
    v, y = vals0.clone().requires_grad_(True), y0.clone().requires_grad_(True)
    if layout == 'coo':
        native = torch.sparse_coo_tensor(torch.stack([ri, ci]), vals0, (n, n))
    else:
        native = torch.sparse_csr_tensor(torch.as_tensor(a.indptr, device=DEV), torch.as_tensor(a.indices, device=DEV), vals0, (n, n))
    x = adjoint.SolveLinear.apply(v, y, native, 'biCG-stab(2)', 1e-12, 0.0, 1000)
    (0.5 * (x * x).sum() + (w * x).sum()).backward()
    np.testing.assert_allclose(x.detach().cpu().numpy(), xd.detach().cpu().numpy(), rtol=0, atol=1e-10)
    np.testing.assert_allclose(y.grad.cpu().numpy(), yd.grad.cpu().numpy(), rtol=0, atol=1e-9 * float(yd.grad.abs().max()))
    np.testing.assert_allclose(v.grad.cpu().numpy(), vd.grad.cpu().numpy(), rtol=0, atol=1e-9 * float(vd.grad.abs().max()))


@pytest.mark.parametrize('dt', ['float32', 'float64'])
def test_native_transposition_is_bit_exact(dt):
    """phb200_csr_transpose (used once per CSC native of a backward pass): indices and values equal SciPy's, rows sorted."""
    import ctypes
    import scipy.sparse as sp
    from phiml_b200 import _engine as E
    rng = np.random.default_rng(31)
    for rows, cols, density in [(700, 500, 0.02), (64, 3000, 0.01), (1, 17, 0.5), (300, 300, 0.0)]:
        a = sp.random(rows, cols, density=density, random_state=rng, format='csr', dtype=np.dtype(dt))
        a.sort_indices()
        ref = a.T.tocsr()
        ref.sort_indices()
        t = lambda v: torch.as_tensor(np.ascontiguousarray(v), device=DEV)
        rp, ci, va = t(a.indptr.astype(np.int32)), t(a.indices.astype(np.int32)), t(a.data)
        o_rp = torch.empty(cols + 1, dtype=torch.int32, device=DEV)
        o_ci = torch.empty(a.nnz, dtype=torch.int32, device=DEV)
        o_va = torch.empty(a.nnz, dtype=va.dtype, device=DEV)
        E.check(E.lib().phb200_csr_transpose(E.ptr(rp), E.ptr(ci), E.ptr(va), rows, cols, a.nnz, E.dtype_code(va.dtype), E.ptr(o_rp), E.ptr(o_ci), E.ptr(o_va), E.stream_ptr(DEV)))
        np.testing.assert_array_equal(o_rp.cpu().numpy(), ref.indptr)
        np.testing.assert_array_equal(o_ci.cpu().numpy(), ref.indices)
        np.testing.assert_array_equal(o_va.cpu().numpy(), ref.data)
    # through the host path: CSC native -> plain CSR
    a = sp.random(400, 400, density=0.03, random_state=rng, format='csc', dtype=np.dtype(dt))
    a.sort_indices()
    nat = torch.sparse_csc_tensor(torch.as_tensor(a.indptr, device=DEV), torch.as_tensor(a.indices, device=DEV), torch.as_tensor(a.data, device=DEV), a.shape)
    m = plain_csr(pb.as_matrix(nat))
    ref = a.tocsr()
    ref.sort_indices()
    np.testing.assert_array_equal(m.rowptr.cpu().numpy(), ref.indptr)
    np.testing.assert_array_equal(m.col.cpu().numpy(), ref.indices)
    np.testing.assert_array_equal(m.values.cpu().numpy(), ref.data)
