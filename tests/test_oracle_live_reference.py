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


@pytest.mark.parametrize('seed', range(24))
def test_random_operators_through_the_tracer_bridge_equal_the_references_matrix(seed):
    """Random linear functions — sums of shifted copies of x (offsets up to +-2 along random axes, ZERO / PERIODIC / ZERO_GRADIENT padding) with scalar or
    per-cell coefficients — traced by the imported reference: the matrix the bridge builds from the tracer (phiml_b200.assembly.analyse_tracer + the
    oracle's assembly here, the device assembler on a GPU, pinned to the oracle byte for byte by the GPU suite) must equal tracer_to_coo + compress_rows:
    same type, dims, int32 pointers / indices, values, rank.  Operators the bridge does not take (it falls back) are counted, not failed."""
    import torch
    NUMPY, math = reference_backend()
    from phiml.math import spatial, extrapolation
    from phiml.math._lin_trace import matrix_from_function
    import phiml.math._lin_trace as lt
    from phiml_b200 import assembly
    from oracle.assemble import stencil_csr, shift_values_csr
    rng = np.random.default_rng(500 + seed)
    rank = int(rng.integers(1, 4))
    names = 'xyz'[:rank]
    grid = {d: int(rng.integers(4, 8)) for d in names}
    pad = [extrapolation.ZERO, extrapolation.PERIODIC, extrapolation.ZERO_GRADIENT][seed % 3]
    n_terms = int(rng.integers(2, 6))
    terms = []
    for _ in range(n_terms):
        d = names[int(rng.integers(0, rank))]
        off = int(rng.choice([-2, -1, 1, 2]))
        coef = float(rng.uniform(-1, 1))
        field = rng.uniform(0.5, 1.5, tuple(grid.values())) if rng.random() < 0.35 else None
        terms.append((d, off, coef, field))
    diag = float(rng.uniform(2, 4))

    def f(x):
        out = diag * x
        for d, off, coef, field in terms:
            shifted = math.shift(x, (off,), dims=d, padding=pad, stack_dim=None)[0]
            c = coef if field is None else math.wrap(field * coef, spatial(','.join(names)))
            out = out + c * shifted
        return out

    def oracle_assemble(g, offsets, coeffs, dtype):
        csr = stencil_csr(g, list(zip(offsets, coeffs)), dtype.type)
        return torch.as_tensor(csr.indptr), torch.as_tensor(csr.indices), torch.as_tensor(csr.data)

    def oracle_assemble_general(g, offsets, values, dtype):
        csr = shift_values_csr(g, offsets, values)
        return torch.as_tensor(csr.indptr), torch.as_tensor(csr.indices), torch.as_tensor(csr.data)

    original = lt.tracer_to_coo if not hasattr(lt.tracer_to_coo, '__wrapped_by_phb200__') else lt.tracer_to_coo
    from phiml_b200 import phiml_plugin
    if phiml_plugin._INSTALLED:
        original = phiml_plugin._INSTALLED['original_t2c']
    saved = lt.tracer_to_coo
    try:
        with math.precision(64 if seed % 2 else 32):
            x0 = math.zeros(spatial(**grid))
            lt.tracer_to_coo = original
            try:
                ref, ref_bias = matrix_from_function(f, x0, auto_compress=True)
            except AssertionError as exc:       # e.g. an offset of 2 wrapping onto itself on a 4-cell periodic axis: the reference cannot compress repeated entries
                pytest.skip(f"the reference does not build this operator: {str(exc)[:80]}")
            assembly.STATS.update(device_assemblies=0, general_assemblies=0, fallbacks=0)
            lt.tracer_to_coo = assembly.make_tracer_to_coo(original, lambda: True, oracle_assemble, oracle_assemble_general)
            mine, mine_bias = matrix_from_function(f, x0, auto_compress=True)
    finally:
        lt.tracer_to_coo = saved
    taken = assembly.STATS['device_assemblies'] + assembly.STATS['general_assemblies']
    assert taken + assembly.STATS['fallbacks'] == 1
    assert type(mine) is type(ref) and mine.shape == ref.shape
    if type(ref).__name__ == 'CompressedSparseMatrix':
        for name, dim in (('_indices', 'sp_entries'), ('_pointers', 'pointers'), ('_values', 'sp_entries')):
            a, b = np.asarray(getattr(mine, name).native(dim)), getattr(ref, name).numpy(dim)
            assert a.dtype == b.dtype and np.array_equal(a, b), (name, pad, terms)
        rank_of = lambda m: np.asarray(m._matrix_rank.numpy(m._matrix_rank.shape.names)) if hasattr(m._matrix_rank, 'numpy') else np.asarray(m._matrix_rank)
        assert np.array_equal(rank_of(mine), rank_of(ref))
    # what the matrix does, whichever way it was built
    probe = math.wrap(rng.standard_normal(tuple(grid.values())), spatial(','.join(names)))
    order = ','.join(names)
    assert np.allclose((mine @ probe).numpy(order), (ref @ probe).numpy(order), rtol=1e-5, atol=1e-5)
    assert np.allclose((mine @ probe).numpy(order), f(probe).numpy(order), rtol=1e-4, atol=1e-4)
    _TAKEN.append(taken)


_TAKEN = []


def test_the_bridge_took_most_random_operators():
    if not _TAKEN:
        pytest.skip("runs after the random-operator cases")
    print(f"the bridge took {sum(_TAKEN)} of {len(_TAKEN)} random operators")
    assert sum(_TAKEN) >= len(_TAKEN) * 0.6, f"the bridge took {sum(_TAKEN)} of {len(_TAKEN)} random operators"


@pytest.mark.parametrize('seed', range(16))
def test_random_ilu_sweeps_agree_with_the_imported_reference(seed):
    """incomplete_lu_coo (phiml/backend/_linalg.py:524-594) on random sparse patterns with a full diagonal — structurally unsymmetric, random sweep counts,
    `safe` on and off — against the oracle's sweeps: same entries in the same order, values to rounding; and the triangular solves of the factors against
    the reference's NumPy backend (SciPy spsolve_triangular)."""
    from oracle import precond
    NUMPY, math = reference_backend()
    from phiml.backend._linalg import incomplete_lu_coo
    rng = np.random.default_rng(7000 + seed)
    n = int(rng.integers(5, 40))
    dt = np.float64 if seed % 2 else np.float32
    a = sp.random(n, n, density=float(rng.uniform(0.05, 0.3)), random_state=np.random.RandomState(seed), format='lil')
    a.setdiag(0)
    a = a.tocsr()
    a = (a + sp.diags(np.asarray(abs(a).sum(axis=1)).ravel() + rng.uniform(0.5, 2.0, n))).tocsr().astype(dt)
    a.sort_indices()
    if sp.tril(a, -1).nnz == 0 or sp.triu(a, 1).nnz == 0:
        pytest.skip("triangular input: the reference has a branch of its own (upper) or raises (lower); pinned in tests/golden/ilu_edges.npz")
    sweeps, safe = int(rng.integers(1, 9)), bool(seed % 3 == 0)
    coo = a.tocoo()
    idx = np.stack([coo.row, coo.col], -1)[None].astype(np.int64)
    with math.precision(64 if dt == np.float64 else 32), np.errstate(all='ignore'):
        (li, lv), (ui, uv) = incomplete_lu_coo(idx, coo.data[None, :, None], a.shape, sweeps, safe)
    L, U = precond.ilu_fixed_point(a, sweeps, safe)
    tol = 1e-12 if dt == np.float64 else 2e-5
    for mine, i, v in ((L, np.asarray(li)[0], np.asarray(lv)[0, :, 0]), (U, np.asarray(ui)[0], np.asarray(uv)[0, :, 0])):
        ref = sp.csr_matrix((v, (i[:, 0], i[:, 1])), shape=a.shape)
        assert (abs(mine - ref) > tol * max(1.0, abs(ref).max())).nnz == 0
        pat_m = sp.csr_matrix((np.ones(mine.nnz), mine.indices, mine.indptr), shape=a.shape)
        pat_r = sp.csr_matrix((np.ones(len(v)), (i[:, 0], i[:, 1])), shape=a.shape)
        assert (pat_m != pat_r).nnz == 0
    rhs = rng.standard_normal((2, n)).astype(dt)
    with np.errstate(all='ignore'):
        ref_l = NUMPY.solve_triangular_sparse(L.astype(np.float64), rhs.astype(np.float64), True, True)
        ref_u = NUMPY.solve_triangular_sparse(U.astype(np.float64), rhs.astype(np.float64), False, False)
    assert np.allclose(precond.sptrsv(L.astype(np.float64), rhs.astype(np.float64), True, True), ref_l, rtol=1e-10, atol=1e-10)
    assert np.allclose(precond.sptrsv(U.astype(np.float64), rhs.astype(np.float64), False, False), ref_u, rtol=1e-9, atol=1e-9)


def test_exotic_linear_functions_through_the_bridge_behave_like_the_reference():
    """Operators outside the plain stencils: biharmonic, grad-div, different paddings per axis and per side, non-zero constant extrapolation and affine
    functions (bias), upwind shifts, slicing + padding, per-cell factors, 1-D and 4-D grids, down-sampling, identity / zero functions, and functions the
    reference's tracer itself rejects.  With the bridge switched on every one gives the reference's matrix type, the same `M @ x + bias`, and raises
    where the reference raises; what the bridge cannot express falls back to the reference's own assembly."""
    import torch
    NUMPY, math = reference_backend()
    from phiml.math import spatial, channel, extrapolation as E_
    from phiml.math._lin_trace import matrix_from_function
    import phiml.math._lin_trace as lt
    from phiml_b200 import assembly, phiml_plugin
    from oracle.assemble import stencil_csr, shift_values_csr

    def oa(g, offsets, coeffs, dtype):
        csr = stencil_csr(g, list(zip(offsets, coeffs)), dtype.type)
        return torch.as_tensor(csr.indptr), torch.as_tensor(csr.indices), torch.as_tensor(csr.data)

    def oag(g, offsets, values, dtype):
        csr = shift_values_csr(g, offsets, values)
        return torch.as_tensor(csr.indptr), torch.as_tensor(csr.indices), torch.as_tensor(csr.data)

    ops = {
        'biharmonic': (lambda x: math.laplace(math.laplace(x, padding=E_.ZERO), padding=E_.ZERO), dict(x=7, y=6)),
        'grad_div': (lambda x: math.sum(math.spatial_gradient(math.sum(math.spatial_gradient(x, padding=E_.ZERO, stack_dim=channel('g')), 'g'), padding=E_.ZERO, stack_dim=channel('h')), 'h'), dict(x=6, y=5)),
        'mixed_axes': (lambda x: -math.laplace(x, padding=E_.combine_sides(x=E_.PERIODIC, y=E_.ZERO_GRADIENT)) + x, dict(x=6, y=5)),
        'sides_differ': (lambda x: -math.laplace(x, padding=E_.combine_sides(x=(E_.ZERO, E_.ZERO_GRADIENT), y=E_.PERIODIC)) + x, dict(x=6, y=5)),
        'one_extrapolation': (lambda x: -math.laplace(x, padding=E_.ONE) + x, dict(x=6, y=5)),
        'constant_3': (lambda x: -math.laplace(x, padding=E_.ConstantExtrapolation(3.0)) + x, dict(x=6, y=5)),
        'symmetric': (lambda x: -math.laplace(x, padding=E_.SYMMETRIC) + x, dict(x=6, y=5)),
        'upwind': (lambda x: x - 0.3 * (x - math.shift(x, (-1,), dims='x', padding=E_.ZERO, stack_dim=None)[0]), dict(x=6, y=5)),
        'affine': (lambda x: -math.laplace(x, padding=E_.ZERO) + 1.5, dict(x=6, y=5)),
        'slice_pad': (lambda x: math.pad(x[{'x': slice(1, None)}], {'x': (0, 1)}, E_.ZERO) - x, dict(x=6, y=5)),
        'mean_removed': (lambda x: x - math.mean(x, 'x,y'), dict(x=5, y=4)),
        'per_cell_factor': (lambda x: x * math.wrap(np.arange(20.).reshape(5, 4), spatial('x,y')), dict(x=5, y=4)),
        'one_d': (lambda x: -math.laplace(x, padding=E_.ZERO), dict(x=9)),
        'four_d': (lambda x: -math.laplace(x, padding=E_.ZERO), dict(x=3, y=4, z=3, w=3)),
        'downsample': (lambda x: x[{'x': slice(0, None, 2)}], dict(x=6, y=5)),
        'identity': (lambda x: x, dict(x=6, y=5)),
        'zero_fn': (lambda x: 0 * x, dict(x=6, y=5)),
    }
    original = phiml_plugin._INSTALLED['original_t2c'] if phiml_plugin._INSTALLED else lt.tracer_to_coo
    saved = lt.tracer_to_coo
    bridge = assembly.make_tracer_to_coo(original, lambda: True, oa, oag)
    took = 0
    try:
        for name, (f, shape) in ops.items():
            out = []
            for t2c in (original, bridge):
                lt.tracer_to_coo = t2c
                assembly.STATS.update(device_assemblies=0, general_assemblies=0, fallbacks=0)
                try:
                    with math.precision(64):
                        x0 = math.zeros(spatial(**shape))
                        m, bias = matrix_from_function(f, x0, auto_compress=True)
                        probe = math.wrap(np.random.default_rng(1).standard_normal(tuple(shape.values())), spatial(','.join(shape)))
                        y = m @ probe + bias
                        out.append(('ok', type(m).__name__, np.asarray(y.numpy(y.shape.names))))
                except Exception as exc:
                    out.append(('raised', type(exc).__name__, None))
            (s0, t0, a), (s1, t1, b) = out
            assert (s0, t0) == (s1, t1), f"{name}: the reference {s0} {t0}, with the bridge {s1} {t1}"
            if s0 == 'ok':
                assert a.shape == b.shape and np.allclose(a, b, rtol=1e-12, atol=1e-12), name
            took += assembly.STATS['device_assemblies'] + assembly.STATS['general_assemblies']
    finally:
        lt.tracer_to_coo = saved
    assert took >= 9, took
