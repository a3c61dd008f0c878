"""
Case table shared by tests/golden/make_golden.py (runs the real reference) and the parity tests
(run the oracle and the CUDA path on the same seeded inputs).  Inputs follow SURVEY.md §6/§8(d):
``y = default_rng(seed).standard_normal``, ``x0 = 0``, ZERO padding.
"""
import numpy as np

ADVDIFF = dict(velocity=(1.0, 0.5, 0.25), c=0.3, nu=0.2)


def sample_index(n: int) -> np.ndarray:
    return np.unique(np.linspace(0, n - 1, 257).astype(np.int64))


def rhs_for(case: dict, n: int) -> np.ndarray:
    rng = np.random.default_rng(case.get('seed', 0))
    nb = case.get('batch', 1)
    y = rng.standard_normal((nb, n))
    scale = case.get('rhs_scale')
    if scale is not None:          # batched-tolerance quirk probe: same vector at different magnitudes
        y = y[:1] * np.asarray(scale)[:, None]
    if case.get('zero_rhs_row') is not None:
        y[case['zero_rhs_row']] = 0
    return y.astype(case['dtype'])


def _c(op, grid, dtype, method, rtol, atol=0.0, **kw):
    return dict(op=op, grid=grid, dtype=dtype, method=method, rtol=rtol, atol=atol, **kw)


SOLVE_CASES = {
    # --- BASELINE config 1: 2-D 128² ---
    'p128_cg_f64_default': _c('poisson', (128, 128), 'float64', 'CG', 1e-12, 1e-12),
    'p128_cg_f64_1e10': _c('poisson', (128, 128), 'float64', 'CG', 1e-10),
    'p128_cg_f32_default': _c('poisson', (128, 128), 'float32', 'CG', 1e-5, 1e-5),
    'p128_cg_f64_torch': _c('poisson', (128, 128), 'float64', 'CG', 1e-12, 1e-12, backend='torch'),
    'p128_cg_f32_torch': _c('poisson', (128, 128), 'float32', 'CG', 1e-5, 1e-5, backend='torch'),
    'p128_auto_f64': _c('poisson', (128, 128), 'float64', 'auto', 1e-10),
    'p128_auto_f32_torch': _c('poisson', (128, 128), 'float32', 'auto', 1e-5, backend='torch'),
    'p128_cgadaptive_f32': _c('poisson', (128, 128), 'float32', 'CG-adaptive', 1e-5),
    'p128_bicg2_f64': _c('poisson', (128, 128), 'float64', 'biCG-stab(2)', 1e-10),
    'p128_bicg2_f64_torch': _c('poisson', (128, 128), 'float64', 'biCG-stab(2)', 1e-10, backend='torch'),   # the reference's own spread on this chaotic case
    'p128_jacobi_f64': _c('poisson', (128, 128), 'float64', 'CG', 1e-10, pre='jacobi'),
    'p128_ilu_f64': _c('poisson', (128, 128), 'float64', 'CG', 1e-10, pre='ilu'),
    'p128_ilu_f32': _c('poisson', (128, 128), 'float32', 'CG', 1e-5, pre='ilu'),
    # --- small cases with full x stored ---
    'p16_cg_f64': _c('poisson', (16, 16), 'float64', 'CG', 1e-10, batch=2, seed=1),
    'p16_cg_f32': _c('poisson', (16, 16), 'float32', 'CG', 1e-5, batch=2, seed=1),
    'p8x6x5_jacobi_f32': _c('poisson', (8, 6, 5), 'float32', 'CG', 1e-5, pre='jacobi', batch=3, seed=2),
    'p8x6x5_ilu_f64': _c('poisson', (8, 6, 5), 'float64', 'CG', 1e-10, pre='ilu', batch=2, seed=2),
    'a8x6x5_bicg1_f64': _c('advdiff', (8, 6, 5), 'float64', 'biCG-stab(1)', 1e-10, batch=2, seed=3),
    'a8x6x5_bicg2_f64': _c('advdiff', (8, 6, 5), 'float64', 'biCG-stab(2)', 1e-10, batch=2, seed=3),
    'a8x6x5_bicg2_f32': _c('advdiff', (8, 6, 5), 'float32', 'biCG-stab(2)', 1e-5, batch=2, seed=3),
    'a8x6x5_bicg3_f64': _c('advdiff', (8, 6, 5), 'float64', 'biCG-stab(3)', 1e-10, batch=2, seed=3),
    'a8x6x5_bicg1_jacobi_f64': _c('advdiff', (8, 6, 5), 'float64', 'biCG-stab(1)', 1e-10, pre='jacobi', batch=2, seed=3),
    # --- 3-D (configs 2 and 4 at sizes the reference finishes in seconds) ---
    'p32c_cg_f32': _c('poisson', (32, 32, 32), 'float32', 'CG', 1e-5),
    'p32c_jacobi_f32': _c('poisson', (32, 32, 32), 'float32', 'CG', 1e-5, pre='jacobi'),
    'p32c_ilu_f32': _c('poisson', (32, 32, 32), 'float32', 'CG', 1e-5, pre='ilu'),
    'p32c_ilu_f64': _c('poisson', (32, 32, 32), 'float64', 'CG', 1e-8, pre='ilu'),
    'p64c_cg_f32': _c('poisson', (64, 64, 64), 'float32', 'CG', 1e-5),
    'p64c_cg_f32_torch': _c('poisson', (64, 64, 64), 'float32', 'CG', 1e-5, backend='torch'),
    # --- config 3 in miniature: shared matrix, batch of RHS, min-tolerance quirk ---
    'p32_batch8_f32': _c('poisson', (32, 32), 'float32', 'CG', 1e-5, batch=8, seed=4),
    'p32_quirk_f32': _c('poisson', (32, 32), 'float32', 'CG', 1e-5, rhs_scale=(1.0, 1e-3), seed=5),
    'p32_quirk_f64_torch': _c('poisson', (32, 32), 'float64', 'CG', 1e-10, rhs_scale=(1.0, 1e-3), seed=5, backend='torch'),
    'p32_zero_rhs_f32': _c('poisson', (32, 32), 'float32', 'CG', 1e-5, 1e-5, batch=3, seed=6, zero_rhs_row=1),
    'p32_maxiter_f32': _c('poisson', (32, 32), 'float32', 'CG', 1e-5, batch=2, seed=7, max_iter=10),
    # --- config 5 in miniature: nonsymmetric, BiCGstab(2) ---
    'a24_bicg2_f32': _c('advdiff', (24, 20, 28), 'float32', 'biCG-stab(2)', 1e-5),
    'a24_bicg2_f64': _c('advdiff', (24, 20, 28), 'float64', 'biCG-stab(2)', 1e-10),
    'a24_bicg1_f32': _c('advdiff', (24, 20, 28), 'float32', 'biCG-stab(1)', 1e-5),
}


# operators with ZERO_GRADIENT / PERIODIC / mixed boundaries and piecewise-constant coefficients (tests/golden/boundaries.npz;
# the operator functions live in make_golden.py: BOUNDARY_OPERATORS)
BOUNDARY_CASES = [   # (operator, grid, dtype, method, preconditioner, rtol, batch)
    ('neumann', (12, 10, 8), 'float64', 'CG', 'jacobi', 1e-10, 2),
    ('neumann', (12, 10, 8), 'float32', 'CG', None, 1e-5, 2),
    ('neumann', (20, 16), 'float32', 'CG', 'jacobi', 1e-5, 1),
    ('periodic', (8, 12, 16), 'float64', 'CG', None, 1e-10, 2),
    ('periodic', (8, 12, 16), 'float32', 'CG', 'jacobi', 1e-5, 1),
    ('mixed', (10, 9, 12), 'float64', 'CG', 'jacobi', 1e-10, 2),
    ('mixed', (10, 9, 12), 'float32', 'CG-adaptive', None, 1e-5, 1),
    ('varcoef', (9, 8, 12), 'float64', 'biCG-stab(2)', None, 1e-10, 2),
    ('periodic_advdiff', (8, 6, 12), 'float64', 'biCG-stab(2)', None, 1e-10, 2),
    ('periodic_advdiff', (8, 6, 12), 'float32', 'biCG-stab(2)', 'jacobi', 1e-5, 1),
    ('varcoef', (9, 8, 12), 'float32', 'biCG-stab(1)', 'jacobi', 1e-5, 2),
    ('neumann', (31,), 'float64', 'CG', None, 1e-10, 1),
]




# `matrix @ x` outside the solver (tests/golden/dense.npz): Backend.mul_csr_dense / mul_coo_dense with a different matrix per batch entry.
# rows with no entries, duplicated coordinates and entries in arbitrary order included.  The reference's mul_coo_dense multiplies
# values (batch, nnz, channels) with gathered rows of width channels * dense_cols, so its COO cases need channels == 1 or dense_cols == 1.
DENSE_CASES = {
    'b3_c2_k4_f32': dict(batch=3, rows=37, cols=29, nnz=211, channels=2, dense_cols=4, dtype='float32', seed=11),
    'b1_c1_k1_f64': dict(batch=1, rows=64, cols=64, nnz=300, channels=1, dense_cols=1, dtype='float64', seed=12),
    'b2_c1_k3_f64': dict(batch=2, rows=50, cols=41, nnz=60, channels=1, dense_cols=3, dtype='float64', seed=13),
    'b4_c3_k1_f32': dict(batch=4, rows=33, cols=47, nnz=400, channels=3, dense_cols=1, dtype='float32', seed=14),
    'b1_c3_k2_f32': dict(batch=1, rows=19, cols=23, nnz=90, channels=3, dense_cols=2, dtype='float32', seed=15),
    'b5_c1_k1_f32': dict(batch=5, rows=300, cols=300, nnz=2100, channels=1, dense_cols=1, dtype='float32', seed=16),
}


def dense_case_inputs(name: str):
    """col (batch, nnz) int32, rowptr (batch, rows+1) int32 [CSR: ordered by row, columns ascending], coordinates (batch, nnz', 2) int64 [COO: the CSR
    entries in a random order followed by 9 repeated coordinates], values for both forms, dense (batch, cols, channels, dense_cols)."""
    c = DENSE_CASES[name]
    rng = np.random.default_rng(c['seed'])
    b, rows, cols, nnz, ch, k = c['batch'], c['rows'], c['cols'], c['nnz'], c['channels'], c['dense_cols']
    col = np.empty((b, nnz), np.int32)
    ptr = np.empty((b, rows + 1), np.int32)
    coo = np.empty((b, nnz + 9, 2), np.int64)
    for i in range(b):
        flat = np.sort(rng.choice(rows * cols, nnz, replace=False))
        r, cc = flat // cols, flat % cols
        col[i] = cc
        ptr[i] = np.searchsorted(r, np.arange(rows + 1))
        order = rng.permutation(nnz)
        coo[i, :nnz, 0], coo[i, :nnz, 1] = r[order], cc[order]
        coo[i, nnz:] = coo[i, :9]
    vals = rng.standard_normal((b, nnz, ch)).astype(c['dtype'])
    coo_vals = rng.standard_normal((b, nnz + 9, ch)).astype(c['dtype'])
    dense = rng.standard_normal((b, cols, ch, k)).astype(c['dtype'])
    return col, ptr, vals, coo, coo_vals, dense


# Edge inputs of Backend.linear_solve (tests/golden/edges.npz, recorded from the reference): zero / non-finite right-hand sides, iteration limits of
# 0 and 1, per-system limits and tolerances, zero tolerances, warm starts, loose absolute tolerances, 1- and 2-cell systems.
def _e(method, backend='numpy', op='poisson', grid=(7, 6), dtype='float32', batch=2, seed=31, y='random', x0='zero', rtol=1e-5, atol=0.0, max_iter=1000, pre=None):
    return dict(method=method, backend=backend, op=op, grid=grid, dtype=dtype, batch=batch, seed=seed, y=y, x0=x0, rtol=rtol, atol=atol, max_iter=max_iter, pre=pre)


EDGE_CASES = {}
for _m, _b in [('CG', 'numpy'), ('CG-adaptive', 'numpy'), ('biCG-stab(1)', 'numpy'), ('biCG-stab(2)', 'numpy'), ('CG', 'torch'), ('auto', 'torch')]:
    _t = f"{_m}_{_b}".replace('(', '').replace(')', '').replace('-', '')
    _op = 'advdiff' if _m.startswith('biCG') else 'poisson'
    _g = (5, 4, 6) if _m.startswith('biCG') else (7, 6)
    EDGE_CASES[f"zero_rhs/{_t}"] = _e(_m, _b, _op, _g, y='zero')
    EDGE_CASES[f"nan_rhs/{_t}"] = _e(_m, _b, _op, _g, batch=3, y='nan_row1')
    EDGE_CASES[f"inf_rhs/{_t}"] = _e(_m, _b, _op, _g, batch=2, y='inf_row0')
    EDGE_CASES[f"max_iter_0/{_t}"] = _e(_m, _b, _op, _g, max_iter=0)
    EDGE_CASES[f"max_iter_1/{_t}"] = _e(_m, _b, _op, _g, max_iter=1)
    EDGE_CASES[f"per_system/{_t}"] = _e(_m, _b, _op, _g, rtol=(1e-2, 1e-6), max_iter=(3, 1000))
    EDGE_CASES[f"zero_tol/{_t}"] = _e(_m, _b, _op, _g, rtol=0.0, atol=0.0, max_iter=7)
    EDGE_CASES[f"warm_start/{_t}"] = _e(_m, _b, _op, _g, x0='random', dtype='float64', rtol=1e-10)
    EDGE_CASES[f"loose_atol/{_t}"] = _e(_m, _b, _op, _g, atol=1e3)
    EDGE_CASES[f"one_cell/{_t}"] = _e(_m, _b, 'poisson', (1,), batch=2)
    EDGE_CASES[f"two_cells/{_t}"] = _e(_m, _b, 'poisson', (2,), batch=1, dtype='float64', rtol=1e-10)
    if _b == 'numpy' and not _m.startswith('CG-adaptive'):
        EDGE_CASES[f"jacobi_zero_rhs/{_t}"] = _e(_m, _b, _op, _g, y='zero_row0', batch=2, pre='jacobi')


# second group: checkpoint lists (Backend.while_loop's trajectory branch), ILU-preconditioned loops and BiCGstab(3 / 4) on edge inputs
for _m, _b in [('CG', 'numpy'), ('CG', 'torch'), ('CG-adaptive', 'numpy'), ('biCG-stab(1)', 'numpy'), ('biCG-stab(2)', 'numpy')]:
    _t = f"{_m}_{_b}".replace('(', '').replace(')', '').replace('-', '')
    _op = 'advdiff' if _m.startswith('biCG') else 'poisson'
    _g = (5, 4, 6) if _m.startswith('biCG') else (7, 6)
    EDGE_CASES[f"traj_repeated_checkpoints/{_t}"] = _e(_m, _b, _op, _g, rtol=(1e-1, 1e-6), max_iter=[[0, 0], [3, 3], [3, 3], [7, 7], [200, 200]])
    EDGE_CASES[f"traj_late_start/{_t}"] = _e(_m, _b, _op, _g, dtype='float64', rtol=1e-10, max_iter=[[4, 4], [9, 9], [1000, 1000]])
    EDGE_CASES[f"traj_zero_rhs_row/{_t}"] = _e(_m, _b, _op, _g, y='zero_row0', atol=1e-6, max_iter=[[0, 0], [1, 1], [2, 2], [100, 100]])
for _m in ('CG', 'biCG-stab(1)', 'biCG-stab(2)'):
    _t = f"{_m}_numpy".replace('(', '').replace(')', '').replace('-', '')
    _op = 'advdiff' if _m.startswith('biCG') else 'poisson'
    EDGE_CASES[f"ilu_zero_rhs_row/{_t}"] = _e(_m, 'numpy', _op, (5, 4, 6), y='zero_row0', atol=1e-7, pre='ilu')
    EDGE_CASES[f"ilu_max_iter_1/{_t}"] = _e(_m, 'numpy', _op, (5, 4, 6), max_iter=1, pre='ilu')
    EDGE_CASES[f"ilu_nan_rhs/{_t}"] = _e(_m, 'numpy', _op, (5, 4, 6), batch=3, y='nan_row1', pre='ilu', max_iter=30)
for _m in ('biCG-stab(3)', 'biCG-stab(4)'):
    _t = f"{_m}_numpy".replace('(', '').replace(')', '').replace('-', '')
    EDGE_CASES[f"zero_tol/{_t}"] = _e(_m, 'numpy', 'advdiff', (5, 4, 6), rtol=0.0, atol=0.0, max_iter=3)
    EDGE_CASES[f"zero_rhs/{_t}"] = _e(_m, 'numpy', 'advdiff', (5, 4, 6), y='zero')
    EDGE_CASES[f"max_iter_0/{_t}"] = _e(_m, 'numpy', 'advdiff', (5, 4, 6), max_iter=0)


def edge_case_inputs(name: str, n: int):
    """y, x0 (batch, n), rtol, atol (batch,), max_iter (1, batch) of an edge case for a matrix with n rows."""
    c = EDGE_CASES[name]
    rng = np.random.default_rng(c['seed'])
    nb = c['batch']
    y = rng.standard_normal((nb, n))
    x0 = rng.standard_normal((nb, n)) if c['x0'] == 'random' else np.zeros((nb, n))
    if c['y'] == 'zero':
        y[:] = 0
    elif c['y'] == 'zero_row0':
        y[0] = 0
    elif c['y'] == 'nan_row1':
        y[1, n // 2] = np.nan
    elif c['y'] == 'inf_row0':
        y[0, 0] = np.inf
    dt = np.dtype(c['dtype']).type
    rtol = np.broadcast_to(np.asarray(c['rtol'], dt), (nb,)).copy()
    atol = np.broadcast_to(np.asarray(c['atol'], dt), (nb,)).copy()
    mi = np.asarray(c['max_iter'])
    max_iter = mi.copy() if mi.ndim == 2 else np.broadcast_to(mi, (1, nb)).copy()
    return y.astype(dt), x0.astype(dt), rtol, atol, max_iter


# Products on natives whose arrays are legal but not canonical (tests/golden/spmv_edges.npz): entries of a row in arbitrary order, repeated
# coordinates, stored zeros, empty rows, rectangular shapes, non-finite operands.  SciPy (the reference's NumPy backend) walks a row's entries
# in STORED order, so does every product kernel here; matrices with stencil structure must not be re-ordered by the stencil recognition.
SPMV_EDGE_CASES = ['shuffled_rows_f32', 'shuffled_rows_f64', 'stored_zeros_f32', 'split_diagonal_f32', 'rectangular_f64', 'inf_times_stored_zero_f32',
                   'nan_in_x_f32', 'neumann_shuffled_rows_f32', 'all_rows_empty_f32']


def spmv_edge_case(name: str):
    """(indptr int32, indices int32, data, shape, x (batch, cols)) of a non-canonical CSR native."""
    import scipy.sparse as sp
    from oracle.assemble import stencil_csr, laplace_shifts
    dt = np.float64 if name.endswith('f64') else np.float32
    rng = np.random.default_rng(sum(map(ord, name)))
    grid = (6, 5, 4)
    base = stencil_csr(grid, laplace_shifts(3), dt)
    n = base.shape[0]
    if name.startswith('neumann'):      # ZERO_GRADIENT-type rows: the diagonal counts the neighbours inside the grid
        off = base - sp.diags(base.diagonal())
        base = (off + sp.diags(-np.asarray(off.sum(axis=1)).ravel() + 0.25)).tocsr().astype(dt)
        base.sort_indices()
    indptr, indices, data, shape = base.indptr.copy(), base.indices.copy(), base.data.copy(), base.shape
    x = rng.standard_normal((3, shape[1])).astype(dt)
    if 'shuffled_rows' in name:
        for r in range(n):
            p = rng.permutation(indptr[r + 1] - indptr[r]) + indptr[r]
            indices[indptr[r]:indptr[r + 1]], data[indptr[r]:indptr[r + 1]] = indices[p], data[p]
    elif name.startswith('stored_zeros') or name.startswith('inf_times'):
        rows, cols, vals = [], [], []
        for r in range(n):
            c = list(indices[indptr[r]:indptr[r + 1]]); v = list(data[indptr[r]:indptr[r + 1]])
            extra = (r + 11) % n
            if extra not in c:
                c.append(extra); v.append(0.0)
            order = np.argsort(c, kind='stable')
            rows += [r] * len(c); cols += [c[k] for k in order]; vals += [v[k] for k in order]
        indices, data = np.asarray(cols, np.int32), np.asarray(vals, dt)
        indptr = np.searchsorted(np.asarray(rows), np.arange(n + 1)).astype(np.int32)
        if name.startswith('inf_times'):
            x[0, 17] = np.inf
            x[1, 40] = -np.inf
    elif name.startswith('split_diagonal'):
        rows, cols, vals = [], [], []
        for r in range(n):
            for k in range(indptr[r], indptr[r + 1]):
                if indices[k] == r:
                    rows += [r, r]; cols += [r, r]; vals += [data[k] * 0.75, data[k] * 0.25]
                else:
                    rows.append(r); cols.append(indices[k]); vals.append(data[k])
        indices, data = np.asarray(cols, np.int32), np.asarray(vals, dt)
        indptr = np.searchsorted(np.asarray(rows), np.arange(n + 1)).astype(np.int32)
    elif name.startswith('rectangular'):
        shape = (37, 29)
        flat = np.sort(rng.choice(37 * 29, 150, replace=False))
        r, c = flat // 29, flat % 29
        keep = (r % 5) != 2           # rows 2, 7, 12, ... stay empty
        r, c = r[keep], c[keep]
        indices, data = c.astype(np.int32), rng.standard_normal(len(c)).astype(dt)
        indptr = np.searchsorted(r, np.arange(38)).astype(np.int32)
        x = rng.standard_normal((3, 29)).astype(dt)
    elif name.startswith('nan_in_x'):
        x[1, 33] = np.nan
    elif name.startswith('all_rows_empty'):
        indptr, indices, data = np.zeros(n + 1, np.int32), np.zeros(0, np.int32), np.zeros(0, dt)
    return indptr.astype(np.int32), indices.astype(np.int32), data, shape, x


# ILU sweeps on matrices that are not stencils (tests/golden/ilu_edges.npz, recorded from incomplete_lu_coo): structurally unsymmetric random patterns,
# a 1-D chain, a diagonal matrix, a dense block, an upper-triangular matrix (the reference returns identity / the matrix itself; a LOWER-triangular one makes
# its pattern helper raise IndexError, _linalg.py:618, so there is nothing to pin), and a rank-deficient matrix with `safe` sweeps.
ILU_EDGE_CASES = {   # name: (builder, dtype, sweeps, safe)
    'random_unsymmetric_f64': ('random', 'float64', 7, False),
    'random_unsymmetric_f32': ('random', 'float32', 5, False),
    'chain_f32': ('chain', 'float32', 9, False),
    'diagonal_f64': ('diagonal', 'float64', 3, False),
    'dense_block_f64': ('dense', 'float64', 6, False),
    'upper_triangular_f64': ('upper', 'float64', 4, False),
    'singular_safe_f64': ('singular', 'float64', 6, True),
    'one_by_one_f32': ('one', 'float32', 2, False),
}


def ilu_edge_matrix(name: str):
    import scipy.sparse as sp
    kind, dt, sweeps, safe = ILU_EDGE_CASES[name]
    rng = np.random.default_rng(sum(map(ord, name)))
    if kind == 'random':
        n = 40
        a = sp.random(n, n, density=0.12, random_state=np.random.RandomState(5), format='lil')
        a.setdiag(0)
        a = a.tocsr()
        a = a + sp.diags(np.asarray(abs(a).sum(axis=1)).ravel() + 1.0)
    elif kind == 'chain':
        n = 17
        a = sp.diags([-np.ones(n - 1), np.full(n, 2.5), -0.5 * np.ones(n - 1)], [-1, 0, 1])
    elif kind == 'diagonal':
        a = sp.diags(rng.uniform(1, 2, 12))
    elif kind == 'dense':
        a = sp.csr_matrix(rng.standard_normal((5, 5)) + 6 * np.eye(5))
    elif kind in ('upper', 'lower'):
        n = 15
        a = sp.random(n, n, density=0.3, random_state=np.random.RandomState(7), format='csr')
        a = (sp.triu(a, 1) if kind == 'upper' else sp.tril(a, -1)) + sp.diags(rng.uniform(1, 2, n))
    elif kind == 'singular':       # periodic 2-D Laplacian: constant null space; two decoupled copies -> rank deficiency 2
        n = 6
        p = sp.diags([-np.ones(n - 1), np.full(n, 2.0), -np.ones(n - 1)], [-1, 0, 1], format='lil')
        p[0, n - 1] = p[n - 1, 0] = -1
        lap = sp.kron(p, sp.identity(n)) + sp.kron(sp.identity(n), p)
        a = sp.block_diag([lap, lap])
    else:
        a = sp.csr_matrix(np.array([[3.0]]))
    a = sp.csr_matrix(a).astype(dt)
    a.sort_indices()
    a.eliminate_zeros()
    return a, sweeps, safe
