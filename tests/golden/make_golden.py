"""
Generates tests/golden/*.npz by RUNNING THE REAL REFERENCE (PhiML 1.15.1, imported from /root/reference).
Run in the authoring container only:   python tests/golden/make_golden.py
The GPU box has no /root/reference; tests read the committed .npz files.

Inputs are never stored: they are re-created from ``np.random.default_rng(seed)`` in the tests
(see tests/_cases.py, which this script shares with the tests).
"""
import hashlib
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))          # tests/
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))          # repo root (oracle.assemble builds the case matrices)
sys.path.insert(0, '/root/reference')

from phiml import math                                                         # noqa: E402
from phiml.math import spatial, batch, channel, Solve, SolveTape, extrapolation  # noqa: E402
from phiml.math._sparse import native_matrix                                   # noqa: E402
from phiml.math._lin_trace import matrix_from_function                         # noqa: E402
from phiml.math._optimize import factor_ilu                                    # noqa: E402
from phiml.backend._backend import get_backend, Preconditioner                 # noqa: E402
from phiml.backend import NUMPY                                                # noqa: E402

from _cases import SOLVE_CASES, rhs_for, sample_index, ADVDIFF, BOUNDARY_CASES, DENSE_CASES, dense_case_inputs, EDGE_CASES, edge_case_inputs, SPMV_EDGE_CASES, spmv_edge_case, ILU_EDGE_CASES, ilu_edge_matrix                   # noqa: E402

warnings.filterwarnings('ignore')
DIMS = 'xyz'


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def neg_laplace(x):
    return -math.laplace(x, padding=extrapolation.ZERO)


def advdiff(x):
    rank = x.shape.spatial_rank
    u = math.wrap(list(ADVDIFF['velocity'][:rank]), channel(gradient=','.join(DIMS[:rank])))
    g = math.spatial_gradient(x, padding=extrapolation.ZERO, difference='central', stack_dim=channel(gradient=','.join(DIMS[:rank])))
    return x + ADVDIFF['c'] * math.sum(g * u, 'gradient') - ADVDIFF['nu'] * math.laplace(x, padding=extrapolation.ZERO)


OPERATORS = {'poisson': neg_laplace, 'advdiff': advdiff}


def build(op: str, grid, dtype):
    """Reference assembly → (phiml matrix tensor, scipy csr, coo index array, coo values)."""
    with math.precision(64 if dtype == np.float64 else 32):
        x0 = math.zeros(spatial(**{DIMS[i]: g for i, g in enumerate(grid)}))
        coo, _ = matrix_from_function(OPERATORS[op], x0, auto_compress=False)
        m, _ = matrix_from_function(OPERATORS[op], x0, auto_compress=True)
        csr = native_matrix(m, NUMPY)
        idx = coo._indices.native([coo._indices.shape.non_channel, coo._indices.shape.channel])
        val = coo._values.numpy()
    return m, csr, np.asarray(idx), np.asarray(val)


class RefJacobi(Preconditioner):
    """5-line user preconditioner the reference's generic cg() is driven with."""
    def __init__(self, csr): self.inv = (1 / csr.diagonal()).astype(csr.dtype)
    def apply(self, vec): return vec * self.inv
    def apply_inv_l(self, vec): return vec * self.inv
    def apply_inv_u(self, vec): return vec
    def __repr__(self): return "jacobi"


def golden_assembly(out):
    for op, grid, full in [('poisson', (3, 4), True), ('poisson', (6, 5, 4), True), ('advdiff', (6, 5, 4), True), ('advdiff', (5, 7), True),
                           ('poisson', (128, 128), False), ('poisson', (32, 32, 32), False), ('poisson', (64, 64, 64), False), ('advdiff', (24, 20, 28), False)]:
        for dtype in (np.float32, np.float64):
            m, csr, idx, val = build(op, grid, dtype)
            key = f"asm/{op}/{'x'.join(map(str, grid))}/{np.dtype(dtype).name}"
            assert csr.indices.dtype == np.int32 and csr.indptr.dtype == np.int32 and csr.has_sorted_indices
            out[key + '/nnz'] = np.int64(csr.nnz)
            out[key + '/sha_indptr'] = sha(csr.indptr)
            out[key + '/sha_indices'] = sha(csr.indices)
            out[key + '/sha_data'] = sha(csr.data)
            out[key + '/sha_coo_idx'] = sha(idx.astype(np.int64))
            out[key + '/sha_coo_val'] = sha(val)
            if full:
                out[key + '/indptr'] = csr.indptr
                out[key + '/indices'] = csr.indices
                out[key + '/data'] = csr.data
                out[key + '/coo_idx'] = idx.astype(np.int64)
                out[key + '/coo_val'] = val
            print(key, csr.nnz)


def golden_spmv(out):
    for op, grid in [('poisson', (128, 128)), ('advdiff', (24, 20, 28))]:
        for dtype in (np.float32, np.float64):
            _, csr, _, _ = build(op, grid, dtype)
            x = np.random.default_rng(7).standard_normal((3, csr.shape[0])).astype(dtype)
            yv = NUMPY.mul_matrix_batched_vector(csr, x)
            out[f"spmv/{op}/{'x'.join(map(str, grid))}/{np.dtype(dtype).name}"] = yv


def golden_solves(out):
    for name, case in SOLVE_CASES.items():
        dtype = np.dtype(case['dtype']).type
        m, csr, _, _ = build(case['op'], case['grid'], dtype)
        y = rhs_for(case, csr.shape[0])
        x0 = np.zeros_like(y)
        nb = y.shape[0]
        rtol = np.full(nb, case['rtol'], dtype)
        atol = np.full(nb, case['atol'], dtype)
        max_iter = np.full((1, nb), case.get('max_iter', 5000))
        pre_name = case.get('pre')
        backend = get_backend(case.get('backend', 'numpy'))
        with math.precision(64 if dtype == np.float64 else 32):
            if pre_name == 'ilu':
                from phiml.math._optimize import compute_preconditioner
                pre = compute_preconditioner('ilu', m, 0, NUMPY, case['method'])
                out[f"solve/{name}/pre_source"] = str(pre)
            elif pre_name == 'jacobi':
                pre = RefJacobi(csr)
            else:
                pre = None
            if backend.name == 'torch':
                import torch
                lin = torch.sparse_csr_tensor(torch.as_tensor(csr.indptr), torch.as_tensor(csr.indices), torch.as_tensor(csr.data), csr.shape)
                args = [torch.as_tensor(a) for a in (y, x0, rtol, atol)]
            else:
                lin, args = csr, [y, x0, rtol, atol]
            with backend:
                ret = backend.linear_solve(case['method'], lin, *args, max_iter, pre, None)
        x = backend.numpy(ret.x)
        res = backend.numpy(ret.residual)
        key = f"solve/{name}"
        out[key + '/method'] = ret.method
        out[key + '/iterations'] = backend.numpy(ret.iterations)
        out[key + '/function_evaluations'] = backend.numpy(ret.function_evaluations)
        out[key + '/converged'] = backend.numpy(ret.converged)
        out[key + '/diverged'] = backend.numpy(ret.diverged)
        out[key + '/x_norm'] = np.linalg.norm(x.astype(np.float64), axis=1)
        out[key + '/res_norm'] = np.linalg.norm(res.astype(np.float64), axis=1)
        si = sample_index(x.shape[1])
        out[key + '/x_sample'] = x[:, si]
        if x.size <= 4096:
            out[key + '/x'] = x
        print(key, ret.method, out[key + '/iterations'][:8], out[key + '/function_evaluations'][:4], out[key + '/converged'][:4], out[key + '/diverged'][:4])


def golden_ilu(out):
    for grid, sweeps in [((5,), 10), ((8, 8), 6), ((6, 5, 4), 8), ((6, 5, 4), 40)]:
        with math.precision(64):
            m, csr, _, _ = build('poisson', grid, np.float64)
            lo, up = factor_ilu(m, sweeps)
            L = native_matrix(lo, NUMPY).tocsr()
            U = native_matrix(up, NUMPY).tocsr()
        L.sort_indices(); U.sort_indices()
        key = f"ilu/{'x'.join(map(str, grid))}/s{sweeps}"
        for nm, t in (('L', L), ('U', U)):
            out[f"{key}/{nm}_indptr"] = t.indptr
            out[f"{key}/{nm}_indices"] = t.indices
            out[f"{key}/{nm}_data"] = t.data
        rhs = np.random.default_rng(3).standard_normal((2, csr.shape[0]))
        out[f"{key}/solve_L"] = NUMPY.solve_triangular_sparse(L, rhs, True, True)
        out[f"{key}/solve_U"] = NUMPY.solve_triangular_sparse(U, rhs, False, False)
        print(key, L.nnz, U.nnz)


def golden_frontend(out):
    """Results through the full math.solve_linear front end (named dims, SolveTape, exceptions) on the reference's own backends."""
    from phiml.math import NotConverged
    f = math.jit_compile_linear(neg_laplace)
    f_adv = math.jit_compile_linear(advdiff)

    def record(key, x, info, order):
        out[key + '/x'] = x.numpy(order)
        out[key + '/iterations'] = info.iterations.numpy(order.split(',')[0]) if order.startswith('b') else np.asarray(info.iterations.numpy())
        out[key + '/function_evaluations'] = info.function_evaluations.numpy(order.split(',')[0]) if order.startswith('b') else np.asarray(info.function_evaluations.numpy())
        out[key + '/method'] = info.method
        print(key, info.method, out[key + '/iterations'], out[key + '/function_evaluations'])

    for backend_name in ('numpy', 'torch'):
        backend = get_backend(backend_name)
        with backend, math.precision(32):
            rng = np.random.default_rng(0)
            y = math.tensor(rng.standard_normal((3, 16, 16)).astype(np.float32), batch('b'), spatial('x,y'))
            x0 = math.zeros(spatial(x=16, y=16))
            for tag, method in (('cg16', 'CG'), ('auto16', 'auto')):
                solve = Solve(method, 1e-5, 0, x0=x0, max_iterations=1000)
                with SolveTape() as tape:
                    x = math.solve_linear(f, y, solve)
                record(f"front/{backend_name}/{tag}", x, tape[solve], 'b,x,y')
        with backend, math.precision(64):
            rng = np.random.default_rng(5)
            y = math.tensor(rng.standard_normal((2, 8, 6, 8)), batch('b'), spatial('x,y,z'))
            x0 = math.zeros(spatial(x=8, y=6, z=8))
            solve = Solve('biCG-stab(2)', 1e-10, 0, x0=x0, max_iterations=1000)
            with SolveTape() as tape:
                x = math.solve_linear(f_adv, y, solve)
            record(f"front/{backend_name}/bicg2_adv", x, tape[solve], 'b,x,y,z')
    # preconditioned solves: only the NumPy backend has sparse triangular solves (Backend.solve_triangular_sparse)
    with NUMPY:
        for prec, rtol, dt in ((32, 1e-5, np.float32), (64, 1e-10, np.float64)):
            with math.precision(prec):
                rng = np.random.default_rng(0)
                y = math.tensor(rng.standard_normal((3, 16, 16)).astype(dt), batch('b'), spatial('x,y'))
                x0 = math.zeros(spatial(x=16, y=16))
                for tag, pre in (('ilu16', 'ilu'), ('autopre16', 'auto')):
                    solve = Solve('CG', rtol, 0, x0=x0, max_iterations=1000, preconditioner=pre)
                    with SolveTape() as tape:
                        x = math.solve_linear(f, y, solve)
                    record(f"front/numpy/{tag}_f{prec}", x, tape[solve], 'b,x,y')
                rng = np.random.default_rng(2)
                y3 = math.tensor(rng.standard_normal((2, 8, 8, 8)).astype(dt), batch('b'), spatial('x,y,z'))
                x03 = math.zeros(spatial(x=8, y=8, z=8))
                solve = Solve('CG', rtol, 0, x0=x03, max_iterations=1000, preconditioner='ilu')
                with SolveTape() as tape:
                    x = math.solve_linear(f, y3, solve)
                record(f"front/numpy/ilu8c_f{prec}", x, tape[solve], 'b,x,y,z')
                # Jacobi has no name in the reference: its generic cg driven with the 5-line preconditioner object
                m, csr, _, _ = build('poisson', (16, 16), dt)
                yn = y.numpy('b,x,y').reshape(3, -1)
                ret = NUMPY.linear_solve('CG', csr, yn, np.zeros_like(yn), np.full(3, rtol, dt), np.zeros(3, dt), np.full((1, 3), 1000), RefJacobi(csr), None)
                key = f"front/numpy/jacobi16_f{prec}"
                out[key + '/x'] = ret.x.reshape(3, 16, 16)
                out[key + '/iterations'] = ret.iterations
                out[key + '/function_evaluations'] = ret.function_evaluations
                out[key + '/method'] = ret.method
                print(key, ret.method, ret.iterations)
        # trajectory recording (Backend.while_loop with one checkpoint per iteration, _backend.py:722-735)
        with math.precision(64):
            rng = np.random.default_rng(8)
            y = math.tensor(rng.standard_normal((2, 12, 12)), batch('b'), spatial('x,y'))
            x0 = math.zeros(spatial(x=12, y=12))
            solve = Solve('CG', 1e-10, 0, x0=x0, max_iterations=60)
            with SolveTape(record_trajectories=True) as tape:
                x = math.solve_linear(f, y, solve)
            info = tape[solve]
            key = "front/numpy/trj12"
            out[key + '/x'] = info.x.numpy('trajectory,b,x,y')
            out[key + '/residual'] = info.residual.numpy('trajectory,b,x,y')
            out[key + '/iterations'] = info.iterations.numpy('trajectory,b')
            out[key + '/function_evaluations'] = info.function_evaluations.numpy('trajectory,b')
            out[key + '/converged'] = info.converged.numpy('trajectory,b')
            out[key + '/final_x'] = x.numpy('b,x,y')
            print(key, out[key + '/x'].shape, out[key + '/iterations'][-1])
        # failure semantics: max_iterations reached -> NotConverged with the partial result attached
        with math.precision(32):
            rng = np.random.default_rng(0)
            y = math.tensor(rng.standard_normal((3, 16, 16)).astype(np.float32), batch('b'), spatial('x,y'))
            solve = Solve('CG', 1e-5, 0, x0=math.zeros(spatial(x=16, y=16)), max_iterations=5)
            try:
                math.solve_linear(f, y, solve)
                raise AssertionError("expected NotConverged")
            except NotConverged as exc:
                out["front/numpy/notconverged5/x"] = exc.result.x.numpy('b,x,y')
                out["front/numpy/notconverged5/iterations"] = exc.result.iterations.numpy('b')
                print("front/numpy/notconverged5", out["front/numpy/notconverged5/iterations"])


def advdiff_k(x, k):
    rank = x.shape.spatial_rank
    u = math.wrap(list(ADVDIFF['velocity'][:rank]), channel(gradient=','.join(DIMS[:rank])))
    g = math.spatial_gradient(x, padding=extrapolation.ZERO, difference='central', stack_dim=channel(gradient=','.join(DIMS[:rank])))
    return x + ADVDIFF['c'] * math.sum(g * u, 'gradient') - k * ADVDIFF['nu'] * math.laplace(x, padding=extrapolation.ZERO)


ADJOINT_CASES = {   # name: (grid, precision, method, rtol, batch, seed, grad_for_f)
    'adj_a8x6x8_bicg2_f64': ((8, 6, 8), 64, 'biCG-stab(2)', 1e-10, 2, 5, True),
    'adj_a8x6x8_bicg2_f32': ((8, 6, 8), 32, 'biCG-stab(2)', 1e-5, 2, 5, True),
    'adj_a12x10_bicg1_f64': ((12, 10), 64, 'biCG-stab(1)', 1e-10, 3, 6, True),
    'adj_a12x8x4_bicg2_f64_tight': ((12, 8, 4), 64, 'biCG-stab(2)', 1e-12, 1, 7, True),
    # without grad_for_f the matrix is constant: compressed natives (CSR forward, CSC = arrays of A^T backward)
    'adj_a8x6x8_bicg2_f64_const': ((8, 6, 8), 64, 'biCG-stab(2)', 1e-10, 2, 5, False),
    'adj_a16x12_bicg2_f32_const': ((16, 12), 32, 'biCG-stab(2)', 1e-5, 2, 8, False),
}


def golden_adjoint(out):
    """Backward pass of math.solve_linear(..., grad_for_f=True) on the reference's torch backend (CPU): what reaches
    Backend.linear_solve in the backward pass (layout of A^T, right-hand side dx = dL/dx), its result dy = dL/dy, and the matrix
    gradient dm on the stored entries (_optimize.py:785-817), recorded where the reference builds it (``matrix._with_values``)."""
    from phiml.math import _sparse
    TORCH = get_backend('torch')
    for name, (grid, prec, method, rtol, nb, seed, grad_f) in ADJOINT_CASES.items():
        rec = {'dm': [], 'calls': []}
        orig_with = _sparse.SparseCoordinateTensor._with_values
        orig_ls = TORCH.__class__.linear_solve

        def spy_with(self, new_values, **kw):
            rec['dm'].append((self, new_values))
            return orig_with(self, new_values, **kw)

        def spy_ls(self, meth, lin, y, x0, rtol_, atol_, max_iter, pre, matrix_offset):
            r = orig_ls(self, meth, lin, y, x0, rtol_, atol_, max_iter, pre, matrix_offset)
            rec['calls'].append((meth, str(lin.layout), lin, y.detach().numpy().copy(), r.x.detach().numpy().copy(), np.asarray(r.iterations).copy()))
            return r
        _sparse.SparseCoordinateTensor._with_values = spy_with
        TORCH.__class__.linear_solve = spy_ls
        try:
            with TORCH, math.precision(prec):
                dt = np.float64 if prec == 64 else np.float32
                f = math.jit_compile_linear(advdiff_k)
                rng = np.random.default_rng(seed)
                names = ','.join(DIMS[:len(grid)])
                y = math.tensor(rng.standard_normal((nb,) + grid).astype(dt), batch('b'), spatial(names))
                x0 = math.zeros(spatial(**{DIMS[i]: g for i, g in enumerate(grid)}))

                def loss(y, k):
                    x = math.solve_linear(f, y, Solve(method, rtol, 0, x0=x0, max_iterations=1000), k=k, grad_for_f=grad_f)
                    return math.l2_loss(x), x
                if grad_f:
                    (l, x), (dy, dk) = math.gradient(loss, 'y,k', get_output=True)(y, math.tensor(1.))
                else:
                    (l, x), dy = math.gradient(loss, 'y', get_output=True)(y, 1.)
                    dk = math.tensor(0.)
        finally:
            _sparse.SparseCoordinateTensor._with_values = orig_with
            TORCH.__class__.linear_solve = orig_ls
        with math.precision(prec):      # Tensor.numpy() converts to the CURRENT precision
            _adjoint_record(out, name, rec, x, dy, dk, grid, names, nb, grad_f)


def _adjoint_record(out, name, rec, x, dy, dk, grid, names, nb, grad_f):
    assert len(rec['calls']) == 2, len(rec['calls'])
    fwd, bwd = rec['calls']
    key = f"adjoint/{name}"
    order = 'b,' + names
    out[key + '/x'] = x.numpy(order).reshape(nb, -1)
    out[key + '/dy'] = dy.numpy(order).reshape(nb, -1)
    out[key + '/dk'] = np.asarray(dk.numpy())
    out[key + '/forward_layout'] = fwd[1]
    out[key + '/backward_layout'] = bwd[1]
    out[key + '/backward_rhs'] = bwd[3]                 # = dL/dx = x for the l2 loss
    out[key + '/backward_iterations'] = bwd[5]
    out[key + '/forward_iterations'] = fwd[5]
    # the matrices as they arrived at the boundary (COO natives, int64 indices, as torch stores them)
    import torch
    for tag, lin in (('forward', fwd[2]), ('backward', bwd[2])):
        lin = lin.detach()
        if lin.layout == torch.sparse_coo:
            parts = dict(indices=lin._indices(), values=lin._values())
        elif lin.layout == torch.sparse_csr:
            parts = dict(crow=lin.crow_indices(), col=lin.col_indices(), values=lin.values())
        else:
            parts = dict(ccol=lin.ccol_indices(), row=lin.row_indices(), values=lin.values())
        for pn, pv in parts.items():
            out[key + f'/{tag}_{pn}'] = pv.numpy()
    if not grad_f:
        print(key, fwd[1], bwd[1], fwd[5], bwd[5])
        return
    # the matrix gradient, entry by entry, and the entry coordinates it refers to
    dm_entries = [(m, v) for m, v in rec['dm'] if v.shape.volume == fwd[2]._values().numel()]
    m, v = dm_entries[-1]
    n = int(np.prod(grid))
    idx = np.asarray(m._indices.native([m._indices.shape.non_channel, m._indices.shape.channel]))
    rank = len(grid)
    col = np.ravel_multi_index(tuple(idx[:, i] for i in range(rank)), grid)         # dual (~x,~y,..) first: tracer_to_coo label order
    row = np.ravel_multi_index(tuple(idx[:, rank + i] for i in range(rank)), grid)
    out[key + '/dm'] = np.asarray(v.numpy())
    out[key + '/dm_row'] = row.astype(np.int64)
    out[key + '/dm_col'] = col.astype(np.int64)
    print(key, fwd[1], bwd[1], fwd[5], bwd[5], 'dk', float(out[key + '/dk']), 'dm', out[key + '/dm'].shape)


CLUSTER_CASES = {            # name: (op, grid, dtype, right-hand sides, seed, method)
    'c40x36_f64': ('poisson', (40, 36), np.float64, 2, 0, 'CG'),
    'c16x12x20_f32': ('poisson', (16, 12, 20), np.float32, 2, 1, 'CG'),
    'c12x10_f64_small': ('poisson', (12, 10), np.float64, 1, 2, 'CG'),          # fewer cells than 3^6 clusters: one cluster per cell
    'c20x14x10_adv_f64': ('advdiff', (20, 14, 10), np.float64, 2, 3, 'biCG-stab(2)'),
}


def golden_cluster(out):
    """The reference's 'cluster' preconditioner (explicit_coarse -> ExplicitClusterSolve, phiml/math/_optimize.py:942-971,
    phiml/backend/_linalg.py:734-799) on its NumPy backend: parts, one application, and solves driven with it."""
    from phiml.math._optimize import compute_preconditioner
    for name, (op, grid, dtype, nb, seed, method) in CLUSTER_CASES.items():
        prec = 64 if dtype == np.float64 else 32
        key = f"cluster/{name}"
        with math.precision(prec):
            m, csr, _, _ = build(op, grid, dtype)
            pre = compute_preconditioner('cluster', m, target_backend=NUMPY, solver=method)
            out[key + '/clusters'] = np.asarray(pre.clusters).astype(np.int32)
            out[key + '/cluster_count'] = np.int64(pre.cluster_count)
            out[key + '/cluster_size'] = np.asarray(pre.cluster_size)
            inv = np.asarray(pre.inv_coarse_matrix)
            out[key + '/inv_coarse'] = inv if inv.shape[0] <= 128 else inv[::7, ::5]      # strided sample keeps the fixture small
            rng = np.random.default_rng(seed)
            v = rng.standard_normal((nb, csr.shape[0])).astype(dtype)
            out[key + '/apply'] = np.asarray(pre.apply(v))
            y = rng.standard_normal((nb, csr.shape[0])).astype(dtype)
            rtol = np.asarray([1e-10 if prec == 64 else 1e-5] * nb, dtype)
            with NUMPY:
                r = NUMPY.linear_solve(method, csr, y, np.zeros_like(y), rtol, np.zeros(nb, dtype), np.array([[1000] * nb]), pre, None)
                r0 = NUMPY.linear_solve(method, csr, y, np.zeros_like(y), rtol, np.zeros(nb, dtype), np.array([[1000] * nb]), None, None)
            out[key + '/x'] = np.asarray(r.x)
            out[key + '/iterations'] = np.asarray(r.iterations)
            out[key + '/function_evaluations'] = np.asarray(r.function_evaluations)
            out[key + '/converged'] = np.asarray(r.converged)
            out[key + '/method'] = np.asarray(r.method)
            out[key + '/iterations_without'] = np.asarray(r0.iterations)
            print(key, repr(pre), r.method, r.iterations, 'without:', r0.iterations, out[key + '/inv_coarse'].dtype, out[key + '/cluster_size'].dtype, out[key + '/clusters'].shape)


# ----------------------------------------------------------------------------------------------------------------------------------
# Operators whose rows are not all alike (second half): ZERO_GRADIENT / PERIODIC / mixed boundaries, piecewise-constant
# coefficients.  The reference assembles them as general sparse matrices; phb200 recognises the CSR native as a pattern-coded stencil.
# ----------------------------------------------------------------------------------------------------------------------------------
from _boundary_ops import BOUNDARY_OPERATORS                                   # noqa: E402


def golden_boundaries(out):
    for op, grid, dt, method, pre_name, rtol, nb in BOUNDARY_CASES:
        dtype = np.dtype(dt).type
        with math.precision(64 if dtype == np.float64 else 32):
            x0 = math.zeros(spatial(**{DIMS[i]: g for i, g in enumerate(grid)}))
            m, _ = matrix_from_function(BOUNDARY_OPERATORS[op], x0, auto_compress=True)
            csr = native_matrix(m, NUMPY)
            assert csr.indices.dtype == np.int32 and csr.indptr.dtype == np.int32, (csr.indices.dtype, csr.indptr.dtype)
            key = f"bnd/{op}/{'x'.join(map(str, grid))}/{dt}/{method}/{pre_name}"
            out[key + '/indptr'] = csr.indptr
            out[key + '/indices'] = csr.indices
            out[key + '/data'] = csr.data
            out[key + '/sorted'] = np.bool_(csr.has_sorted_indices)
            rng = np.random.default_rng(11)
            v = rng.standard_normal((3, csr.shape[0])).astype(dtype)
            out[key + '/spmv'] = NUMPY.mul_matrix_batched_vector(csr, v)
            y = rng.standard_normal((nb, csr.shape[0])).astype(dtype)
            pre = RefJacobi(csr) if pre_name == 'jacobi' else None
            rt, at = np.full(nb, rtol, dtype), np.zeros(nb, dtype)
            with NUMPY:
                r = NUMPY.linear_solve(method, csr, y, np.zeros_like(y), rt, at, np.full((1, nb), 2000), pre, None)
            out[key + '/x'] = np.asarray(r.x)
            out[key + '/iterations'] = np.asarray(r.iterations)
            out[key + '/function_evaluations'] = np.asarray(r.function_evaluations)
            out[key + '/converged'] = np.asarray(r.converged)
            out[key + '/method'] = np.asarray(r.method)
            rows = np.diff(csr.indptr)
            print(key, csr.shape, csr.nnz, 'row lengths', np.unique(rows), 'sorted', csr.has_sorted_indices, 'zeros stored', int((csr.data == 0).sum()), r.method, r.iterations, r.converged)


def golden_dense(out):
    """`matrix @ x` outside the solver: NUMPY.mul_csr_dense (_numpy_backend.py:510-519) and Backend.mul_coo_dense (_backend.py:1303-1328; NumPy scatter =
    np.add.at, i.e. entry order) on a different matrix per batch entry."""
    for name, c in DENSE_CASES.items():
        col, ptr, vals, coo, coo_vals, dense = dense_case_inputs(name)
        shape = (c['rows'], c['cols'])
        with math.precision(64 if c['dtype'] == 'float64' else 32):      # auto_cast follows the global precision, not the operands
            out[f"dense/{name}/csr"] = NUMPY.mul_csr_dense(col, ptr, vals, shape, dense)
            if c['channels'] == 1 or c['dense_cols'] == 1:
                out[f"dense/{name}/coo"] = NUMPY.mul_coo_dense(coo, coo_vals, shape, dense)
            assert out[f"dense/{name}/csr"].dtype == np.dtype(c['dtype'])


def golden_edges(out):
    """Edge inputs through the reference's Backend.linear_solve (NumPy backend = generic loops; torch backend on the CPU = torch_sparse_cg rules)."""
    for name, c in EDGE_CASES.items():
        dtype = np.dtype(c['dtype']).type
        m, csr, _, _ = build(c['op'], c['grid'], dtype)
        y, x0, rtol, atol, max_iter = edge_case_inputs(name, csr.shape[0])
        backend = get_backend(c['backend'])
        pre = RefJacobi(csr) if c['pre'] == 'jacobi' else None
        with math.precision(64 if dtype == np.float64 else 32), np.errstate(all='ignore'):
            if c['pre'] == 'ilu':
                from phiml.math._optimize import compute_preconditioner
                pre = compute_preconditioner('ilu', m, 0, NUMPY, c['method'])
            if backend.name == 'torch':
                import torch
                lin = torch.sparse_csr_tensor(torch.as_tensor(csr.indptr), torch.as_tensor(csr.indices), torch.as_tensor(csr.data), csr.shape)
                args = [torch.as_tensor(a) for a in (y, x0, rtol, atol)]
            else:
                lin, args = csr, [y, x0, rtol, atol]
            with backend:
                ret = backend.linear_solve(c['method'], lin, *args, max_iter, pre, None)
        key = f"edge/{name}"
        out[key + '/method'] = ret.method
        for f in ('x', 'residual', 'iterations', 'function_evaluations', 'converged', 'diverged'):
            out[f"{key}/{f}"] = backend.numpy(getattr(ret, f))
        print(key, ret.method, out[key + '/iterations'], out[key + '/function_evaluations'], out[key + '/converged'], out[key + '/diverged'])


def golden_spmv_edges(out):
    """Products on non-canonical natives through the reference: NUMPY.mul_matrix_batched_vector on the CSR native as given (_numpy_backend.py:282-283)
    and on the same entries as a COO native."""
    import scipy.sparse as sp
    for name in SPMV_EDGE_CASES:
        indptr, indices, data, shape, x = spmv_edge_case(name)
        with np.errstate(all='ignore'):
            csr = sp.csr_matrix((data, indices, indptr), shape=shape)
            assert csr.indices is indices or np.array_equal(csr.indices, indices)      # SciPy keeps the arrays as they are
            out[f"spmv_edge/{name}/csr"] = NUMPY.mul_matrix_batched_vector(csr, x)
            rows = np.repeat(np.arange(shape[0]), np.diff(indptr))
            coo = sp.coo_matrix((data, (rows, indices)), shape=shape)
            out[f"spmv_edge/{name}/coo"] = NUMPY.mul_matrix_batched_vector(coo, x)


def golden_ilu_edges(out):
    """incomplete_lu_coo (phiml/backend/_linalg.py:524-594) on matrices that are not stencils; factors stored as the (indices, values) pairs it returns,
    in the entry order it returns them."""
    from phiml.backend._linalg import incomplete_lu_coo
    for name in ILU_EDGE_CASES:
        a, sweeps, safe = ilu_edge_matrix(name)
        coo = a.tocoo()      # CSR order: rows ascending, columns ascending
        idx = np.stack([coo.row, coo.col], -1)[None].astype(np.int64)
        with math.precision(64 if a.dtype == np.float64 else 32), np.errstate(all='ignore'):
            (li, lv), (ui, uv) = incomplete_lu_coo(idx, coo.data[None, :, None], a.shape, sweeps, safe)
        key = f"ilu_edge/{name}"
        out[key + '/L_idx'], out[key + '/L_val'] = np.asarray(li)[0], np.asarray(lv)[0, :, 0]
        out[key + '/U_idx'], out[key + '/U_val'] = np.asarray(ui)[0], np.asarray(uv)[0, :, 0]
        print(key, a.shape, a.nnz, out[key + '/L_val'].shape, out[key + '/U_val'].shape, out[key + '/L_val'].dtype, np.isfinite(out[key + '/U_val']).all())


def golden_frontend_products(out):
    """`matrix @ x` through the front end (dot_coordinate_dense -> Backend.mul_coo_dense, phiml/math/_sparse.py:1613; dot_compressed_dense ->
    Backend.mul_csr_dense, :1590) with batch and channel dimensions on x, on the reference's NumPy backend."""
    with NUMPY:
        for prec, dt in ((32, np.float32), (64, np.float64)):
            with math.precision(prec):
                f = math.jit_compile_linear(neg_laplace)          # per precision: the trace cache of a linear function does not key on it
                f_adv = math.jit_compile_linear(advdiff)
                rng = np.random.default_rng(9)
                x0 = math.zeros(spatial(x=16, y=12))
                v = math.tensor(rng.standard_normal((3, 16, 12, 2)).astype(dt), batch('b'), spatial('x,y'), channel('c'))
                coo, _ = matrix_from_function(neg_laplace, x0)
                csr, _ = f.sparse_matrix_and_bias(x0)
                assert type(coo).__name__ == 'SparseCoordinateTensor' and type(csr).__name__ == 'CompressedSparseMatrix'
                out[f"prod/coo_laplace_f{prec}"] = (coo @ v).numpy('b,x,y,c')
                out[f"prod/csr_laplace_f{prec}"] = (csr @ v).numpy('b,x,y,c')
                x03 = math.zeros(spatial(x=6, y=5, z=7))
                v3 = math.tensor(rng.standard_normal((6, 5, 7)).astype(dt), spatial('x,y,z'))
                csr3, _ = f_adv.sparse_matrix_and_bias(x03)
                out[f"prod/csr_advdiff_f{prec}"] = (csr3 @ v3).numpy('x,y,z')
                print(prec, [out[k].shape for k in out])


def main():
    groups = {'frontend_products': golden_frontend_products, 'ilu_edges': golden_ilu_edges, 'spmv_edges': golden_spmv_edges, 'edges': golden_edges, 'dense': golden_dense, 'boundaries': golden_boundaries, 'cluster': golden_cluster, 'assembly': golden_assembly, 'spmv': golden_spmv, 'solves': golden_solves, 'ilu': golden_ilu, 'frontend': golden_frontend, 'adjoint': golden_adjoint}
    only = sys.argv[1:] or list(groups)
    for g in only:
        out = {}
        groups[g](out)
        np.savez_compressed(os.path.join(HERE, f"{g}.npz"), **out)
        print(g, "->", os.path.getsize(os.path.join(HERE, f"{g}.npz")), "bytes")


if __name__ == '__main__':
    main()
