"""
CPU suite, part 3 (needs a copy of the reference: baseline/_ref or /root/reference): the PhiML plug-in keeps the reference's own front end
working and routes the path's calls to the phb200 boundary with the argument layout of SURVEY.md §8b.  The CUDA engine is
replaced by the oracle HERE, in the test, to observe the plumbing without a GPU.
"""
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.reference


@pytest.fixture()
def phiml_env():
    from _phiml import import_phiml
    assert import_phiml() is not None
    from phiml_b200 import phiml_plugin
    torch_backend = phiml_plugin.install()
    yield phiml_plugin, torch_backend
    phiml_plugin.uninstall()


def test_install_keeps_singleton_and_enables_triangular_solves(phiml_env):
    plugin, TORCH = phiml_env
    from phiml.backend._backend import Backend, get_backend
    assert get_backend('torch') is TORCH
    assert TORCH.supports(Backend.solve_triangular_sparse)
    assert TORCH.supports(Backend.mul_csr_dense) and TORCH.supports(Backend.conjugate_gradient)
    plugin.uninstall()
    assert not TORCH.supports(Backend.solve_triangular_sparse)


def test_cpu_operands_still_take_the_reference_path(phiml_env):
    """tests/commit/math/test__optimize.py:62-70 must keep passing on CPU torch tensors (not the phb200 path)."""
    from phiml import math
    from phiml.math import spatial, Solve, extrapolation
    plugin, TORCH = phiml_env
    with TORCH:
        y = math.ones(spatial(x=3))
        x0 = math.zeros(spatial(x=3))
        for method in ['CG', 'CG-adaptive', 'biCG-stab(1)']:
            f = math.jit_compile_linear(lambda x: math.laplace(x, padding=extrapolation.ZERO))
            x = math.solve_linear(f, y, Solve(method, 0, 1e-10, x0=x0))
            np.testing.assert_allclose(x.numpy('x'), [-1.5, -2, -1.5], atol=1e-3)
    with pytest.raises(NotImplementedError):
        from phiml.math._optimize import compute_preconditioner
        compute_preconditioner('jacobi', None, 0, None, 'CG')


def test_front_end_reaches_the_boundary_with_reference_layout(phiml_env, monkeypatch):
    """math.solve_linear -> Backend.linear_solve(method, native CSR, y (batch,N), x0, rtol (batch,), atol, max_iter (1,batch), pre)."""
    import torch
    import scipy.sparse as sp
    from phiml import math
    from phiml.math import spatial, batch, Solve, SolveTape, extrapolation
    from oracle import krylov
    from phiml_b200 import solve as host
    plugin, TORCH = phiml_env
    seen = {}

    def fake_cg(lin, y, x0, rtol, atol, max_iter, pre=None, matrix_offset=None):
        seen.update(layout=lin.layout, crow_dtype=lin.crow_indices().dtype, y=tuple(y.shape), x0=tuple(x0.shape), rtol=tuple(np.shape(rtol)),
                    max_iter=np.asarray(max_iter).shape, pre=pre, y_dtype=y.dtype)
        a = sp.csr_matrix((lin.values().numpy(), lin.col_indices().numpy(), lin.crow_indices().numpy()), shape=tuple(lin.shape))
        r = krylov.torch_fastpath_cg(a, y.numpy(), x0.numpy(), np.asarray(rtol), np.asarray(atol), np.asarray(max_iter))
        t = torch.as_tensor
        return host.SolveResult(r.method, t(r.x), t(r.residual), t(r.iterations), t(r.function_evaluations), t(r.converged), t(r.diverged), r.message)

    monkeypatch.setattr(plugin, '_on_path', lambda lin, y: True)
    monkeypatch.setattr(host, 'conjugate_gradient', fake_cg)
    f = math.jit_compile_linear(lambda x: -math.laplace(x, padding=extrapolation.ZERO))
    with TORCH, math.precision(32):
        rng = np.random.default_rng(0)
        y = math.tensor(rng.standard_normal((3, 16, 16)).astype(np.float32), batch('b'), spatial('x,y'))
        solve = Solve('CG', 1e-5, 0, x0=math.zeros(spatial(x=16, y=16)), max_iterations=1000)
        with SolveTape() as tape:
            x = math.solve_linear(f, y, solve)
        info = tape[solve]
    assert seen['layout'] == torch.sparse_csr and seen['crow_dtype'] == torch.int32
    assert seen['y'] == (3, 256) and seen['x0'] == (3, 256) and seen['rtol'] == (3,) and seen['max_iter'] == (1, 3)
    assert seen['pre'] is None and seen['y_dtype'] == torch.float32
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'frontend.npz'))
    np.testing.assert_allclose(x.numpy('b,x,y'), g['front/torch/cg16/x'], rtol=0, atol=2e-5)
    np.testing.assert_array_equal(info.iterations.numpy('b'), g['front/torch/cg16/iterations'])
    assert info.method == str(g['front/torch/cg16/method'])


def test_matrix_native_stays_resident_across_solves(phiml_env):
    """the reference converts / uploads the matrix on every solve_linear call (_optimize.py:650); with the plug-in
    the native of a live matrix tensor is built once per backend and dies with the tensor."""
    import gc
    from phiml import math
    from phiml.math import spatial, Solve, extrapolation
    plugin, TORCH = phiml_env
    plugin.NATIVE_CACHE.clear()
    plugin.NATIVE_CACHE.hits = plugin.NATIVE_CACHE.misses = 0
    f = math.jit_compile_linear(lambda x: -math.laplace(x, padding=extrapolation.ZERO))
    with TORCH:
        rng = np.random.default_rng(1)
        x0 = math.zeros(spatial(x=12, y=12))
        xs = []
        for k in range(3):
            y = math.tensor(rng.standard_normal((12, 12)).astype(np.float32), spatial('x,y'))
            xs.append(math.solve_linear(f, y, Solve('CG', 1e-5, 0, x0=x0)))
    assert plugin.NATIVE_CACHE.misses == 1 and plugin.NATIVE_CACHE.hits == 2
    assert len(plugin.NATIVE_CACHE._d) == 1
    # results are those of independent solves (no state leaks between calls)
    from oracle.assemble import poisson_csr
    a = poisson_csr((12, 12), np.float32)
    for x in xs:
        assert np.isfinite(x.numpy('x,y')).all()
    del f, xs
    gc.collect()
    assert len(plugin.NATIVE_CACHE._d) <= 1      # the entry goes away with the matrix tensor (LinearFunction dropped)


def test_backward_pass_reaches_the_boundary_with_the_transposed_native(phiml_env, monkeypatch):
    """math.gradient through solve_linear runs implicit_gradient_solve (_optimize.py:785-817); with the plug-in both
    solves arrive at phb200's host entry — forward with the CSR native, backward with the SAME arrays as a CSC native (A^T)."""
    import torch
    import scipy.sparse as sp
    from phiml import math
    from phiml.math import spatial, batch, channel, Solve
    from oracle import krylov
    from phiml_b200 import solve as host
    from _util import ADJOINT_CASES, ADVDIFF, adjoint_golden
    plugin, TORCH = phiml_env
    name = 'adj_a8x6x8_bicg2_f64_const'
    grid, dt, method, rtol, nb, seed, _ = ADJOINT_CASES[name]
    rec, mats = adjoint_golden(name)
    seen = []

    def fake_bicg(lin, y, x0, rtol_, atol_, max_iter, pre=None, matrix_offset=None, poly_order=2):
        if lin.layout == torch.sparse_csr:
            a = sp.csr_matrix((lin.values().numpy(), lin.col_indices().numpy(), lin.crow_indices().numpy()), shape=tuple(lin.shape))
        else:
            a = sp.csc_matrix((lin.values().numpy(), lin.row_indices().numpy(), lin.ccol_indices().numpy()), shape=tuple(lin.shape)).tocsr()
        seen.append((str(lin.layout), lin.values().data_ptr(), tuple(y.shape), poly_order))
        r = krylov.bicgstab_l(a, y.detach().numpy(), x0.detach().numpy(), np.asarray(rtol_), np.asarray(atol_), np.asarray(max_iter), None, poly_order, 'torch')
        t = torch.as_tensor
        return host.SolveResult(r.method, t(r.x), t(r.residual), t(r.iterations), t(r.function_evaluations), t(r.converged), t(r.diverged), r.message)

    monkeypatch.setattr(plugin, '_on_path', lambda lin, y: True)
    monkeypatch.setattr(host, 'bi_conjugate_gradient', fake_bicg)

    def advdiff(x):
        u = math.wrap(list(ADVDIFF['velocity']), channel(gradient='x,y,z'))
        g = math.spatial_gradient(x, padding=math.extrapolation.ZERO, difference='central', stack_dim=channel(gradient='x,y,z'))
        return x + ADVDIFF['c'] * math.sum(g * u, 'gradient') - ADVDIFF['nu'] * math.laplace(x, padding=math.extrapolation.ZERO)

    with TORCH, math.precision(64):
        f = math.jit_compile_linear(advdiff)
        y = math.tensor(np.random.default_rng(seed).standard_normal((nb,) + grid), batch('b'), spatial('x,y,z'))
        x0 = math.zeros(spatial(x=grid[0], y=grid[1], z=grid[2]))

        def loss(y):
            x = math.solve_linear(f, y, Solve(method, rtol, 0, x0=x0, max_iterations=1000))
            return math.l2_loss(x), x
        (l, x), dy = math.gradient(loss, 'y', get_output=True)(y)
        dy_np, x_np = dy.numpy('b,x,y,z').reshape(nb, -1), x.numpy('b,x,y,z').reshape(nb, -1)
    assert [s[0] for s in seen] == ['torch.sparse_csr', 'torch.sparse_csc']
    assert seen[0][1] == seen[1][1], "the backward native must view the forward arrays (zero copy transposition)"
    assert seen[0][2] == seen[1][2] == (nb, 384) and seen[1][3] == 2
    np.testing.assert_allclose(x_np, rec['x'], rtol=0, atol=1e-8 * np.abs(rec['x']).max())
    np.testing.assert_allclose(dy_np, rec['dy'], rtol=0, atol=1e-8 * np.abs(rec['dy']).max())


def test_tracer_bridge_builds_the_reference_matrix(phiml_env):
    """the bridge from ShiftLinTracer.val to the assembler returns the SAME CompressedSparseMatrix as the reference's
    tracer_to_coo + compress_rows (dims, int32 indices / pointers, values, rank) for constant-coefficient ZERO-boundary operators (stencil
    form) AND for shift tracers with one value per cell — PERIODIC wrap, ZERO_GRADIENT, variable coefficients (general form).  Arrays come from the oracle HERE (no GPU); on the device they come from
    phb200_stencil_to_csr / phb200_shift_count + phb200_shift_fill, pinned byte for byte by test_device_assembly_bit_exact and
    tests/test_gpu_coded.py::test_general_device_assembly_is_byte_identical."""
    import torch
    from phiml import math
    from phiml.math import spatial, channel, extrapolation
    from phiml.math._lin_trace import matrix_from_function
    import phiml.math._lin_trace as lt
    from phiml_b200 import assembly
    from oracle.assemble import stencil_csr, shift_values_csr
    from _util import ADVDIFF
    from _boundary_ops import varcoef, periodic_advdiff

    def neg_laplace(x): return -math.laplace(x, padding=extrapolation.ZERO)
    def periodic(x): return -math.laplace(x, padding=extrapolation.PERIODIC)
    def zero_gradient(x): return -math.laplace(x, padding=extrapolation.ZERO_GRADIENT)

    def advdiff(x):
        dims = 'xyz'[:x.shape.spatial_rank]
        u = math.wrap(list(ADVDIFF['velocity'][:len(dims)]), channel(gradient=','.join(dims)))
        g = math.spatial_gradient(x, padding=extrapolation.ZERO, difference='central', stack_dim=channel(gradient=','.join(dims)))
        return x + ADVDIFF['c'] * math.sum(g * u, 'gradient') - ADVDIFF['nu'] * math.laplace(x, padding=extrapolation.ZERO)

    def oracle_assemble(grid, offsets, coeffs, dtype):
        csr = stencil_csr(grid, list(zip(offsets, coeffs)), dtype.type)
        return torch.as_tensor(csr.indptr), torch.as_tensor(csr.indices), torch.as_tensor(csr.data)

    def oracle_assemble_general(grid, offsets, values, dtype):
        csr = shift_values_csr(grid, offsets, values)
        assert csr.data.dtype == dtype
        return torch.as_tensor(csr.indptr), torch.as_tensor(csr.indices), torch.as_tensor(csr.data)

    bridged = lt.tracer_to_coo            # the plug-in's wrapper (not eligible here: no GPU)
    original = phiml_env[0]._INSTALLED['original_t2c']
    try:
        for prec in (32, 64):
            for f, shape, expect in ((neg_laplace, dict(x=6, y=5), 'stencil'), (neg_laplace, dict(x=5, y=4, z=8), 'stencil'), (advdiff, dict(x=6, y=5, z=4), 'stencil'),
                                     (neg_laplace, dict(x=9), 'stencil'), (periodic, dict(x=6, y=5), 'general'), (periodic, dict(x=4, y=5, z=6), 'general'),
                                     (varcoef, dict(x=9, y=8, z=12), 'general'), (periodic_advdiff, dict(x=8, y=6, z=12), 'general'), (zero_gradient, dict(x=6, y=5), 'general'), (zero_gradient, dict(x=5, y=6, z=4), 'general')):
                with math.precision(prec):
                    x0 = math.zeros(spatial(**shape))
                    lt.tracer_to_coo = original
                    ref, _ = matrix_from_function(f, x0, auto_compress=True)
                    assembly.STATS.update(device_assemblies=0, general_assemblies=0, fallbacks=0)
                    lt.tracer_to_coo = assembly.make_tracer_to_coo(original, lambda: True, oracle_assemble, oracle_assemble_general)
                    mine, _ = matrix_from_function(f, x0, auto_compress=True)
                assert (assembly.STATS['device_assemblies'] == 1) == (expect == 'stencil'), (f.__name__, assembly.STATS)
                assert (assembly.STATS['general_assemblies'] == 1) == (expect == 'general'), (f.__name__, assembly.STATS)
                assert type(mine) is type(ref) and mine.shape == ref.shape
                assert mine._compressed_dims == ref._compressed_dims and mine._uncompressed_dims == ref._uncompressed_dims
                for name, dim in (('_indices', 'sp_entries'), ('_pointers', 'pointers'), ('_values', 'sp_entries')):
                    a, b = np.asarray(getattr(mine, name).native(dim)), getattr(ref, name).numpy(dim)
                    assert a.dtype == b.dtype and np.array_equal(a, b), (f.__name__, name)
                assert float(mine._matrix_rank) == float(ref._matrix_rank)
                # the coordinate form: compress_rows() keeps the coordinates it came from, the bridge rebuilds them on demand (phiml_plugin: decompress)
                assert mine._uncompressed_indices is None
                c_mine, c_ref = mine.decompress(), ref.decompress()
                assert type(c_mine) is type(c_ref) and c_mine.shape == c_ref.shape
                assert c_mine._indices.shape == c_ref._indices.shape and c_mine._indices.dtype == c_ref._indices.dtype
                order = c_ref._indices.shape.names
                assert np.array_equal(np.asarray(c_mine._indices.native(order)), c_ref._indices.numpy(order)), f.__name__
                assert np.array_equal(np.asarray(c_mine._values.native('sp_entries')), c_ref._values.numpy('sp_entries'))
                assert (c_mine._can_contain_double_entries, c_mine._indices_sorted, c_mine._indices_constant) == (c_ref._can_contain_double_entries, c_ref._indices_sorted, c_ref._indices_constant)
                assert c_mine._indices.default_backend.name == 'numpy'        # what the reference's packing utilities insist on (_sparse.py:394)
                probe = math.tensor(np.random.default_rng(2).standard_normal(tuple(shape.values())), spatial(','.join(shape)))
                order_x = ','.join(shape)
                assert np.allclose((c_mine @ probe).numpy(order_x), (c_ref @ probe).numpy(order_x), rtol=1e-6 if prec == 32 else 1e-13, atol=1e-6 if prec == 32 else 1e-13)
        # auto_compress=False (the default of the public function) asks for the coordinate form: the bridge leaves it to the reference
        with math.precision(32):
            x0 = math.zeros(spatial(x=6, y=5))
            lt.tracer_to_coo = original
            ref, _ = matrix_from_function(neg_laplace, x0)
            assembly.STATS.update(device_assemblies=0, general_assemblies=0, fallbacks=0)
            lt.tracer_to_coo = assembly.make_tracer_to_coo(original, lambda: True, oracle_assemble, oracle_assemble_general)
            mine, _ = matrix_from_function(neg_laplace, x0)
        assert type(ref).__name__ == 'SparseCoordinateTensor' and type(mine) is type(ref)
        assert assembly.STATS['device_assemblies'] == 0 and assembly.STATS['fallbacks'] == 1
        assert np.array_equal(mine._indices.numpy(mine._indices.shape.names), ref._indices.numpy(ref._indices.shape.names))
        assert np.array_equal(mine._values.numpy(mine._values.shape.names), ref._values.numpy(ref._values.shape.names))
    finally:
        lt.tracer_to_coo = bridged


def test_front_end_solves_on_the_bridged_matrix(phiml_env):
    """math.solve_linear with a jit_compile_linear function whose matrix came through the bridge (torch CPU natives here) gives the
    reference's result: the bridged CompressedSparseMatrix is a drop-in for everything downstream (native_matrix, packing, SolveTape)."""
    import os
    import torch
    from phiml import math
    from phiml.math import spatial, batch, Solve, SolveTape, extrapolation
    import phiml.math._lin_trace as lt
    from phiml_b200 import assembly
    from oracle.assemble import stencil_csr
    plugin, TORCH = phiml_env

    def oracle_assemble(grid, offsets, coeffs, dtype):
        csr = stencil_csr(grid, list(zip(offsets, coeffs)), dtype.type)
        return torch.as_tensor(csr.indptr), torch.as_tensor(csr.indices), torch.as_tensor(csr.data)

    bridged = lt.tracer_to_coo
    lt.tracer_to_coo = assembly.make_tracer_to_coo(plugin._INSTALLED['original_t2c'], lambda: True, oracle_assemble)
    try:
        assembly.STATS.update(device_assemblies=0, fallbacks=0)
        f = math.jit_compile_linear(lambda x: -math.laplace(x, padding=extrapolation.ZERO))
        with TORCH, math.precision(32):
            y = math.tensor(np.random.default_rng(0).standard_normal((3, 16, 16)).astype(np.float32), batch('b'), spatial('x,y'))
            solve = Solve('CG', 1e-5, 0, x0=math.zeros(spatial(x=16, y=16)), max_iterations=1000)
            with SolveTape() as tape:
                x = math.solve_linear(f, y, solve)
            info = tape[solve]
        assert assembly.STATS['device_assemblies'] == 1
    finally:
        lt.tracer_to_coo = bridged
    g = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'frontend.npz'))
    np.testing.assert_allclose(x.numpy('b,x,y'), g['front/torch/cg16/x'], rtol=0, atol=2e-5)
    np.testing.assert_array_equal(info.iterations.numpy('b'), g['front/torch/cg16/iterations'])


def test_jit_compiled_solve_is_one_registered_operator_node(phiml_env, monkeypatch):
    """under ``math.jit_compile`` PhiML traces with torch.jit.trace; the plug-in must put the solve into the graph as ONE
    registered operator (phb200::linear_solve) that runs the device loop on every execution — not an unrolled Python loop and not a
    constant baked in at trace time.  Engine replaced by the oracle here; the operator, its schema and the routing are the real ones."""
    import torch
    import scipy.sparse as sp
    from phiml import math
    from phiml.math import spatial, Solve, extrapolation
    from oracle import krylov
    from phiml_b200 import solve as host, torch_op
    plugin, TORCH = phiml_env
    seen = []

    def fake_linear_solve(method, lin, y, x0, rtol, atol, max_iter, pre=None, matrix_offset=None):
        seen.append((method, lin.layout, tuple(y.shape), np.asarray(max_iter).shape, pre))
        a = sp.csr_matrix((lin.values().numpy(), lin.col_indices().numpy(), lin.crow_indices().numpy()), shape=tuple(lin.shape))
        r = krylov.torch_fastpath_cg(a, y.numpy(), x0.numpy(), rtol.numpy(), atol.numpy(), np.asarray(max_iter))
        t = torch.as_tensor
        return host.SolveResult(r.method, t(r.x), t(r.residual), t(r.iterations), t(r.function_evaluations), t(r.converged), t(r.diverged), r.message)

    monkeypatch.setattr(plugin, '_traceable', lambda lin, y: torch._C._get_tracing_state() is not None and isinstance(lin, torch.Tensor) and lin.layout != torch.strided)
    monkeypatch.setattr(host, 'linear_solve', fake_linear_solve)
    lap = math.jit_compile_linear(lambda x: -math.laplace(x, padding=extrapolation.ZERO))

    @math.jit_compile
    def solve_for(y):
        return math.solve_linear(lap, y, Solve('CG', 1e-5, 0, x0=math.zeros(spatial(x=12, y=10)), max_iterations=500))

    from oracle.assemble import poisson_csr
    a = poisson_csr((12, 10), np.float32)
    rng = np.random.default_rng(3)
    calls0 = torch_op._STATE['calls']
    with TORCH, math.precision(32):
        for k in range(3):                      # first call traces, the next ones execute the traced graph
            y_np = rng.standard_normal((12, 10)).astype(np.float32)
            x = solve_for(math.tensor(y_np, spatial('x,y')))
            ref = krylov.torch_fastpath_cg(a, y_np.reshape(1, -1), np.zeros((1, 120), np.float32), np.float32([1e-5]), np.float32([0]), np.array([[500]]))
            np.testing.assert_allclose(x.numpy('x,y').reshape(-1), ref.x[0], rtol=0, atol=1e-6)
    assert torch_op._STATE['calls'] - calls0 >= 3, "every execution of the traced graph must run the operator (no baked-in constants)"
    assert all(s[0] == 'CG' and s[1] == torch.sparse_csr and s[2] == (1, 120) and s[3] == (1, 1) and s[4] is None for s in seen)
    # the traced graph of the compiled function holds the operator as a node
    parts, layout = torch_op.parts_of_native(torch.sparse_csr_tensor(torch.as_tensor(a.indptr), torch.as_tensor(a.indices), torch.as_tensor(a.data), a.shape))

    def direct(y):
        return torch_op.register()(parts, layout, 120, 120, y, torch.zeros_like(y), torch.full((1,), 1e-5), torch.zeros(1), [500], 1, 'CG', '')[0]
    graph = torch.jit.trace(direct, (torch.ones(1, 120),), check_trace=False).graph
    assert 'phb200::linear_solve' in str(graph)


def test_matrix_gradient_inside_phimls_backward_pass_comes_from_the_boundary(phiml_env, monkeypatch):
    """grad_for_f=True: implicit_gradient_solve forms dm = -(dy[row] * x[col]) summed over the batch (_optimize.py:803-807).  With the
    plug-in PhiML's own backward pass hands (row, col, dy, x) to phb200's matrix-gradient entry; here that entry is the oracle's
    restatement, and dL/dk must equal what the unmodified reference gives."""
    import torch
    from phiml import math
    from phiml.math import spatial, batch, channel, Solve
    from oracle import adjoint as oracle_adjoint
    from phiml_b200 import adjoint as host_adjoint
    from _util import ADVDIFF
    plugin, TORCH = phiml_env
    seen = []

    def fake_matrix_gradient(native, dy, x):
        idx = native._indices().numpy()
        seen.append((tuple(dy.shape), tuple(x.shape), idx.shape))
        return torch.as_tensor(oracle_adjoint.matrix_gradient(idx[0], idx[1], dy.numpy(), x.numpy()))

    def advdiff_k(x, k):
        u = math.wrap(list(ADVDIFF['velocity'][:2]), channel(gradient='x,y'))
        g = math.spatial_gradient(x, padding=math.extrapolation.ZERO, difference='central', stack_dim=channel(gradient='x,y'))
        return x + ADVDIFF['c'] * math.sum(g * u, 'gradient') - k * ADVDIFF['nu'] * math.laplace(x, padding=math.extrapolation.ZERO)

    def run():
        with TORCH, math.precision(64):
            f = math.jit_compile_linear(advdiff_k)
            y = math.tensor(np.random.default_rng(4).standard_normal((2, 10, 8)), batch('b'), spatial('x,y'))
            x0 = math.zeros(spatial(x=10, y=8))

            def loss(y, k):
                x = math.solve_linear(f, y, Solve('biCG-stab(2)', 1e-10, 0, x0=x0, max_iterations=1000), k=k, grad_for_f=True)
                return math.l2_loss(x), x
            (l, x), (dy, dk) = math.gradient(loss, 'y,k', get_output=True)(y, math.tensor(1.))
            return dy.numpy('b,x,y'), float(dk.numpy())

    dy_ref, dk_ref = run()                       # operands on the CPU: the reference's own rule
    assert plugin.DM_STATS['device'] == 0
    monkeypatch.setattr(plugin, '_device_tensor', lambda t: isinstance(t, torch.Tensor))
    monkeypatch.setattr(host_adjoint, 'matrix_gradient', fake_matrix_gradient)
    dy_dev, dk_dev = run()
    assert plugin.DM_STATS['device'] == 1 and len(seen) == 1
    assert seen[0][0] == (2, 80) and seen[0][1] == (2, 80) and seen[0][2][0] == 2
    np.testing.assert_allclose(dy_dev, dy_ref, rtol=0, atol=1e-12 * np.abs(dy_ref).max())
    assert abs(dk_dev - dk_ref) <= 1e-10 * max(1.0, abs(dk_ref))
    plugin.DM_STATS.update(device=0, reference=0)


def test_bridged_matrix_behaves_like_the_reference_matrix_under_phimls_operations(phiml_env):
    """A device-assembled (bridged) CompressedSparseMatrix is a PhiML tensor like any other: 40 operations a user can apply to the matrix — products with
    batch / channel dimensions, arithmetic, reductions, format changes, the coordinate form and its products, math.factor_ilu, casts, conversion to the
    NumPy backend, natives for SciPy and torch — give the reference matrix' result, and raise where the reference itself raises."""
    import torch
    from phiml import math
    from phiml.math import spatial, batch, channel, dual, extrapolation
    from phiml.backend import NUMPY
    from phiml.math._sparse import native_matrix
    import phiml.math._lin_trace as lt
    from phiml_b200 import assembly
    from oracle.assemble import stencil_csr
    TORCH = phiml_env[1]

    def neg_laplace(x): return -math.laplace(x, padding=extrapolation.ZERO)

    def oracle_assemble(grid, offsets, coeffs, dtype):
        csr = stencil_csr(grid, list(zip(offsets, coeffs)), dtype.type)
        return torch.as_tensor(csr.indptr), torch.as_tensor(csr.indices), torch.as_tensor(csr.data)

    ops = {
        'matmul': lambda M, v: (M @ v).numpy('x,y'),
        'matmul_batched': lambda M, v: (M @ math.stack([v, 2 * v], batch('b'))).numpy('b,x,y'),
        'matmul_channels': lambda M, v: (M @ math.stack([v, 2 * v], channel('c'))).numpy('x,y,c'),
        'dense': lambda M, v: math.dense(M).numpy('x,y,~x,~y'),
        'scale': lambda M, v: ((M * 2.5) @ v).numpy('x,y'),
        'neg': lambda M, v: ((-M) @ v).numpy('x,y'),
        'add': lambda M, v: ((M + M) @ v).numpy('x,y'),
        'sub': lambda M, v: ((M - 0.5 * M) @ v).numpy('x,y'),
        'abs': lambda M, v: (abs(M) @ v).numpy('x,y'),
        'pow': lambda M, v: ((M ** 2) @ v).numpy('x,y'),
        'where': lambda M, v: (math.where(M > 0, M, 0) @ v).numpy('x,y'),
        'mul_tensor': lambda M, v: ((M * v) @ v).numpy('x,y'),
        'stored_values': lambda M, v: math.stored_values(M).numpy('entries'),
        'stored_indices': lambda M, v: math.stored_indices(M).numpy('entries,index'),
        'sum_dual': lambda M, v: math.sum(M, dual).numpy('x,y'),
        'sum_primal': lambda M, v: math.sum(M, 'x,y').numpy('~x,~y'),
        'sum_all': lambda M, v: np.asarray(float(math.sum(M, M.shape))),
        'max_all': lambda M, v: np.asarray(float(math.max(M, M.shape))),
        'sparsity': lambda M, v: np.asarray(float(math.get_sparsity(M))),
        'is_sparse': lambda M, v: np.asarray(math.is_sparse(M)),
        'matrix_rank': lambda M, v: np.asarray(float(M._matrix_rank)),
        'repr': lambda M, v: np.asarray(len(repr(M)) > 0),
        'close': lambda M, v: np.asarray(bool(math.close(M, M))),
        'copy': lambda M, v: (math.copy(M) @ v).numpy('x,y'),
        'rename': lambda M, v: math.dense(math.rename_dims(M, 'x', 'a')).numpy('a,y,~x,~y'),
        'slice_row': lambda M, v: math.dense(M[{'x': 1}]).numpy('y,~x,~y'),
        'to_dense': lambda M, v: (math.to_format(M, 'dense') @ v).numpy('x,y'),
        'to_csc': lambda M, v: (math.to_format(M, 'csc') @ v).numpy('x,y'),
        'to_coo': lambda M, v: (math.to_format(M, 'coo') @ v).numpy('x,y'),
        'coo_dense': lambda M, v: math.dense(math.to_format(M, 'coo')).numpy('x,y,~x,~y'),
        'coo_scaled': lambda M, v: ((math.to_format(M, 'coo') * 3) @ v).numpy('x,y'),
        'coo_to_csr': lambda M, v: (math.to_format(math.to_format(M, 'coo'), 'csr') @ v).numpy('x,y'),
        'decompress': lambda M, v: (M.decompress() @ v).numpy('x,y'),
        'factor_ilu': lambda M, v: np.concatenate([math.dense(t).numpy('x,y,~x,~y').ravel() for t in math.factor_ilu(M, 4)]),
        'to_float': lambda M, v: (math.to_float(M) @ v).numpy('x,y'),
        'cast32': lambda M, v: (math.cast(M, math.DType(float, 32)) @ math.cast(v, math.DType(float, 32))).numpy('x,y'),
        'convert_numpy': lambda M, v: (math.convert(M, NUMPY) @ math.convert(v, NUMPY)).numpy('x,y'),
        'scipy_native': lambda M, v: native_matrix(M, NUMPY).toarray(),
        'torch_native': lambda M, v: native_matrix(M, TORCH).to_dense().numpy(),
        'trace_through': lambda M, v: (math.matrix_from_function(lambda x: M @ x, v)[0] @ v).numpy('x,y'),
    }
    bridged = lt.tracer_to_coo
    original = phiml_env[0]._INSTALLED['original_t2c']
    try:
        with TORCH, math.precision(64):
            x0 = math.zeros(spatial(x=6, y=5))
            v = math.tensor(np.random.default_rng(1).standard_normal((6, 5)), spatial('x,y'))
            lt.tracer_to_coo = original
            ref, _ = math.jit_compile_linear(lambda x: neg_laplace(x)).sparse_matrix_and_bias(x0)
            lt.tracer_to_coo = assembly.make_tracer_to_coo(original, lambda: True, oracle_assemble, None)
            mine, _ = math.jit_compile_linear(lambda x: neg_laplace(x)).sparse_matrix_and_bias(x0)
            assert getattr(mine, '_phb200_device_assembled', False) and mine._uncompressed_indices is None
            ok = 0
            for name, op in ops.items():
                res = []
                for M in (ref, mine):
                    try:
                        res.append(('ok', op(M, v)))
                    except Exception as exc:
                        res.append(('raised', type(exc).__name__))
                (s0, a), (s1, b) = res
                assert s0 == s1, f"{name}: the reference matrix {s0} ({a if s0 == 'raised' else ''}), the bridged matrix {s1} ({b if s1 == 'raised' else ''})"
                if s0 == 'ok':
                    ok += 1
                    assert np.shape(a) == np.shape(b) and np.allclose(a, b, rtol=1e-13, atol=1e-13), name
            assert ok >= 33, ok
    finally:
        lt.tracer_to_coo = bridged


def test_numpy_backed_inputs_keep_the_references_numpy_matrix(phiml_env):
    """The backend a matrix is FOR is the one of the traced tensor (matrix_from_function: target_backend = backend_for(*tensors), _lin_trace.py:714): with
    torch as the default backend but NumPy-backed y / x0 (math.wrap(ndarray)), the reference solves on its NumPy backend with a NumPy-backed matrix — a
    bridged (torch-backed) matrix would reach NUMPY.linear_solve as a torch sparse tensor.  The bridge follows the traced tensor: NumPy-backed inputs are
    left to the reference, torch-backed ones are bridged; both give the reference's results for Φ-ML and SciPy methods."""
    import torch
    from phiml import math
    from phiml.math import spatial, batch, extrapolation, Solve, SolveTape
    from phiml.backend import default_backend
    import phiml.math._lin_trace as lt
    from phiml_b200 import assembly
    from oracle.assemble import stencil_csr
    TORCH = phiml_env[1]

    def neg_laplace(x): return -math.laplace(x, padding=extrapolation.ZERO)

    def oracle_assemble(grid, offsets, coeffs, dtype):
        csr = stencil_csr(grid, list(zip(offsets, coeffs)), dtype.type)
        return torch.as_tensor(csr.indptr), torch.as_tensor(csr.indices), torch.as_tensor(csr.data)

    bridged = lt.tracer_to_coo
    original = phiml_env[0]._INSTALLED['original_t2c']
    # the plug-in's own rule (phiml_plugin.wants_device_matrix) with "a CUDA device is present" taken for granted
    rule = lambda tb=None: (tb is TORCH) if tb is not None else (default_backend() is TORCH)
    try:
        for kind, make in (('numpy-backed', math.wrap), ('torch-backed', math.tensor)):
            res = {}
            for name, t2c in (('reference', original), ('bridge', assembly.make_tracer_to_coo(original, rule, oracle_assemble, None))):
                lt.tracer_to_coo = t2c
                assembly.STATS.update(device_assemblies=0, general_assemblies=0, fallbacks=0)
                f = math.jit_compile_linear(lambda x: neg_laplace(x))
                with TORCH, math.precision(64):
                    y = make(np.random.default_rng(0).standard_normal((2, 8, 6)), batch('b'), spatial('x,y'))
                    x0 = make(np.zeros((8, 6)), spatial('x,y'))
                    for method in ('CG', 'auto', 'biCG-stab(2)', 'scipy-direct'):
                        solve = Solve(method, 1e-9, 0, x0=x0, max_iterations=500)
                        with SolveTape() as tape:
                            x = math.solve_linear(f, y, solve)
                        res[(name, method)] = (x.numpy('b,x,y'), tape[solve].method)
                if name == 'bridge':
                    assert assembly.STATS['device_assemblies'] == (1 if kind == 'torch-backed' else 0), (kind, assembly.STATS)
            for method in ('CG', 'auto', 'biCG-stab(2)', 'scipy-direct'):
                (a, ma), (b, mb) = res[('reference', method)], res[('bridge', method)]
                assert ma == mb and np.allclose(a, b, rtol=1e-7, atol=1e-9), (kind, method, ma, mb)
                assert ('numpy' in ma) == (kind == 'numpy-backed') or method == 'scipy-direct' or 'PyTorch' in ma, (kind, method, ma)
    finally:
        lt.tracer_to_coo = bridged


def test_batch_and_channel_dimensions_on_x0_through_the_bridge(phiml_env):
    """solve_linear with batch / channel dimensions on x0 or y, and functions whose trace the reference rejects (batched coefficients, channel mixing):
    with the bridge on, the same x and iteration counts as with the reference's own assembly, and the same exceptions."""
    import torch
    from phiml import math
    from phiml.math import spatial, batch, channel, extrapolation as E_, Solve, SolveTape
    from phiml.backend import default_backend
    import phiml.math._lin_trace as lt
    from phiml_b200 import assembly
    from oracle.assemble import stencil_csr, shift_values_csr
    TORCH = phiml_env[1]

    def oa(g, offsets, coeffs, dtype):
        csr = stencil_csr(g, list(zip(offsets, coeffs)), dtype.type)
        return torch.as_tensor(csr.indptr), torch.as_tensor(csr.indices), torch.as_tensor(csr.data)

    def oag(g, offsets, values, dtype):
        csr = shift_values_csr(g, offsets, values)
        return torch.as_tensor(csr.indptr), torch.as_tensor(csr.indices), torch.as_tensor(csr.data)

    k = math.tensor(np.array([1.0, 2.0]), batch('k'))
    lap = lambda: (lambda x: -math.laplace(x, padding=E_.ZERO) + 0.1 * x)
    cases = {
        'batched_x0': (lap, lambda: math.zeros(batch(b=2), spatial(x=6, y=5)), lambda r: math.tensor(r.standard_normal((2, 6, 5)), batch('b'), spatial('x,y')), 'ok'),
        'channel_x0': (lap, lambda: math.zeros(spatial(x=6, y=5), channel(c=2)), lambda r: math.tensor(r.standard_normal((6, 5, 2)), spatial('x,y'), channel('c')), 'ok'),
        'y_batch_only': (lap, lambda: math.zeros(spatial(x=6, y=5)), lambda r: math.tensor(r.standard_normal((3, 6, 5)), batch('b'), spatial('x,y')), 'ok'),
        'batched_coefficient': (lambda: (lambda x: -math.laplace(x, padding=E_.ZERO) + k * x), lambda: math.zeros(spatial(x=6, y=5)), lambda r: math.tensor(r.standard_normal((6, 5)), spatial('x,y')), 'raised'),
        'channel_mixing': (lambda: (lambda x: -math.laplace(x, padding=E_.ZERO) + math.sum(x, 'c')), lambda: math.zeros(spatial(x=6, y=5), channel(c=2)), lambda r: math.tensor(r.standard_normal((6, 5, 2)), spatial('x,y'), channel('c')), 'raised'),
    }
    bridged = lt.tracer_to_coo
    original = phiml_env[0]._INSTALLED['original_t2c']
    bridge = assembly.make_tracer_to_coo(original, lambda: default_backend() is TORCH, oa, oag)
    try:
        for name, (mkf, mkx0, mky, expect) in cases.items():
            out = []
            for t2c in (original, bridge):
                lt.tracer_to_coo = t2c
                assembly.STATS.update(device_assemblies=0, general_assemblies=0, fallbacks=0)
                f = math.jit_compile_linear(mkf())
                try:
                    with TORCH, math.precision(64):
                        x0, y = mkx0(), mky(np.random.default_rng(0))
                        solve = Solve('CG', 1e-9, 0, x0=x0, max_iterations=500)
                        with SolveTape() as tape:
                            x = math.solve_linear(f, y, solve)
                        its = tape[solve].iterations
                        out.append(('ok', np.asarray(x.numpy(x.shape.names)), its.numpy(its.shape.names).tolist()))
                except Exception as exc:
                    out.append(('raised', type(exc).__name__, None))
            (s0, a, i0), (s1, b, i1) = out
            assert s0 == s1 == expect, (name, out)
            if s0 == 'ok':
                assert a.shape == b.shape and np.allclose(a, b, rtol=1e-8, atol=1e-9) and i0 == i1, name
                assert assembly.STATS['device_assemblies'] == 1, (name, assembly.STATS)
            else:
                assert a == b, (name, a, b)
    finally:
        lt.tracer_to_coo = bridged
