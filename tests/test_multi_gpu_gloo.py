"""
CPU suite, part 4: the N>1 host logic with world_size 2 over gloo — shard / slab bounds, halo exchange pattern, the
collective sequence of SlabCG and the batch-minimum tolerance of sharded solves.  The CUDA engine is replaced by the CPU test
double tests/_slab_engine.py; the result is compared with the single-process oracle on the whole system.
"""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))


def test_bounds():
    from phiml_b200.multi_gpu import shard_bounds, slab_bounds
    for total, world in [(4096, 8), (10, 3), (7, 7), (5, 8)]:
        parts = [shard_bounds(total, world, r) for r in range(world)]
        assert parts[0][0] == 0 and parts[-1][1] == total
        assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
        assert max(b - a for a, b in parts) - min(b - a for a, b in parts) <= 1
    with pytest.raises(ValueError):
        slab_bounds(3, 4, 3)


def _worker(rank, world, port, grid, out_dir):
    sys.path.insert(0, os.path.dirname(HERE))
    sys.path.insert(0, HERE)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from phiml_b200.multi_gpu import SlabCG, shard_bounds
        from _slab_engine import CpuSlabEngine
        from oracle.assemble import laplace_shifts
        shifts = laplace_shifts(3)
        offsets, coeffs = [s for s, _ in shifts], [c for _, c in shifts]
        dtype = torch.float64
        cg = SlabCG(grid, offsets, coeffs, dtype, jacobi=True, group=None,
                    engine_factory=lambda g, xb, xe: CpuSlabEngine(g, xb, xe, offsets, coeffs, dtype, True))
        n = int(np.prod(grid))
        y = np.random.default_rng(0).standard_normal((2, n))
        y[1] *= 1e-3                                            # different magnitudes: exercises the minimum-tolerance rule
        plane = grid[1] * grid[2]
        y_loc = torch.as_tensor(y[:, cg.x_lo * plane: cg.x_hi * plane].copy())
        x, r, its, fev, conv, div = cg.solve(y_loc, torch.zeros_like(y_loc), [1e-8, 1e-8], [0.0, 0.0], 500, poll_every=7)
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), x=x.numpy(), its=its.numpy(), fev=fev.numpy(), conv=conv.numpy(), lo=cg.x_lo, hi=cg.x_hi)
        # batch sharding: the global minimum tolerance is what one process would compute over the whole batch
        lo, hi = shard_bounds(5, world, rank)
        tol_all = torch.tensor([3e-4, 1e-6, 2e-5, 7e-7, 9e-3], dtype=torch.float64)
        t = tol_all[lo:hi].min().reshape(1)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        assert float(t) == float(tol_all.min())
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_slab_cg_two_ranks_matches_single_process_oracle(tmp_path):
    from oracle import krylov, precond
    from oracle.assemble import poisson_csr
    grid = (7, 6, 8)
    port = 29600 + (os.getpid() % 300)
    mp.spawn(_worker, args=(2, port, grid, str(tmp_path)), nprocs=2, join=True)
    parts = [np.load(os.path.join(tmp_path, f"rank{r}.npz")) for r in range(2)]
    assert parts[0]['lo'] == 0 and parts[0]['hi'] == parts[1]['lo'] and parts[1]['hi'] == grid[0]
    x = np.concatenate([p['x'] for p in parts], axis=1)
    csr = poisson_csr(grid, np.float64)
    n = csr.shape[0]
    y = np.random.default_rng(0).standard_normal((2, n))
    y[1] *= 1e-3
    ref = krylov.cg(csr, y, np.zeros_like(y), np.float64([1e-8, 1e-8]), np.float64([0, 0]), np.array([[500, 500]]), precond.Jacobi(csr))
    for p in parts:                                             # scalars identical on every rank
        np.testing.assert_array_equal(p['its'], parts[0]['its'])
        assert p['conv'].all()
    assert np.max(np.abs(parts[0]['its'].astype(int) - ref.iterations.astype(int))) <= 1
    np.testing.assert_array_equal(parts[0]['fev'], parts[0]['its'] + 1)
    scale = np.abs(ref.x).max(axis=1, keepdims=True)
    assert np.max(np.abs(x - ref.x) / scale) < 1e-7      # rel_tol 1e-8; the stopping iteration may differ by one


def _worker_any_method(rank, world, port, grid, method, jacobi, out_dir):
    sys.path.insert(0, os.path.dirname(HERE))
    sys.path.insert(0, HERE)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from phiml_b200.multi_gpu import SlabSolver
        from _slab_engine import CpuCallbackEngine
        from _cases import ADVDIFF
        from oracle.assemble import advdiff_shifts
        shifts = advdiff_shifts(3, ADVDIFF['velocity'], ADVDIFF['c'], ADVDIFF['nu'], np.float64)
        offsets, coeffs = [s for s, _ in shifts], [float(c) for _, c in shifts]
        dtype = torch.float64
        sol = SlabSolver(grid, offsets, coeffs, dtype, method=method, jacobi=jacobi, group=None,
                         engine_factory=lambda g, xb, xe: CpuCallbackEngine(g, xb, xe, offsets, coeffs, dtype, method, jacobi))
        n = int(np.prod(grid))
        y = np.random.default_rng(3).standard_normal((2, n))
        plane = grid[1] * grid[2]
        y_loc = torch.as_tensor(y[:, sol.x_lo * plane: sol.x_hi * plane].copy())
        x, r, its, fev, conv, div = sol.solve(y_loc, torch.zeros_like(y_loc), [1e-10, 1e-10], [0.0, 0.0], 500)
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), x=x.numpy(), r=r.numpy(), its=its.numpy(), fev=fev.numpy(), conv=conv.numpy(), div=div.numpy(),
                 lo=sol.x_lo, hi=sol.x_hi, n_allreduce=sol.counts['allreduce'], n_halo=sol.counts['halo'])
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
@pytest.mark.parametrize('method,jacobi', [('biCG-stab(2)', False), ('biCG-stab', True), ('CG-adaptive', False)])
def test_slab_solver_two_ranks_any_method_matches_single_process_oracle(tmp_path, method, jacobi):
    """BASELINE config 5 in miniature: nonsymmetric advection-diffusion, row-partitioned over 2 ranks, halo exchange before
    every product, all-reduce for every dot product; the result must be the single-process solve."""
    from oracle import krylov, precond
    from _util import oracle_matrix
    grid = (9, 6, 8)
    port = 29900 + (os.getpid() % 300) + (7 if jacobi else 0) + len(method)
    mp.spawn(_worker_any_method, args=(2, port, grid, method, jacobi, str(tmp_path)), nprocs=2, join=True)
    parts = [np.load(os.path.join(tmp_path, f"rank{r}.npz")) for r in range(2)]
    assert parts[0]['lo'] == 0 and parts[0]['hi'] == parts[1]['lo'] and parts[1]['hi'] == grid[0]
    x = np.concatenate([p['x'] for p in parts], axis=1)
    csr = oracle_matrix('advdiff', grid, 'float64')
    n = csr.shape[0]
    y = np.random.default_rng(3).standard_normal((2, n))
    ref = krylov.linear_solve(method, csr, y, np.zeros_like(y), np.float64([1e-10, 1e-10]), np.float64([0, 0]), np.array([[500, 500]]),
                              precond.Jacobi(csr) if jacobi else None, flavour='numpy')
    for p in parts:
        np.testing.assert_array_equal(p['its'], parts[0]['its'])
        np.testing.assert_array_equal(p['fev'], parts[0]['fev'])
        assert p['conv'].all() and not p['div'].any()
        assert p['n_allreduce'] == parts[0]['n_allreduce'] and p['n_halo'] == parts[0]['n_halo'] and p['n_halo'] >= int(parts[0]['fev'].max())
    assert np.max(np.abs(parts[0]['its'].astype(int) - ref.iterations.astype(int))) <= 1
    scale = np.abs(ref.x).max(axis=1, keepdims=True)
    assert np.max(np.abs(x - ref.x) / scale) < 1e-8
