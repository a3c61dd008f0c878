"""
GPU suite, multi-GPU part (needs >= 2 GPUs on the box, skipped otherwise): the slab drivers launched as the bench is —
one process per GPU under torchrun, NCCL for the plumbing — and compared with the single-GPU device loop on rank 0.
The slab decomposition changes neither the kernels nor the per-rank reduction order, and the cross-rank sums are added in rank
order on every rank, so the results must be IDENTICAL to one GPU: same iteration count, max |x_slab - x_1gpu| == 0.
"""
import json
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def run_ranks(world, script, *args, timeout=600):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs, the box has {torch.cuda.device_count()}")
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', f'--nproc-per-node={world}', '--master-addr', '127.0.0.1', '--master-port', str(_free_port()),
           os.path.join(ROOT, 'scripts', script), *args]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-3000:]
    lines = [l for l in p.stdout.splitlines() if l.startswith('{')]
    assert lines, p.stdout[-2000:] + p.stderr[-2000:]
    return json.loads(lines[-1])


@pytest.mark.parametrize('world', [2, 4])
@pytest.mark.parametrize('mode', ['peer', 'nccl'])
def test_slab_cg_is_bit_identical_to_one_gpu(world, mode):
    """Jacobi-CG 64^3 / 128^3 on x-slabs: peer-memory halo stores + in-kernel all-reduce (default) and the NCCL variants."""
    extra = [] if mode == 'peer' else ['--nccl-reduce', '--nccl-halos']
    for size in (64, 128):
        out = run_ranks(world, 'run_slab.py', '--size', str(size), '--check', *extra)
        assert out['converged'] and out['iterations'] == out['ref_iterations'], out
        assert out['rel_err'] == 0.0, out
        assert out['true_residual'] < 2e-5, out
        if mode == 'peer':
            assert out['all_reduce'].startswith('in-kernel') and out['halo_exchange'].startswith('peer-memory'), out


@pytest.mark.parametrize('method,op,jacobi', [('biCG-stab(2)', 'advdiff', False), ('biCG-stab', 'advdiff', True), ('CG-adaptive', 'poisson', False)])
def test_slab_any_method_is_bit_identical_to_one_gpu(method, op, jacobi):
    """BASELINE config 5 in miniature: every method of the device loop on x-slabs of ONE system over 2 GPUs."""
    args = ['--size', '64', '--check', '--method', method, '--op', op] + (['--jacobi'] if jacobi else [])
    out = run_ranks(2, 'run_slab_any.py', *args)
    assert out['converged'] and out['iterations'] == out['ref_iterations'], out
    assert out['rel_err_vs_single_gpu'] == 0.0, out
    assert out['true_residual'] < 5e-5, out


def test_sharded_batch_matches_one_gpu():
    """BASELINE config 3 in miniature: 256 systems of 128^2 sharded over 2 GPUs; iterations and x equal the unsharded solve."""
    out = run_ranks(2, 'run_sharded.py', '--batch', '256', '--check')
    assert out['converged'], out
    assert out['single_gpu_iterations_equal_on_shard'], out
    assert out['max_rel_err_vs_single_gpu'] <= 1e-6, out


def test_bench_multi_gpu_times_a_loop_with_a_collective():
    """bench.py --gpus 2: the headline is ONE 256^3 system on x-slabs (strong scaling), not independent replicas."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node=2', '--master-addr', '127.0.0.1', '--master-port', str(_free_port()),
           os.path.join(ROOT, 'bench.py'), '--gpus', '2', '--steps', '20', '--warmup', '5', '--extras', '0']
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-3000:]
    line = json.loads([l for l in p.stdout.splitlines() if l.startswith('{')][-1])
    assert line['n_gpus'] == 2 and line['scaling'] == 'strong' and line['value'] > 0
    assert line['headline_detail']['collectives_per_iteration']['all_reduce_in_kernel'] == 2
    assert line['e2e']['iterations_per_solve'] in range(551, 556)      # the reference's 553 +- 2


def test_operands_on_a_non_current_device():
    """scratch, set-up temporaries and launches follow the operands' device, not the caller's current device."""
    if torch.cuda.device_count() < 2:
        pytest.skip(f"needs 2 GPUs, the box has {torch.cuda.device_count()}")
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    import phiml_b200 as pb
    from _util import oracle_matrix
    csr = oracle_matrix('poisson', (24, 20, 28), 'float32')
    y = np.random.default_rng(0).standard_normal((2, csr.shape[0])).astype(np.float32)
    rt, at, mi = np.float32([1e-5, 1e-5]), np.float32([0, 0]), np.array([[1000, 1000]])
    results = []
    assert torch.cuda.current_device() == 0
    for dev in ('cuda:0', 'cuda:1'):
        nat = torch.sparse_csr_tensor(torch.as_tensor(csr.indptr, device=dev), torch.as_tensor(csr.indices, device=dev), torch.as_tensor(csr.data, device=dev), csr.shape)
        m = pb.as_matrix(nat)
        assert m.stencil is not None
        for pre in (pb.JacobiPreconditioner(m), pb.IncompleteLU(m)):
            yd = torch.as_tensor(y, device=dev)
            res = pb.generic_conjugate_gradient(m, yd, torch.zeros_like(yd), rt, at, mi, pre)
            assert res.x.device == torch.device(dev) and bool(res.converged.all())
            results.append((res.iterations.cpu(), res.x.cpu()))
        assert torch.cuda.current_device() == 0, "the caller's current device must be restored"
    for a, b in zip(results[:2], results[2:]):
        assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
