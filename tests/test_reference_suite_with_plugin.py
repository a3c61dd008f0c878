"""
CPU suite: the reference's OWN tests for the path (tests/commit/math/test__optimize.py, test__sparse.py, test__trace.py, test__functional.py and
tests/commit/backend/test__backend.py of PhiML 1.15.1) must keep passing with the phb200 plug-in installed — the plug-in re-binds eight things inside
PhiML and none of them may change what CPU operands, NumPy-backed matrices or the other backends see.  Runs only where the reference's
test files exist (/root/reference; the GPU box carries the package alone), in a child process so that the patched singleton does not leak.
"""
import os
import subprocess
import sys

import pytest

REF = '/root/reference'
FILES = ['math/test__optimize.py', 'math/test__sparse.py', 'math/test__trace.py', 'math/test__functional.py', 'backend/test__backend.py']

RUNNER = r'''
import sys
sys.path.insert(0, %(ref)r)
sys.path.insert(0, %(root)r)
import phiml
from phiml_b200 import phiml_plugin
TORCH = phiml_plugin.install()
assert type(TORCH).__name__ == 'B200TorchBackend'
import pytest
sys.exit(int(pytest.main(['-q', '--no-header', '-p', 'no:cacheprovider', '--rootdir=' + %(tmp)r, '-W', 'ignore'] + %(files)r)))
'''


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, 'tests', 'commit', 'math')), reason="the reference's test files are not on this machine")
def test_the_references_own_path_tests_pass_with_the_plugin_installed(tmp_path):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    files = [os.path.join(REF, 'tests', 'commit', f) for f in FILES]
    code = RUNNER % dict(ref=REF, root=root, tmp=str(tmp_path), files=files)
    r = subprocess.run([sys.executable, '-c', code], cwd=str(tmp_path), capture_output=True, text=True, timeout=900)
    tail = '\n'.join((r.stdout + r.stderr).splitlines()[-15:])
    assert r.returncode == 0, tail
    assert ' passed' in r.stdout and 'failed' not in r.stdout.splitlines()[-1], tail


BRIDGED_RUNNER = r'''
import sys
sys.path.insert(0, %(ref)r)
sys.path.insert(0, %(root)r)
import torch
import phiml
from phiml_b200 import phiml_plugin, assembly
TORCH = phiml_plugin.install()
import phiml.math._lin_trace as lt
from phiml.backend import default_backend
from oracle.assemble import stencil_csr, shift_values_csr

def assemble(grid, offsets, coeffs, dtype):
    csr = stencil_csr(grid, list(zip(offsets, coeffs)), dtype.type)
    return torch.as_tensor(csr.indptr), torch.as_tensor(csr.indices), torch.as_tensor(csr.data)

def assemble_general(grid, offsets, values, dtype):
    csr = shift_values_csr(grid, offsets, values)
    return torch.as_tensor(csr.indptr), torch.as_tensor(csr.indices), torch.as_tensor(csr.data)

original = phiml_plugin._INSTALLED['original_t2c']
lt.tracer_to_coo = assembly.make_tracer_to_coo(original, lambda: default_backend() is TORCH, assemble, assemble_general)
import pytest
rc = int(pytest.main(['-q', '--no-header', '-p', 'no:cacheprovider', '--rootdir=' + %(tmp)r, '-W', 'ignore'] + %(files)r))
print('BRIDGED', assembly.STATS['device_assemblies'] + assembly.STATS['general_assemblies'])
sys.exit(rc)
'''


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, 'tests', 'commit', 'math')), reason="the reference's test files are not on this machine")
def test_the_references_own_tests_pass_on_bridged_matrices(tmp_path):
    """The reference's whole tests/commit/math directory (+ backend/test__backend.py) with the tracer bridge switched ON whenever PhiML's default backend is torch — what a CUDA
    machine does — and the CSR arrays coming from the oracle instead of the device assembler (no GPU here): every matrix_from_function /
    jit_compile_linear matrix the reference's tests build under torch is then a bridged CompressedSparseMatrix, and their expectations must still hold."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    files = [os.path.join(REF, 'tests', 'commit', 'math'), os.path.join(REF, 'tests', 'commit', 'backend', 'test__backend.py')]      # 400 tests
    code = BRIDGED_RUNNER % dict(ref=REF, root=root, tmp=str(tmp_path), files=files)
    r = subprocess.run([sys.executable, '-c', code], cwd=str(tmp_path), capture_output=True, text=True, timeout=900)
    tail = '\n'.join((r.stdout + r.stderr).splitlines()[-15:])
    assert r.returncode == 0, tail
    bridged = [int(l.split()[1]) for l in r.stdout.splitlines() if l.startswith('BRIDGED')]
    assert bridged and bridged[0] >= 10, f"the bridge was taken {bridged} times\n{tail}"
