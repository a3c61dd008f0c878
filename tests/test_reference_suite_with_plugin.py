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
