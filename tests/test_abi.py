"""CPU suite, part 2: the C-ABI library builds/loads without a GPU and exports exactly what include/phb200.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    text = open(os.path.join(ROOT, 'include', 'phb200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(phb200_[a-z0-9_]+)\s*\(', text)))


def test_library_exports_every_declared_symbol():
    from phiml_b200 import _engine as E
    lib = E.load()
    names = header_functions()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"libphb200.so lacks {n} declared in include/phb200.h"
    assert sorted(E.SIGNATURES) == names, "phiml_b200/_engine.py prototypes out of sync with include/phb200.h"
    assert lib.phb200_version() == 100
    assert lib.phb200_launch_count() == 0


def test_argument_errors_do_not_need_a_gpu():
    from phiml_b200 import _engine as E
    lib = E.load()
    h = ctypes.c_void_p()
    rc = lib.phb200_matrix_csr(ctypes.byref(h), None, None, None, 4, 4, 0, E.F32, 0)
    assert rc == E.ERR_ARG and b'null' in lib.phb200_last_error()
    with pytest.raises(ValueError):
        E.check(rc)
    g = (ctypes.c_int64 * 2)(4, 4)
    off = (ctypes.c_int32 * 4)(0, 0, 0, 0)
    co = (ctypes.c_double * 2)(1.0, 2.0)
    assert lib.phb200_matrix_stencil(ctypes.byref(h), 2, g, 2, off, co, E.F64) == E.ERR_ARG      # duplicate offsets
    off = (ctypes.c_int32 * 4)(0, 0, 0, 1)
    assert lib.phb200_matrix_stencil(ctypes.byref(h), 2, g, 2, off, co, E.F64) == E.OK
    nnz = ctypes.c_int64()
    assert lib.phb200_stencil_nnz(h, ctypes.byref(nnz)) == E.OK and nnz.value == 16 + 12
    assert lib.phb200_matrix_destroy(h) == E.OK


def test_product_path_fails_loudly_without_cuda():
    """No CPU fallback: host tensors are rejected instead of being routed somewhere else."""
    import torch
    import phiml_b200
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    a = torch.sparse_csr_tensor(torch.tensor([0, 1, 2]), torch.tensor([0, 1]), torch.tensor([1., 2.]), (2, 2))
    with pytest.raises(phiml_b200.EngineError):
        phiml_b200.mul_matrix_batched_vector(a, torch.ones(1, 2))
    with pytest.raises(phiml_b200.EngineError):
        phiml_b200.linear_solve('CG', a, torch.ones(1, 2), torch.zeros(1, 2), [1e-5], [0], [[10]])


def test_missing_library_raises(tmp_path):
    from phiml_b200 import _engine as E
    with pytest.raises(E.EngineError):
        E.load(str(tmp_path / 'nope.so'))


def test_workload_stencil_tables_equal_the_oracle_tables():
    """bench.py / scripts take their operators from phiml_b200.stencils (product side); same points, order and coefficient bits
    as the oracle's tables, which are pinned to the reference's assembly (tests/test_oracle_golden.py::test_assembly_bit_exact)."""
    import numpy as np
    from phiml_b200 import stencils
    from oracle.assemble import advdiff_shifts, laplace_shifts
    for rank in (1, 2, 3):
        assert stencils.neg_laplace_pairs(rank) == [(s, float(c)) for s, c in laplace_shifts(rank)]
        for dt in (np.float32, np.float64):
            ours = stencils.advection_diffusion_pairs(rank, stencils.ADVDIFF['velocity'], stencils.ADVDIFF['c'], stencils.ADVDIFF['nu'], dt)
            assert ours == [(s, float(c)) for s, c in advdiff_shifts(rank, (1.0, 0.5, 0.25), 0.3, 0.2, dt)]
    from _cases import ADVDIFF
    assert ADVDIFF == stencils.ADVDIFF
