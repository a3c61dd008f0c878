"""
phiml_b200 — B200-native (sm_100a) drop-in for the sparse linear-solve hot path of PhiML's torch backend.

Only what the path needs: the C-ABI CUDA library (csrc/ -> lib/libphb200.so, header include/phb200.h), a host-side
mirror of the reference's Backend operator interface (solve.py, backend.py, precond.py, sparse.py), the PhiML plug-in
(phiml_plugin.py) and the multi-GPU drivers (multi_gpu.py).  No CPU fallback anywhere.
"""
from ._engine import EngineError, load as load_library, launch_count
from .sparse import Matrix, as_matrix, clear_cache
from .precond import Preconditioner, JacobiPreconditioner, IncompleteLU, ClusterPreconditioner, compute_preconditioner
from .solve import (SolveResult, linear_solve, conjugate_gradient, conjugate_gradient_adaptive, bi_conjugate_gradient,
                    generic_conjugate_gradient, generic_conjugate_gradient_adaptive, mul_matrix_batched_vector, mul_csr_dense, mul_coo_dense,
                    solve_triangular_sparse, solve_triangular, DeviceSolver)
from .backend import B200Backend
from . import adjoint

__all__ = [n for n in dir() if not n.startswith('_')]
