"""
PhiML plug-in: makes the phb200 path the implementation behind PhiML's *own* torch backend object, so that unmodified
``phiml.math`` code (named batch dims, ``SolveTape``, ``NotConverged``/``Diverged``) runs on it.

    import phiml_b200.phiml_plugin as plugin
    plugin.install()          # once, after importing phiml

What it does:

* ``TORCH.__class__`` is switched to a subclass of ``TorchBackend`` (phiml/backend/torch/_torch_backend.py:19) that overrides
  ``linear_solve`` / ``conjugate_gradient`` / ``conjugate_gradient_adaptive`` / ``bi_conjugate_gradient``,
  ``mul_matrix_batched_vector``, ``mul_csr_dense``, ``mul_coo_dense`` and adds ``solve_triangular_sparse`` — the singleton identity is
  kept (``get_backend('torch') is TORCH``) and ``TORCH.supports(Backend.solve_triangular_sparse)`` turns true, so
  ``preconditioner='auto'`` selects ILU (phiml/math/_optimize.py:842-846).
* ``phiml.math._optimize.compute_preconditioner`` (looked up as a module global at :644) is re-bound to a wrapper that adds
  ``'jacobi'`` and builds ``'ilu'`` on the device for CUDA torch matrices.
* Device assembly: ``phiml.math._lin_trace.tracer_to_coo`` (looked up as a module global by
  ``matrix_from_function``, :737) is re-bound to the bridge of ``phiml_b200/assembly.py``: a traced constant-coefficient
  ZERO-boundary operator becomes the reference's own ``CompressedSparseMatrix`` with device-resident, byte-identical CSR arrays
  without the host-side COO build (36.7 s / 17.5 GB at 256^3, infeasible from 512^3 on).
* Residency: ``phiml.math._optimize.native_matrix`` (:650) and the preconditioner wrapper keep the device
  native / the factorisation of a matrix for as long as the PhiML matrix tensor itself lives (``LinearFunction`` caches that tensor
  per signature, so repeated ``solve_linear`` calls with one ``jit_compile_linear`` function skip the per-call upload — 0.94 GB at
  256^3 — and the per-call ILU factorisation that the reference repeats, ``_optimize.py:644,650``).

Calls whose operands are not CUDA torch tensors / torch sparse matrices (CPU tensors, dense matrices, matrix-free functions,
``scipy-*`` methods, tracing) are outside the phb200 path and go to the reference's own ``TorchBackend`` implementation unchanged.
"""
import warnings

import numpy as np
import torch

from . import precond as _precond
from . import solve as _solve
from .sparse import Matrix

_INSTALLED = {}


class _IdentityCache:
    """value per (live object, extra key): entries die with the object (weak reference), ids are never trusted alone."""

    def __init__(self, max_entries=8):
        self._d = {}
        self.max_entries = max_entries
        self.hits = self.misses = 0

    def get(self, obj, extra, make, pins=()):
        """``pins``: objects whose ids are part of ``extra``; the entry holds them so that their ids cannot be recycled under it."""
        import weakref
        key = (id(obj), extra)
        hit = self._d.get(key)
        if hit is not None and hit[0]() is obj:
            self.hits += 1
            return hit[1]
        self.misses += 1
        value = make()
        try:
            ref = weakref.ref(obj, lambda _r, k=key: self._d.pop(k, None))
        except TypeError:
            return value                      # not weak-referenceable: no caching
        if len(self._d) >= self.max_entries:
            self._d.pop(next(iter(self._d)))
        self._d[key] = (ref, value, pins)
        return value

    def clear(self):
        self._d.clear()


NATIVE_CACHE = _IdentityCache()
PRECOND_CACHE = _IdentityCache()


def _is_sparse_cuda(m) -> bool:
    if isinstance(m, Matrix):
        return True
    return isinstance(m, torch.Tensor) and m.layout != torch.strided and m.is_cuda


def _on_path(lin, y) -> bool:
    """True if this call belongs to the phb200 path: explicit sparse matrix and right-hand sides on a CUDA device, eager mode."""
    if not _is_sparse_cuda(lin):
        return False
    if torch._C._get_tracing_state() is not None:
        return False
    return isinstance(y, torch.Tensor) and y.is_cuda


def _own_pre(pre) -> bool:
    return pre is None or isinstance(pre, (_precond.JacobiPreconditioner, _precond.IncompleteLU))


def install():
    """Idempotent.  Returns the patched ``TORCH`` backend object."""
    if _INSTALLED:
        return _INSTALLED['torch']
    from phiml.backend import _backend as pb_backend
    from phiml.backend._backend import SolveResult as PhiSolveResult, Preconditioner as PhiPreconditioner
    from phiml.backend.torch import TORCH
    from phiml.backend.torch._torch_backend import TorchBackend
    import phiml.math._optimize as opt

    def convert(r: _solve.SolveResult) -> PhiSolveResult:
        return PhiSolveResult(r.method, r.x, r.residual, r.iterations, r.function_evaluations, r.converged, r.diverged, r.message)

    class B200TorchBackend(TorchBackend):
        """TorchBackend whose sparse linear-solve path runs on libphb200 (CUDA operands only)."""

        # --- products ---
        def mul_matrix_batched_vector(self, A, b):
            if _is_sparse_cuda(A) and isinstance(b, torch.Tensor) and b.is_cuda and torch._C._get_tracing_state() is None:
                A_, b_ = self.auto_cast(A, b) if isinstance(A, torch.Tensor) else (A, b)
                return _solve.mul_matrix_batched_vector(A_, b_)
            return TorchBackend.mul_matrix_batched_vector(self, A, b)

        def mul_csr_dense(self, column_indices, row_pointers, values, shape, dense):
            values_t, dense_t = self.as_tensor(values), self.as_tensor(dense)
            if values_t.is_cuda and dense_t.is_cuda and torch._C._get_tracing_state() is None:
                values_t, dense_t = self.auto_cast(values_t, dense_t, bool_to_int=True, int_to_float=True)
                return _solve.mul_csr_dense(self.as_tensor(column_indices), self.as_tensor(row_pointers), values_t, shape, dense_t)
            return TorchBackend.mul_csr_dense(self, column_indices, row_pointers, values, shape, dense)

        def mul_coo_dense(self, indices, values, shape, dense):
            values_t, dense_t = self.as_tensor(values), self.as_tensor(dense)
            if values_t.is_cuda and dense_t.is_cuda and torch._C._get_tracing_state() is None:
                values_t, dense_t = self.auto_cast(values_t, dense_t)
                return _solve.mul_coo_dense(self.as_tensor(indices), values_t, shape, dense_t)
            return TorchBackend.mul_coo_dense(self, indices, values, shape, dense)

        # --- solves ---
        def conjugate_gradient(self, lin, y, x0, rtol, atol, max_iter, pre, matrix_offset):
            if matrix_offset is None and _on_path(lin, y) and _own_pre(pre):
                return convert(_solve.conjugate_gradient(lin, self.to_float(y), self.to_float(x0), rtol, atol, max_iter, pre, None))
            return TorchBackend.conjugate_gradient(self, lin, y, x0, rtol, atol, max_iter, pre, matrix_offset)

        def conjugate_gradient_adaptive(self, lin, y, x0, rtol, atol, max_iter, pre, matrix_offset):
            if matrix_offset is None and _on_path(lin, y) and _own_pre(pre):
                return convert(_solve.conjugate_gradient_adaptive(lin, self.to_float(y), self.to_float(x0), rtol, atol, max_iter, pre, None))
            return TorchBackend.conjugate_gradient_adaptive(self, lin, y, x0, rtol, atol, max_iter, pre, matrix_offset)

        def bi_conjugate_gradient(self, lin, y, x0, rtol, atol, max_iter, pre, matrix_offset, poly_order=2):
            if matrix_offset is None and poly_order >= 1 and poly_order <= 4 and _on_path(lin, y) and _own_pre(pre):
                return convert(_solve.bi_conjugate_gradient(lin, self.to_float(y), self.to_float(x0), rtol, atol, max_iter, pre, None, poly_order))
            return TorchBackend.bi_conjugate_gradient(self, lin, y, x0, rtol, atol, max_iter, pre, matrix_offset, poly_order)

        def solve_triangular_sparse(self, matrix, rhs, lower: bool, unit_diagonal: bool):
            if _is_sparse_cuda(matrix) and isinstance(rhs, torch.Tensor) and rhs.is_cuda:
                return _solve.solve_triangular_sparse(matrix, rhs, lower, unit_diagonal)
            raise NotImplementedError("sparse triangular solves run on CUDA torch tensors only (phb200); use the NumPy backend on the CPU")

    # the device-side preconditioners must be accepted wherever PhiML checks the protocol type
    for cls in (_precond.JacobiPreconditioner, _precond.IncompleteLU):
        if PhiPreconditioner not in cls.__mro__:
            try:
                cls.__bases__ = cls.__bases__ + (PhiPreconditioner,)
            except TypeError:
                pass

    original_compute = opt.compute_preconditioner
    original_native = opt.native_matrix

    def anchor(matrix):
        """(object whose life time bounds the cache entry, key part) of a constant sparse matrix tensor, or (None, None).
        PhiML re-wraps the matrix tensor on every call (custom-gradient plumbing); the natives below the wrappers are stable."""
        if not getattr(matrix, 'available', False) or torch._C._get_tracing_state() is not None:
            return None, None, ()
        try:
            values = matrix._values._native
            indices = matrix._indices._native
            pointers = getattr(getattr(matrix, '_pointers', None), '_native', None)
        except AttributeError:
            return None, None, ()
        if isinstance(values, torch.Tensor) and values.requires_grad:
            return None, None, ()
        # which dims are compressed / how the coordinate columns are labelled distinguishes A from the transposed view the backward
        # pass builds by renaming dims over the SAME arrays (transpose_matrix, _optimize.py:823-833)
        layout = (repr(getattr(matrix, '_compressed_dims', None)), repr(getattr(matrix, '_uncompressed_dims', None)),
                  repr(getattr(matrix, '_dense_shape', None)), repr(matrix._indices.shape))
        # in-place edits of the arrays (an optimizer step under no_grad, .copy_) bump torch's version counters: a stale native or
        # stale factors must not be served for them
        versions = tuple(getattr(t, '_version', None) for t in (values, indices, pointers))
        return values, (type(matrix).__name__, id(indices), id(pointers), repr(matrix.shape), versions) + layout, (indices, pointers)

    def native_matrix(value, target_backend):
        """phiml.math._sparse.native_matrix with residency: one conversion / upload per live matrix and backend."""
        obj, key, pins = anchor(value)
        if target_backend is None or obj is None:
            return original_native(value, target_backend)
        return NATIVE_CACHE.get(obj, (target_backend.name,) + key, lambda: original_native(value, target_backend), pins)

    def compute_preconditioner(method, matrix, rank_deficiency=0, target_backend=None, solver=None):
        from phiml.math._sparse import is_sparse
        cuda_torch = target_backend is TORCH and is_sparse(matrix) and torch.cuda.is_available() and str(TORCH.get_default_device().device_type).upper() == 'GPU'
        if method == 'auto' and cuda_torch and solver not in ('direct', 'scipy-direct') and not (solver or '').startswith('scipy-'):
            # the reference resolves 'auto' INSIDE its own function (_optimize.py:838-854: 'ilu' whenever the backend has sparse
            # triangular solves, which the patched backend does) and would go on to its CPU sweeps; resolve it here instead
            method = 'ilu'
        if method == 'jacobi':
            if not cuda_torch:
                raise NotImplementedError("preconditioner='jacobi' is provided by phb200 for sparse matrices on the CUDA torch backend only")
            make = lambda: _precond.JacobiPreconditioner(native_matrix(matrix, TORCH))
            obj, key, pins = anchor(matrix)
            return PRECOND_CACHE.get(obj, ('jacobi',) + key, make, pins) if obj is not None else make()
        if method == 'ilu' and cuda_torch and not (solver or '').startswith('scipy-'):
            rd = np.asarray(rank_deficiency.numpy() if hasattr(rank_deficiency, 'numpy') else rank_deficiency)
            safe = bool(np.any(rd > 0))                     # factor_ilu(matrix, iterations, safe=(rank_deficiency > 0).any), _optimize.py:869
            make = lambda: _precond.IncompleteLU(native_matrix(matrix, TORCH), safe=safe, rank_deficiency=rd)
            obj, key, pins = anchor(matrix)
            return PRECOND_CACHE.get(obj, ('ilu', safe) + key, make, pins) if obj is not None else make()
        return original_compute(method, matrix, rank_deficiency=rank_deficiency, target_backend=target_backend, solver=solver)

    # constant-coefficient ZERO-boundary operators go from the tracer straight to the device assembler
    import phiml.math._lin_trace as lin_trace
    from . import assembly
    original_t2c = lin_trace.tracer_to_coo

    def wants_device_matrix() -> bool:
        from phiml.backend import default_backend
        return (default_backend() is TORCH and torch.cuda.is_available() and str(TORCH.get_default_device().device_type).upper() == 'GPU'
                and torch._C._get_tracing_state() is None)

    def assemble(grid, offsets, coeffs, dtype):
        return assembly.cuda_assemble(grid, offsets, coeffs, dtype, TORCH.get_default_device().ref)

    lin_trace.tracer_to_coo = assembly.make_tracer_to_coo(original_t2c, wants_device_matrix, assemble)

    TORCH.__class__ = B200TorchBackend
    opt.compute_preconditioner = compute_preconditioner
    opt.native_matrix = native_matrix
    _INSTALLED.update(torch=TORCH, original_class=TorchBackend, original_compute=original_compute, original_native=original_native, opt=opt,
                      lin_trace=lin_trace, original_t2c=original_t2c)
    return TORCH


def uninstall():
    if not _INSTALLED:
        return
    _INSTALLED['torch'].__class__ = _INSTALLED['original_class']
    _INSTALLED['opt'].compute_preconditioner = _INSTALLED['original_compute']
    _INSTALLED['opt'].native_matrix = _INSTALLED['original_native']
    _INSTALLED['lin_trace'].tracer_to_coo = _INSTALLED['original_t2c']
    NATIVE_CACHE.clear()
    PRECOND_CACHE.clear()
    _INSTALLED.clear()
