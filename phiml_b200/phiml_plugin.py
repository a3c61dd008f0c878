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
DM_STATS = {'device': 0, 'reference': 0}      # matrix gradients of backward passes computed by the kernel / by the reference's rule


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


def _device_tensor(t) -> bool:
    return isinstance(t, torch.Tensor) and t.is_cuda


def _traceable(lin, y) -> bool:
    """True while ``torch.jit.trace`` is recording (``math.jit_compile``) and the operands are a torch sparse matrix / tensors on CUDA:
    the solve then goes through the registered operator ``phb200::linear_solve`` and becomes one node of the traced graph."""
    if torch._C._get_tracing_state() is None:
        return False
    return isinstance(lin, torch.Tensor) and lin.layout != torch.strided and lin.is_cuda and isinstance(y, torch.Tensor) and y.is_cuda


def _own_pre(pre) -> bool:
    return pre is None or isinstance(pre, (_precond.JacobiPreconditioner, _precond.IncompleteLU, _precond.ClusterPreconditioner))


def install():
    """Idempotent.  Returns the patched ``TORCH`` backend object."""
    if _INSTALLED:
        return _INSTALLED['torch']
    from phiml.backend import _backend as pb_backend
    from phiml.backend._backend import SolveResult as PhiSolveResult, Preconditioner as PhiPreconditioner
    from phiml.backend.torch import TORCH
    from phiml.backend.torch._torch_backend import TorchBackend
    import phiml.math._optimize as opt
    from . import assembly

    def convert(r: _solve.SolveResult) -> PhiSolveResult:
        return PhiSolveResult(r.method, r.x, r.residual, r.iterations, r.function_evaluations, r.converged, r.diverged, r.message)

    def traced(kind, method, lin, y, x0, rtol, atol, max_iter, pre, poly_order=2):
        from . import torch_op
        parts = torch_op.traced_linear_solve(method, lin, y, x0, rtol, atol, max_iter, pre)
        name = _solve.method_string(kind, pre, np.asarray(max_iter).shape[0] > 1, poly_order, traced=True)
        return PhiSolveResult(name, *parts, [""] * int(y.shape[0]))

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
            if matrix_offset is None and _traceable(lin, y) and _own_pre(pre):
                return traced('CG', 'CG', lin, self.to_float(y), self.to_float(x0), rtol, atol, max_iter, pre)
            if matrix_offset is None and _on_path(lin, y) and _own_pre(pre):
                return convert(_solve.conjugate_gradient(lin, self.to_float(y), self.to_float(x0), rtol, atol, max_iter, pre, None))
            return TorchBackend.conjugate_gradient(self, lin, y, x0, rtol, atol, max_iter, pre, matrix_offset)

        def conjugate_gradient_adaptive(self, lin, y, x0, rtol, atol, max_iter, pre, matrix_offset):
            if matrix_offset is None and _traceable(lin, y) and _own_pre(pre):
                return traced('CG-adaptive', 'CG-adaptive', lin, self.to_float(y), self.to_float(x0), rtol, atol, max_iter, pre)
            if matrix_offset is None and _on_path(lin, y) and _own_pre(pre):
                return convert(_solve.conjugate_gradient_adaptive(lin, self.to_float(y), self.to_float(x0), rtol, atol, max_iter, pre, None))
            return TorchBackend.conjugate_gradient_adaptive(self, lin, y, x0, rtol, atol, max_iter, pre, matrix_offset)

        def bi_conjugate_gradient(self, lin, y, x0, rtol, atol, max_iter, pre, matrix_offset, poly_order=2):
            if matrix_offset is None and poly_order >= 1 and poly_order <= 4 and _traceable(lin, y) and _own_pre(pre):
                method = 'biCG-stab' if poly_order == 1 else f'biCG-stab({poly_order})'
                return traced('biCG-stab', method, lin, self.to_float(y), self.to_float(x0), rtol, atol, max_iter, pre, poly_order)
            if matrix_offset is None and poly_order >= 1 and poly_order <= 4 and _on_path(lin, y) and _own_pre(pre):
                return convert(_solve.bi_conjugate_gradient(lin, self.to_float(y), self.to_float(x0), rtol, atol, max_iter, pre, None, poly_order))
            return TorchBackend.bi_conjugate_gradient(self, lin, y, x0, rtol, atol, max_iter, pre, matrix_offset, poly_order)

        def solve_triangular_sparse(self, matrix, rhs, lower: bool, unit_diagonal: bool):
            if _is_sparse_cuda(matrix) and isinstance(rhs, torch.Tensor) and rhs.is_cuda:
                return _solve.solve_triangular_sparse(matrix, rhs, lower, unit_diagonal)
            raise NotImplementedError("sparse triangular solves run on CUDA torch tensors only (phb200); use the NumPy backend on the CPU")

    # the device-side preconditioners must be accepted wherever PhiML checks the protocol type
    for cls in (_precond.JacobiPreconditioner, _precond.IncompleteLU, _precond.ClusterPreconditioner):
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

        def ident(t):      # torch tensors by storage address (detached copies share it), anything else by object identity
            return ('t', int(t.data_ptr()), int(t.numel())) if isinstance(t, torch.Tensor) else ('o', id(t))
        obj = values
        if isinstance(values, torch.Tensor):
            # device-assembled constants arrive as detached copies (new objects over the same storage on every call): tie the
            # entry to the bridged matrix that owns the storage
            owner = assembly.owner_of(values)
            if owner is not None:
                obj = owner
        return obj, (type(matrix).__name__, ident(indices), ident(pointers), repr(matrix.shape), versions) + layout, (indices, pointers)

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
            sweeps = None                                   # eager rule: ceil(sqrt(d n^(1/d))), precond.ilu_sweeps
            if opt._TRACING_JIT and getattr(matrix, 'available', False):
                # "high-quality preconditioner when jit-compiling with constant matrix" (_optimize.py:863-865): ceil(n^(1/d)) sweeps
                import math as _m
                from phiml.math import dual as _dual
                from phiml.math._sparse import stored_values as _stored
                n_ = _dual(matrix).volume
                d_ = (_stored(matrix).shape.volume / n_ - 1) / 2
                if d_ >= 1:
                    sweeps = int(_m.ceil(n_ ** (1 / d_)))
            make = lambda: _precond.IncompleteLU(native_matrix(matrix, TORCH), sweeps=sweeps, safe=safe, rank_deficiency=rd)
            obj, key, pins = anchor(matrix)
            return PRECOND_CACHE.get(obj, ('ilu', safe, sweeps) + key, make, pins) if obj is not None else make()
        if method == 'cluster' and cuda_torch and not (solver or '').startswith('scipy-') and torch._C._get_tracing_state() is None:
            # ExplicitClusterSolve would be applied by ~8 backend ops per iteration from Python; here the parts go into the device
            # preconditioner that the Krylov loop applies itself.  Grid matrices: the parts are built on the device from the native
            # (precond.ClusterPreconditioner.from_matrix restates explicit_coarse / coarse_explicit_preconditioner_coo, _optimize.py:942-971,
            # _linalg.py:769-799 — the reference's own builder needs the COO form, which a device-assembled matrix never had and which
            # CompressedSparseMatrix.decompress cannot rebuild for rank > 1, _sparse.py:895-903); anything else: the reference builds them.
            def make():
                from phiml.math import spatial, instance
                axes = spatial(matrix)
                if axes and not instance(matrix):
                    return _precond.ClusterPreconditioner.from_matrix(native_matrix(matrix, TORCH), grid=tuple(int(v) for v in axes.sizes))
                ref = original_compute(method, matrix, rank_deficiency=rank_deficiency, target_backend=target_backend, solver=solver)
                dev = TORCH.get_default_device().ref
                t = lambda a: torch.as_tensor(a).to(dev)
                return _precond.ClusterPreconditioner(t(ref.inv_coarse_matrix), t(ref.clusters).reshape(-1), int(ref.cluster_count), t(ref.cluster_size))
            obj, key, pins = anchor(matrix)
            return PRECOND_CACHE.get(obj, ('cluster',) + key, make, pins) if obj is not None else make()
        return original_compute(method, matrix, rank_deficiency=rank_deficiency, target_backend=target_backend, solver=solver)

    # constant-coefficient ZERO-boundary operators go from the tracer straight to the device assembler
    import phiml.math._lin_trace as lin_trace
    from . import assembly
    original_t2c = lin_trace.tracer_to_coo

    def wants_device_matrix(target_backend=None) -> bool:
        # the backend the matrix is FOR is the one of the traced tensor (matrix_from_function: target_backend = backend_for(*tensors), _lin_trace.py:714):
        # NumPy-backed x0 / y are solved on the NumPy backend even while torch is the default backend, and need the reference's NumPy-backed matrix
        from phiml.backend import default_backend
        if not (torch.cuda.is_available() and str(TORCH.get_default_device().device_type).upper() == 'GPU' and torch._C._get_tracing_state() is None):
            return False
        return (target_backend is TORCH) if target_backend is not None else (default_backend() is TORCH)

    def assemble(grid, offsets, coeffs, dtype):
        return assembly.cuda_assemble(grid, offsets, coeffs, dtype, TORCH.get_default_device().ref)

    lin_trace.tracer_to_coo = assembly.make_tracer_to_coo(original_t2c, wants_device_matrix, assemble)

    # ---- backward pass: phiml.math._optimize.attach_gradient_solve is looked up as a module global by solve_linear (:670)
    original_attach = opt.attach_gradient_solve

    def device_matrix_gradient(matrix, like_dy, x):
        """Prepares ``dm`` of implicit_gradient_solve (_optimize.py:798-807) as ONE kernel (phb200_matrix_gradient_coo: both gathers, the
        product, the batch sum and the negation).  Returns ``compute(dy) -> dm`` or None when the operands are not on the phb200 path
        (then the reference's own rule runs).  ``like_dy`` has the dims of dy (the forward right-hand side)."""
        from phiml.math import dual, instance
        from phiml.math._sparse import CompressedSparseMatrix, SparseCoordinateTensor
        from phiml.math._tensors import reshaped_native, reshaped_tensor, Tensor as PhiTensor
        from phiml.math._tree import disassemble_tree, variable_attributes
        from phiml.math._magic_ops import value_attributes
        if torch._C._get_tracing_state() is not None:
            return None
        coo = matrix.decompress() if isinstance(matrix, CompressedSparseMatrix) else matrix
        if not isinstance(coo, SparseCoordinateTensor):
            return None

        def single_tensor(tree, attr_type):
            try:
                _, tensors = disassemble_tree(tree, cache=False, attr_type=attr_type)
            except Exception:
                return None
            return tensors[0] if len(tensors) == 1 and isinstance(tensors[0], PhiTensor) else None

        y_t, x_t = single_tensor(like_dy, value_attributes), single_tensor(x, variable_attributes)
        if y_t is None or x_t is None or x_t.default_backend is not TORCH:
            return None
        ind_batch, channels, indices, _values, shape = coo._native_coo_components(dual, matrix=True)
        entries = instance(coo._values)
        if not ind_batch.is_empty or not channels.is_empty or entries.rank != 1:
            return None
        row_names = coo._dense_shape.non_dual.names
        col_names = tuple(n.lstrip('~') for n in coo._dense_shape.dual.names)
        if any(n not in y_t.shape for n in row_names) or any(n not in x_t.shape for n in col_names):
            return None
        extra = y_t.shape.without(row_names) & x_t.shape.without(col_names)      # what the reference sums over (:806)
        if any(n in coo.shape for n in extra.names):
            return None
        x_n = reshaped_native(x_t, [extra, x_t.shape.only(col_names, reorder=True)], force_expand=True)
        if not _device_tensor(x_n):
            return None
        idx = torch.as_tensor(indices).to(x_n.device)
        rows_cols = torch.stack([idx[0, :, 0], idx[0, :, 1]]).to(torch.int64)

        def compute(dy):
            from . import adjoint
            dy_t = single_tensor(dy, value_attributes)
            dy_n = reshaped_native(dy_t, [extra, dy_t.shape.only(row_names, reorder=True)], force_expand=True)
            dt = torch.float64 if torch.float64 in (dy_n.dtype, x_n.dtype) else torch.float32
            native = torch.sparse_coo_tensor(rows_cols, torch.empty(rows_cols.shape[1], dtype=dt, device=x_n.device), shape)
            out = adjoint.matrix_gradient(native, dy_n.detach().to(dt), x_n.detach().to(dt))
            DM_STATS['device'] += 1
            return coo._with_values(reshaped_tensor(out, [entries], convert=False))
        return compute

    def attach_gradient_solve(forward_solve, auxiliary_args, matrix_adjoint):
        aux = [a for a in auxiliary_args.split(',') if a]
        if not matrix_adjoint:
            normal = original_attach(forward_solve, auxiliary_args, matrix_adjoint)
            if 'matrix' in aux:
                return normal
            # A device-assembled matrix (assembly.build_compressed) is torch-backed, so the reference would treat it as a differentiable
            # input (auxiliary only "if matrix.backend == NUMPY", _optimize.py:670) although it is the same constant the reference holds in
            # NumPy arrays.  Constants must stay auxiliary: that is what keeps jit-compiled solves (one tensor argument, the right-hand
            # side) and the per-matrix caches working.
            constant = original_attach(forward_solve, auxiliary_args + ',matrix', matrix_adjoint)

            def dispatch(y, solve, matrix, **kwargs):
                return (constant if getattr(matrix, '_phb200_device_assembled', False) else normal)(y, solve, matrix, **kwargs)
            return dispatch

        reference_rule = original_attach(forward_solve, auxiliary_args, matrix_adjoint)      # complete reference behaviour, used as it is

        def backward(fwd_args: dict, x, dx):
            compute_dm = device_matrix_gradient(fwd_args['matrix'], fwd_args['y'], x) if 'matrix' in fwd_args else None
            if compute_dm is None:
                DM_STATS['reference'] += 1
                return reference_rule.gradient(fwd_args, x, dx)
            # the steps of implicit_gradient_solve (_optimize.py:786-797): adjoint solve on the transposed matrix with Solve.gradient_solve
            solve, matrix = fwd_args['solve'], fwd_args['matrix']
            transposed = opt.transpose_matrix(matrix, solve.x0.shape, fwd_args['y'].shape)
            grad_solve = solve.gradient_solve
            grad_solve = opt.copy_with(grad_solve, x0=grad_solve.x0 if grad_solve.x0 is not None else opt.zeros_like(fwd_args['y']))
            fwd_args.pop('is_backprop', None)
            dy = solve_with_grad(dx, grad_solve, transposed, is_backprop=True, **fwd_args)
            return {'y': dy, 'matrix': compute_dm(dy)}

        solve_with_grad = opt.custom_gradient(forward_solve, backward, auxiliary_args=auxiliary_args)
        return solve_with_grad

    # ---- the coordinate form of a bridged matrix.  compress_rows() keeps the coordinates it started from (_uncompressed_indices, phiml/math/_sparse.py:383-390)
    #      and decompress() hands them back (:895-908); the bridge never had them (they are the 2 * rank * 8 bytes per entry it avoids), so decompress() —
    #      called by math.factor_ilu (_optimize.py:922), explicit_coarse (:966), the reference's matrix gradient (:801) or a user — would raise
    #      NotImplementedError.  They are rebuilt from the CSR arrays on demand, in the reference's layout: (sp_entries, sparse_idx = dual names + primal names), int64.
    from phiml.math._sparse import CompressedSparseMatrix
    original_decompress = CompressedSparseMatrix.decompress

    def decompress(self):
        # (1-D grids: the reference has a branch of its own for matrices without kept coordinates, with the labels the other way round; only matrices the
        #  bridge made are given the compress_rows() layout there, everything else keeps the reference's branch)
        rebuild = not (self._compressed_dims.rank == self._uncompressed_dims.rank == 1) or getattr(self, '_phb200_device_assembled', False)
        if self._uncompressed_indices is None and rebuild and self._uncompressed_offset is None \
                and self._indices.shape.names == ('sp_entries',) and self._pointers.shape.names == ('pointers',):
            self._uncompressed_indices = assembly.coordinates_of_compressed(self)
            self._uncompressed_indices_perm = None
        return original_decompress(self)

    CompressedSparseMatrix.decompress = decompress

    # ---- math.factor_ilu (_optimize.py:883-939) calls incomplete_lu_coo (looked up as a module global at :925): one matrix with one channel whose values are a
    #      CUDA tensor is factored by the device sweeps (bit-identical factors, tests/golden/ilu_edges.npz) and handed back in the reference's layout; index
    #      arrays living on the device (bridged matrices) are brought to the host first, where the reference insists on NumPy (:545)
    original_ilu_coo = opt.incomplete_lu_coo

    def incomplete_lu_coo(indices, values, shape, iterations, safe):
        out = _precond.incomplete_lu_coo_device(indices, values, shape, iterations, safe)
        if out is not None:
            return out
        if isinstance(indices, torch.Tensor):
            indices = indices.detach().cpu().numpy()
        return original_ilu_coo(indices, values, shape, iterations, safe)

    opt.incomplete_lu_coo = incomplete_lu_coo

    TORCH.__class__ = B200TorchBackend
    opt.compute_preconditioner = compute_preconditioner
    opt.native_matrix = native_matrix
    opt.attach_gradient_solve = attach_gradient_solve
    _INSTALLED.update(torch=TORCH, original_class=TorchBackend, original_compute=original_compute, original_native=original_native, opt=opt, original_attach=original_attach,
                      lin_trace=lin_trace, original_t2c=original_t2c, csm=CompressedSparseMatrix, original_decompress=original_decompress, original_ilu_coo=original_ilu_coo)
    return TORCH


def uninstall():
    if not _INSTALLED:
        return
    _INSTALLED['torch'].__class__ = _INSTALLED['original_class']
    _INSTALLED['opt'].compute_preconditioner = _INSTALLED['original_compute']
    _INSTALLED['opt'].native_matrix = _INSTALLED['original_native']
    _INSTALLED['opt'].attach_gradient_solve = _INSTALLED['original_attach']
    _INSTALLED['lin_trace'].tracer_to_coo = _INSTALLED['original_t2c']
    _INSTALLED['csm'].decompress = _INSTALLED['original_decompress']
    _INSTALLED['opt'].incomplete_lu_coo = _INSTALLED['original_ilu_coo']
    NATIVE_CACHE.clear()
    PRECOND_CACHE.clear()
    _INSTALLED.clear()
