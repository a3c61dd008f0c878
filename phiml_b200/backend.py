"""
``B200Backend`` — stand-alone mirror of the part of the reference's ``Backend`` operator API that lies on the sparse
linear-solve path (phiml/backend/_backend.py:88 ff.; torch implementation phiml/backend/torch/_torch_backend.py).
It needs no PhiML install: the GPU box runs the parity tests through this class.  ``phiml_b200.phiml_plugin`` grafts the
same functions onto PhiML's own ``TORCH`` backend object when PhiML is importable.
"""
from typing import Tuple

import numpy as np
import torch

from . import solve as _solve
from . import precond as _precond
from .sparse import as_matrix


class B200Backend:
    name = 'torch'

    def __init__(self, device='cuda:0'):
        self.device = torch.device(device)

    def __repr__(self):
        return "phb200"

    # --- tensors ---
    def as_tensor(self, x, convert_external=True):
        if isinstance(x, torch.Tensor):
            return x.to(self.device)
        return torch.as_tensor(np.asarray(x), device=self.device)

    def numpy(self, t):
        return t.detach().cpu().numpy() if isinstance(t, torch.Tensor) else np.asarray(t)

    def is_sparse(self, x) -> bool:
        return isinstance(x, torch.Tensor) and x.layout != torch.strided

    def supports(self, feature) -> bool:
        name = feature if isinstance(feature, str) else feature.__name__
        return hasattr(self, name)

    # --- sparse natives (TorchBackend.sparse_coo_tensor / csr_matrix / csc_matrix, _torch_backend.py:857-896) ---
    def sparse_coo_tensor(self, indices, values, shape):
        indices = self.as_tensor(indices).to(torch.int64).t()
        values = self.as_tensor(values)
        return torch.sparse_coo_tensor(indices, values, size=tuple(shape))

    def csr_matrix(self, column_indices, row_pointers, values, shape: Tuple[int, int]):
        return torch.sparse_csr_tensor(self.as_tensor(row_pointers), self.as_tensor(column_indices), self.as_tensor(values), tuple(shape), device=self.device)

    def csc_matrix(self, column_pointers, row_indices, values, shape: Tuple[int, int]):
        return torch.sparse_csc_tensor(self.as_tensor(column_pointers), self.as_tensor(row_indices), self.as_tensor(values), tuple(shape), device=self.device)

    # --- products ---
    def mul_matrix_batched_vector(self, A, b):
        return _solve.mul_matrix_batched_vector(A, self.as_tensor(b))

    def linear(self, lin, vector):
        if callable(lin) and not isinstance(lin, torch.Tensor):
            return lin(vector)
        return self.mul_matrix_batched_vector(lin, vector)

    def mul_csr_dense(self, column_indices, row_pointers, values, shape, dense):
        return _solve.mul_csr_dense(self.as_tensor(column_indices), self.as_tensor(row_pointers), self.as_tensor(values), shape, self.as_tensor(dense))

    def mul_coo_dense(self, indices, values, shape, dense):
        return _solve.mul_coo_dense(self.as_tensor(indices), self.as_tensor(values), shape, self.as_tensor(dense))

    # --- solves ---
    def linear_solve(self, method, lin, y, x0, rtol, atol, max_iter, pre, matrix_offset):
        return _solve.linear_solve(method, lin, self.as_tensor(y), self.as_tensor(x0), rtol, atol, max_iter, pre, matrix_offset)

    def conjugate_gradient(self, lin, y, x0, rtol, atol, max_iter, pre, matrix_offset):
        return _solve.conjugate_gradient(lin, self.as_tensor(y), self.as_tensor(x0), rtol, atol, max_iter, pre, matrix_offset)

    def conjugate_gradient_adaptive(self, lin, y, x0, rtol, atol, max_iter, pre, matrix_offset):
        return _solve.conjugate_gradient_adaptive(lin, self.as_tensor(y), self.as_tensor(x0), rtol, atol, max_iter, pre, matrix_offset)

    def bi_conjugate_gradient(self, lin, y, x0, rtol, atol, max_iter, pre, matrix_offset, poly_order=2):
        return _solve.bi_conjugate_gradient(lin, self.as_tensor(y), self.as_tensor(x0), rtol, atol, max_iter, pre, matrix_offset, poly_order)

    def solve_triangular(self, matrix, rhs, lower: bool, unit_diagonal: bool):
        return _solve.solve_triangular(matrix, self.as_tensor(rhs), lower, unit_diagonal)

    def solve_triangular_sparse(self, matrix, rhs, lower: bool, unit_diagonal: bool):
        return _solve.solve_triangular_sparse(matrix, self.as_tensor(rhs), lower, unit_diagonal)

    # --- preconditioners ---
    def compute_preconditioner(self, method, matrix):
        return _precond.compute_preconditioner(method, matrix)
