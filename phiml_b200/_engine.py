"""
ctypes binding of libphb200.so (C ABI declared in include/phb200.h).

There is no CPU fallback: if the library is missing or a call fails, an exception is raised.
Build the library with ``python -c "import __graft_entry__ as g; g.build()"`` or ``make -C phiml_b200/csrc``.
"""
import ctypes
import os
from ctypes import POINTER, byref, c_char_p, c_double, c_int, c_int32, c_int64, c_size_t, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'lib', 'libphb200.so')

F32, F64 = 0, 1
OK, ERR_ARG, ERR_CUDA, ERR_UNSUPPORTED, ERR_WORKSPACE, ERR_NOT_STENCIL = 0, -1, -2, -3, -4, -5
CG, CG_ADAPTIVE, CG_TORCH, CG_ADAPTIVE_TORCH, BICGSTAB = 0, 1, 2, 3, 4
PRE_NONE, PRE_JACOBI, PRE_ILU = 0, 1, 2
MAT_CSR, MAT_STENCIL = 0, 1

_P = c_void_p
_PP = POINTER(c_void_p)

# name -> (restype, argtypes); must list every function of include/phb200.h (tests/test_abi.py checks it)
SIGNATURES = {
    'phb200_version': (c_int, []),
    'phb200_last_error': (c_char_p, []),
    'phb200_launch_count': (c_int64, []),
    'phb200_matrix_csr': (c_int, [_PP, _P, _P, _P, c_int64, c_int64, c_int64, c_int, c_int]),
    'phb200_matrix_stencil': (c_int, [_PP, c_int, POINTER(c_int64), c_int, POINTER(c_int32), POINTER(c_double), c_int]),
    'phb200_stencil_from_csr': (c_int, [_PP, _P, c_int, POINTER(c_int64), _P]),
    'phb200_shift_count': (c_int, [c_int, POINTER(c_int64), c_int, POINTER(c_int32), _P, c_int, _P, POINTER(c_int64), _P]),
    'phb200_shift_fill': (c_int, [c_int, POINTER(c_int64), c_int, POINTER(c_int32), _P, c_int, _P, _P, _P, _P]),
    'phb200_coded_from_csr': (c_int, [_PP, _P, _P]),
    'phb200_coded_info': (c_int, [_P, POINTER(c_int), POINTER(c_int), POINTER(c_int64), POINTER(c_int)]),
    'phb200_stencil_nnz': (c_int, [_P, POINTER(c_int64)]),
    'phb200_stencil_to_csr': (c_int, [_P, _P, _P, _P, _P]),
    'phb200_matrix_set_slab': (c_int, [_P, c_int64, c_int64]),
    'phb200_matrix_info': (c_int, [_P, POINTER(c_int), POINTER(c_int64), POINTER(c_int64), POINTER(c_int64), POINTER(c_int)]),
    'phb200_matrix_destroy': (c_int, [_P]),
    'phb200_spmv': (c_int, [_P, _P, _P, c_int64, c_int64, c_int64, _P]),
    'phb200_spmv_coo': (c_int, [_P, _P, _P, c_int64, c_int64, c_int64, _P, _P, c_int64, c_int64, c_int64, c_int, _P]),
    'phb200_csr_dense_batched': (c_int, [_P, _P, _P, _P, _P, c_int64, c_int64, c_int64, c_int64, c_int64, c_int64, c_int, _P]),
    'phb200_csr_transpose': (c_int, [_P, _P, _P, c_int64, c_int64, c_int64, c_int, _P, _P, _P, _P]),
    'phb200_matrix_gradient_coo': (c_int, [_P, _P, c_int64, c_int64, c_int64, _P, _P, c_int64, c_int64, c_int64, _P, c_int, _P]),
    'phb200_matrix_gradient_csr': (c_int, [_P, _P, c_int64, c_int64, c_int64, _P, _P, c_int64, c_int64, c_int64, _P, c_int, _P]),
    'phb200_jacobi_setup': (c_int, [_P, _P, _P]),
    'phb200_precond_jacobi': (c_int, [_PP, _P, c_int64, c_int, c_int]),
    'phb200_precond_ilu0': (c_int, [_PP, _P, _P, _P, _P]),
    'phb200_precond_ilu_sweeps': (c_int, [_PP, _P, _P, c_int, c_int, _P, _P]),
    'phb200_precond_cluster': (c_int, [_PP, _P, _P, _P, c_int64, c_int, _P, _P, c_int]),
    'phb200_solver_set_dist_callbacks': (c_int, [_P, _P, _P, _P, _P]),
    'phb200_solver_set_peer_halos': (c_int, [_P, _P, c_int64, _P, c_int64]),
    'phb200_solver_set_peer_native': (c_int, [_P, _P, POINTER(c_void_p), c_int, c_int, _P, _P, _P, _P]),
    'phb200_solver_set_peer_reduce': (c_int, [_P, POINTER(c_void_p), c_int, c_int, _P]),
    'phb200_peer_mailbox_bytes': (c_size_t, [c_int, c_int64]),
    'phb200_solver_dist_advance': (c_int, [_P, c_int, _P]),
    'phb200_ipc_export': (c_int, [_P, _P, POINTER(c_int64)]),
    'phb200_ipc_open': (c_int, [_P, _PP]),
    'phb200_ipc_close': (c_int, [_P]),
    'phb200_solver_option': (c_int, [_P, c_int, c_int, POINTER(c_int)]),
    'phb200_precond_info': (c_int, [_P, POINTER(c_int), POINTER(c_int)]),
    'phb200_precond_levels': (c_int, [_P, POINTER(c_int64), POINTER(c_int64)]),
    'phb200_sptrsv_plan': (c_int, [_PP, _P, c_int, _P]),
    'phb200_sptrsv': (c_int, [_P, _P, c_int, c_int, _P, _P, c_int64, c_int64, _P]),
    'phb200_sptrsv_plan_destroy': (c_int, [_P]),
    'phb200_precond_apply': (c_int, [_P, _P, _P, c_int64, c_int64, _P]),
    'phb200_precond_destroy': (c_int, [_P]),
    'phb200_solver_create': (c_int, [_PP, _P, _P, c_int, c_int, c_int64]),
    'phb200_solver_workspace_bytes': (c_size_t, [_P]),
    'phb200_solver_start': (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, c_size_t, _P]),
    'phb200_solver_tol_min': (c_int, [_P, _P, c_int, _P]),
    'phb200_solver_set_dist': (c_int, [_P, _P]),
    'phb200_solver_dist_step': (c_int, [_P, c_int, _P]),
    'phb200_solver_advance': (c_int, [_P, c_int, _P]),
    'phb200_solver_run': (c_int, [_P, c_int64, _P]),
    'phb200_solver_poll': (c_int, [_P, POINTER(c_int), _P]),
    'phb200_solver_result': (c_int, [_P, _P, _P, _P, _P, _P]),
    'phb200_solver_stats': (c_int, [_P, POINTER(c_int64), POINTER(c_int64)]),
    'phb200_solver_profile': (c_int, [_P, c_int]),
    'phb200_solver_profile_read': (c_int, [_P, POINTER(c_double), POINTER(c_int64), c_int, POINTER(c_int)]),
    'phb200_solver_destroy': (c_int, [_P]),
}

_lib = None


class EngineError(RuntimeError):
    pass


class NotAStencil(Exception):
    pass


def load(path: str = None) -> ctypes.CDLL:
    """Loads libphb200.so and declares all prototypes.  Raises (never falls back) if the library is absent."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    path = path or LIB_PATH
    if not os.path.exists(path):
        raise EngineError(f"{path} not found. The CUDA extension is required (no CPU fallback); build it with "
                          f"`make -C phiml_b200/csrc` or `python -c 'import __graft_entry__ as g; g.build()'`.")
    lib = ctypes.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)      # AttributeError if the library lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    if lib.phb200_version() < 100:
        raise EngineError("libphb200.so is older than this package; rebuild")
    _lib = lib
    return lib


def lib() -> ctypes.CDLL:
    return _lib if _lib is not None else load()


def check(rc: int):
    if rc == OK:
        return
    msg = lib().phb200_last_error().decode('utf-8', 'replace')
    if rc == ERR_ARG:
        raise ValueError(msg)
    if rc == ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    if rc == ERR_NOT_STENCIL:
        raise NotAStencil(msg)
    if rc == ERR_WORKSPACE:
        raise MemoryError(msg)
    raise EngineError(f"phb200 error {rc}: {msg}")


def dtype_code(torch_dtype) -> int:
    import torch
    if torch_dtype == torch.float32:
        return F32
    if torch_dtype == torch.float64:
        return F64
    raise TypeError(f"phb200 computes in float32 or float64, got {torch_dtype}")


def ptr(t) -> c_void_p:
    return c_void_p(t.data_ptr()) if t is not None else c_void_p(0)


def stream_ptr(device=None) -> c_void_p:
    """Current stream of ``device``.  The library allocates its scratch and launches on the CURRENT device, so the operands' device must be
    the current one — the public entry points enter it (``on_operand_device``); anything else fails loudly here instead of running across GPUs."""
    import torch
    if device is not None:
        idx = torch.device(device).index
        if idx is not None and idx != torch.cuda.current_device():
            raise EngineError(f"operands live on cuda:{idx} but the current device is cuda:{torch.cuda.current_device()}; call inside torch.cuda.device({idx})")
    return c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _device_of(obj):
    import torch
    if isinstance(obj, torch.Tensor):
        return obj.device if obj.is_cuda else None
    dev = getattr(obj, 'device', None)           # Matrix, preconditioners (through .matrix)
    if isinstance(dev, torch.device) and dev.type == 'cuda':
        return dev
    m = getattr(obj, 'matrix', None)
    return _device_of(m) if m is not None and m is not obj else None


def on_operand_device(fn):
    """Runs ``fn`` with the device of its first CUDA operand (tensor, Matrix or preconditioner) as the current device (ADVICE round 1:
    scratch, set-up temporaries and launches must not land on another GPU when the caller's current device differs from the operands')."""
    import functools

    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        import torch
        dev = None
        for a in list(args) + list(kwargs.values()):
            dev = _device_of(a)
            if dev is not None:
                break
        if dev is None or dev.index is None or dev.index == torch.cuda.current_device():
            return fn(*args, **kwargs)
        with torch.cuda.device(dev):
            return fn(*args, **kwargs)
    return wrapper


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise EngineError("phb200 operates on CUDA tensors only (there is no CPU fallback); got a tensor on " + str(t.device))


def launch_count() -> int:
    return int(lib().phb200_launch_count())
