"""ctypes binding of the C ABI in include/lanefit_b200.h.

The shared library is built in-tree (``python -m lanedetection_end2end_b200.csrc.build``
or ``__graft_entry__.build()``) as ``lanedetection_end2end_b200/liblanefit_b200.so``.
There is NO fallback: if the library is missing, every device op raises.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liblanefit_b200.so")

LF_OK = 0
LF_F32, LF_BF16 = 0, 1
ACT_IDS = {"none": 0, "square": 1, "abs": 2, "relu": 3, "sigmoid": 4, "softplus": 5}
SOLVER_INVERSE, SOLVER_CHOLESKY = 0, 1
STATUS_SINGULAR, STATUS_NONFINITE, STATUS_NOT_POSDEF = 1, 2, 4
MAX_ORDER = 4

_c_void_p = ctypes.c_void_p
_c_int = ctypes.c_int
_c_double = ctypes.c_double
_c_float = ctypes.c_float
_c_size_t = ctypes.c_size_t

# name -> (restype, argtypes); every symbol include/lanefit_b200.h declares
PROTOTYPES = {
    "lf_version": (_c_int, []),
    "lf_error_string": (ctypes.c_char_p, [_c_int]),
    "lf_last_cuda_error": (ctypes.c_char_p, []),
    "lf_lsq_workspace_bytes": (_c_size_t, [_c_int] * 5),
    "lf_lsq_fwd": (_c_int, [_c_void_p, _c_int, _c_void_p, _c_void_p, _c_void_p,
                            _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_double, _c_int,
                            _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_size_t, _c_void_p]),
    "lf_lsq_bwd": (_c_int, [_c_void_p, _c_int, _c_void_p, _c_void_p, _c_void_p,
                            _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int,
                            _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p]),
}

_lib = None


class LanefitError(RuntimeError):
    pass


def lib():
    """Load (once) and return the ctypes handle; raises if the library was not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise LanefitError(
                "liblanefit_b200.so not found at %s -- build it with "
                "`python -m lanedetection_end2end_b200.csrc.build` (there is no CPU/eager fallback)" % LIB_PATH)
        h = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(h, name)   # AttributeError if the .so is stale
            fn.restype = res
            fn.argtypes = args
        _lib = h
    return _lib


def check(rc, what):
    if rc != LF_OK:
        h = lib()
        msg = h.lf_error_string(rc).decode()
        if rc == -4:
            msg += ": " + h.lf_last_cuda_error().decode()
        raise LanefitError("%s failed: %s (code %d)" % (what, msg, rc))


def ptr(t):
    """Device pointer of a tensor (or None)."""
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def stream_ptr():
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def dtype_id(t):
    import torch
    if t.dtype == torch.float32:
        return LF_F32
    if t.dtype == torch.bfloat16:
        return LF_BF16
    raise LanefitError("unsupported dtype %s (float32 or bfloat16 expected)" % t.dtype)


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise LanefitError(
                "lanedetection_end2end_b200 runs on CUDA (sm_100a) only; got a %s tensor. "
                "There is no CPU fallback by design." % t.device)
