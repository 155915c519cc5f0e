"""ctypes binding of libp3d_b200.so (the C ABI declared in include/p3d_b200.h).

There is no CPU fallback: if the shared library is missing, or no CUDA device is
visible when a compute entry point is called, an exception is raised.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("P3D_LIB") or os.path.join(_HERE, "csrc", "libp3d_b200.so")   # P3D_LIB: A/B builds (tools)

OK = 0
ERR_BAD_ARG, ERR_NOT_IMPLEMENTED, ERR_CUDA, ERR_OOM, ERR_NUMERIC = -1, -2, -3, -4, -5
MEM_HOST, MEM_DEVICE = 0, 1
OPS = {"hard": 0, "soft": 1, "garrote": 2, "garotte": 2}
MODELS = {"linear": 0, "exponential": 1, "data-driven": 2, "inverse_proportional": 3}
VERSIONS = {"regular": 0, "fast": 1, "adaptive": 2}
PROFILE_KINDS = ("rows_init", "cols_stats", "cols_iter", "rows_iter", "time_fft", "time_ifft", "sort", "fft2",
                 "cols_iter64", "rows_iter64", "init64", "replay")


class PocsParams(C.Structure):
    _fields_ = [
        ("niter", C.c_int32), ("thresh_op", C.c_int32), ("thresh_model", C.c_int32), ("version", C.c_int32),
        ("q", C.c_double), ("eps", C.c_double), ("alpha", C.c_double), ("p_max", C.c_double), ("p_min", C.c_double),
        ("p_min_adaptive", C.c_int32), ("sqrt_decay", C.c_int32), ("decay_factors", C.c_int32),
        ("absmax_threshold", C.c_int32), ("thresh_percentile", C.c_int32),
    ]


class P3dError(RuntimeError):
    pass


_lib = None

# name -> (restype, argtypes); also the list the ABI test checks against include/p3d_b200.h
SIGNATURES = {
    "p3d_abi_version": (C.c_int, []),
    "p3d_last_error": (C.c_char_p, []),
    "p3d_device_count": (C.c_int, []),
    "p3d_plan_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int]),
    "p3d_plan_destroy": (C.c_int, [C.c_void_p]),
    "p3d_pocs_run": (C.c_int, [C.c_void_p, C.POINTER(PocsParams), C.c_void_p, C.c_int, C.c_void_p, C.c_int64,
                               C.c_void_p, C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "p3d_pocs_schedule": (C.c_int, [C.c_void_p, C.POINTER(PocsParams), C.c_void_p, C.c_int, C.c_int64, C.c_void_p]),
    "p3d_kxky_filter_run": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int64]),
    "p3d_fft2": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int64, C.c_int]),
    "p3d_time_fft": (C.c_int, [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.c_int64,
                               C.c_double, C.c_double, C.c_int, C.c_void_p]),
    "p3d_time_ifft": (C.c_int, [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.c_int64,
                                C.c_double, C.c_double, C.c_int, C.c_int]),
    "p3d_time_envelope": (C.c_int, [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int64, C.c_int64]),
    "p3d_time_last_kernel_ms": (C.c_int, [C.POINTER(C.c_double)]),
    "p3d_time_last_path": (C.c_char_p, []),
    "p3d_host_alloc": (C.c_int, [C.POINTER(C.c_void_p), C.c_int64]),
    "p3d_host_free": (C.c_int, [C.c_void_p]),
    "p3d_device_alloc": (C.c_int, [C.c_int, C.POINTER(C.c_void_p), C.c_int64]),
    "p3d_device_free": (C.c_int, [C.c_int, C.c_void_p]),
    "p3d_memcpy": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_int]),
    "p3d_device_synchronize": (C.c_int, [C.c_int]),
    "p3d_plan_set_profiling": (C.c_int, [C.c_void_p, C.c_int]),
    "p3d_plan_get_profile": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]),
    "p3d_plan_event_record": (C.c_int, [C.c_void_p, C.c_int]),
    "p3d_plan_event_elapsed_ms": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_double)]),
    "p3d_plan_describe": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int64]),
    "p3d_plan_set_option": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int64]),
    "p3d_plan_get_escalation": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
}


def load():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise P3dError(
            f"{LIB_PATH} not found: build it with `python pseudo-3d-interpolation_b200/build.py` "
            "(there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc == OK:
        return
    msg = load().p3d_last_error().decode(errors="replace")
    if rc == ERR_BAD_ARG:
        raise ValueError(msg)
    if rc == ERR_NOT_IMPLEMENTED:
        raise NotImplementedError(msg)
    if rc == ERR_OOM:
        raise MemoryError(msg)
    if rc == ERR_NUMERIC:
        raise IndexError(msg)          # the reference raises IndexError (v[0] on an empty candidate set)
    raise P3dError(msg)


def require_gpu():
    n = load().p3d_device_count()
    if n <= 0:
        raise P3dError("no CUDA device visible: the B200 path has no CPU fallback")
    return n


def ptr(a):
    """Raw pointer of a numpy array / int address / None."""
    if a is None:
        return None
    if isinstance(a, (int, np.integer)):
        return C.c_void_p(int(a))
    return C.c_void_p(a.ctypes.data)


class PinnedArray:
    """numpy view over cudaMallocHost memory (freed with the object)."""

    def __init__(self, shape, dtype):
        lib = load()
        self.nbytes = int(np.prod(shape, dtype=np.int64)) * np.dtype(dtype).itemsize
        p = C.c_void_p()
        check(lib.p3d_host_alloc(C.byref(p), self.nbytes))
        self._p = p
        buf = (C.c_char * max(self.nbytes, 1)).from_address(p.value)
        self.array = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape, dtype=np.int64))).reshape(shape)

    def __del__(self):
        try:
            if getattr(self, "_p", None) is not None and self._p.value:
                self.array = None
                load().p3d_host_free(self._p)
                self._p = None
        except Exception:
            pass


class DeviceBuffer:
    """Raw device allocation on one GPU (for device-resident runs without torch)."""

    def __init__(self, nbytes, device=0):
        self.device, self.nbytes = device, int(nbytes)
        p = C.c_void_p()
        check(load().p3d_device_alloc(device, C.byref(p), self.nbytes))
        self.ptr = p.value

    def upload(self, arr):
        arr = np.ascontiguousarray(arr)
        check(load().p3d_memcpy(self.device, C.c_void_p(self.ptr), ptr(arr), arr.nbytes, 0))

    def download(self, arr):
        check(load().p3d_memcpy(self.device, ptr(arr), C.c_void_p(self.ptr), arr.nbytes, 1))
        return arr

    def free(self):
        if self.ptr:
            load().p3d_device_free(self.device, C.c_void_p(self.ptr))
            self.ptr = 0

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass
