"""ctypes binding of ``include/sspslam_b200.h`` (the only bridge to the CUDA library).

The shared object is built in-tree by ``__graft_entry__.build()`` (plain ``nvcc`` for
sm_100a).  If it is missing or cannot be loaded this module raises — there is no CPU
fallback for the stepped path or the SSP kernels.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libssb.so")
KERNEL_KINDS = ("ens_small", "ens_wide", "decode", "pes", "cleanup_scan", "cleanup_pick", "gate", "lin", "advance",
                "begin", "ens_voja", "extra", "pes_hist", "pes_fold")

_lib = None


class SsbError(RuntimeError):
    pass


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise SsbError(f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                       "(nvcc, sm_100a). The B200 backend has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    P, I, LL, SZ, D, CP = C.c_void_p, C.c_int, C.c_longlong, C.c_size_t, C.c_double, C.c_char_p
    protos = {
        "ssb_create": (I, [I, I, C.POINTER(P)]),
        "ssb_set_array": (I, [P, CP, P, SZ]),
        "ssb_set_scalar": (I, [P, CP, D]),
        "ssb_finalize": (I, [P]),
        "ssb_upload": (I, [P, CP, SZ, SZ, P]),
        "ssb_download": (I, [P, CP, SZ, SZ, P]),
        "ssb_set_tables": (I, [P, P, LL, I]),
        "ssb_rebase_tables": (I, [P, LL]),
        "ssb_run_steps": (I, [P, I]),
        "ssb_run_steps_io": (I, [P, P, I, P]),
        "ssb_io_wait": (I, [P]),
        "ssb_synth_setup": (I, [P, P, P, P, P, P, P, P]),
        "ssb_synth_steps": (I, [P, P, LL, I]),
        "ssb_read_probes": (I, [P, P, LL, I]),
        "ssb_n_steps": (LL, [P]),
        "ssb_n_trials_padded": (I, [P]),
        "ssb_sync": (I, [P]),
        "ssb_reset": (I, [P]),
        "ssb_destroy": (None, [P]),
        "ssb_set_profiling": (I, [P, I]),
        "ssb_last_run_ms": (I, [P, C.POINTER(C.c_float)]),
        "ssb_kernel_times": (I, [P, P, P, I]),
        "ssb_total_launches": (LL, [P]),
        "ssb_timeline": (I, [P, P, P, P, I, C.POINTER(I)]),
        "ssb_mark": (I, [P, I]),
        "ssb_mark_elapsed_ms": (I, [P, I, I, C.POINTER(C.c_float)]),
        "ssb_ssp_encode": (I, [I, P, P, P, LL, I, I]),
        "ssb_ssp_decode_argmax": (I, [I, P, P, P, LL, LL, I]),
        "ssb_host_alloc": (P, [SZ]),
        "ssb_host_free": (None, [P]),
        "ssb_last_error": (CP, []),
        "ssb_version": (CP, []),
    }
    for name, (res, args) in protos.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


BUILDER_LIB_PATH = os.path.join(_HERE, "libssb_builder.so")
BUILDER_EXPORTS = ("ssb_solve_decoders", "ssb_builder_last_error")
_builder_lib = None


def load_builder():
    """``libssb_builder.so`` (include/sspslam_b200_builder.h): batched decoder solves on the device (cuBLAS / cuSOLVER)."""
    global _builder_lib
    if _builder_lib is not None:
        return _builder_lib
    if not os.path.isfile(BUILDER_LIB_PATH):
        raise SsbError(f"{BUILDER_LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'`")
    lib = C.CDLL(BUILDER_LIB_PATH)
    lib.ssb_solve_decoders.restype = C.c_int
    lib.ssb_solve_decoders.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_double,
                                       C.c_void_p]
    lib.ssb_builder_last_error.restype = C.c_char_p
    lib.ssb_builder_last_error.argtypes = []
    _builder_lib = lib
    return lib


def solve_decoders(A, Y, reg, device=0):
    """Batched LstsqL2 on the device: ``A`` [n_sys, m, n], ``Y`` [n_sys, m, k] -> ``X`` [n_sys, n, k] (float64)."""
    lib = load_builder()
    A = np.ascontiguousarray(A, dtype=np.float64)
    Y = np.ascontiguousarray(Y, dtype=np.float64)
    n_sys, m, n = A.shape
    k = Y.shape[2]
    X = np.empty((n_sys, n, k), dtype=np.float64)
    rc = lib.ssb_solve_decoders(int(device), n_sys, m, n, k, _ptr(A), _ptr(Y), float(reg), _ptr(X))
    if rc != 0:
        raise SsbError(f"ssb_solve_decoders failed ({rc}): {lib.ssb_builder_last_error().decode(errors='replace')}")
    return X


EXPORTS = ("ssb_create", "ssb_set_array", "ssb_set_scalar", "ssb_finalize", "ssb_upload", "ssb_download",
           "ssb_set_tables", "ssb_rebase_tables", "ssb_run_steps", "ssb_run_steps_io", "ssb_io_wait", "ssb_synth_setup", "ssb_synth_steps", "ssb_read_probes", "ssb_n_steps",
           "ssb_n_trials_padded", "ssb_sync", "ssb_reset", "ssb_destroy", "ssb_set_profiling", "ssb_last_run_ms",
           "ssb_kernel_times", "ssb_total_launches", "ssb_timeline", "ssb_mark", "ssb_mark_elapsed_ms", "ssb_ssp_encode", "ssb_ssp_decode_argmax", "ssb_host_alloc",
           "ssb_host_free", "ssb_last_error", "ssb_version")


def check(rc, what=""):
    if rc != 0:
        msg = load().ssb_last_error().decode(errors="replace")
        raise SsbError(f"{what} failed ({rc}): {msg}")


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class PinnedBuffer:
    """float32 host staging buffer in page-locked memory (``cudaHostAlloc``)."""

    def __init__(self, n_floats):
        lib = load()
        self.nbytes = int(n_floats) * 4
        self.ptr = lib.ssb_host_alloc(max(self.nbytes, 4))
        if not self.ptr:
            raise SsbError("pinned allocation failed: " + lib.ssb_last_error().decode())
        buf = (C.c_float * int(n_floats)).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=np.float32)

    def free(self):
        if self.ptr:
            self.array = None
            load().ssb_host_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def ssp_encode(a_scaled, x, device=0):
    lib = load()
    a = np.ascontiguousarray(a_scaled, dtype=np.float64)
    x = np.ascontiguousarray(x, dtype=np.float64)
    d, n = a.shape
    if x.ndim != 2 or x.shape[1] != n:
        raise ValueError(f"expected points of shape (N, {n})")
    out = np.empty((x.shape[0], d), dtype=np.float64)
    check(lib.ssb_ssp_encode(device, _ptr(a), _ptr(x), _ptr(out), x.shape[0], n, d), "ssb_ssp_encode")
    return out


def ssp_decode_argmax(sample_ssps, queries, device=0):
    lib = load()
    s = np.ascontiguousarray(sample_ssps, dtype=np.float64)
    q = np.ascontiguousarray(queries, dtype=np.float64)
    if q.ndim != 2 or q.shape[1] != s.shape[1]:
        raise ValueError("query / sample dimensionality mismatch")
    idx = np.empty(q.shape[0], dtype=np.int32)
    check(lib.ssb_ssp_decode_argmax(device, _ptr(s), _ptr(q), _ptr(idx), q.shape[0], s.shape[0], s.shape[1]),
          "ssb_ssp_decode_argmax")
    return idx.astype(np.int64)
