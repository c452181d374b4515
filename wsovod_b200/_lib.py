"""ctypes binding of libwsovod_b200.so (the C ABI declared in include/wsovod_b200.h).

There is no CPU implementation and no fallback: if the shared library is missing or a call returns
a non-zero code this module raises.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libwsovod_b200.so")

_lib = None

c_i64 = ctypes.c_int64
c_int = ctypes.c_int
c_f = ctypes.c_float
c_d = ctypes.c_double
c_p = ctypes.c_void_p
c_sz = ctypes.c_size_t

# name -> (restype, argtypes); mirrors include/wsovod_b200.h one to one
SIGNATURES = {
    "wsovod_b200_abi_version": (c_int, []),
    "wsovod_b200_strerror": (ctypes.c_char_p, [c_int]),
    "wsovod_b200_launch_count": (ctypes.c_uint64, []),
    "wsovod_b200_tune": (c_int, [c_int, c_int]),
    "wsovod_b200_roi_pool_workspace": (c_sz, [c_i64, c_i64, c_int, c_int]),
    "wsovod_b200_roi_pool_fwd": (c_int, [c_p, c_i64, c_i64, c_i64, c_i64, c_p, c_i64, c_f, c_int, c_int,
                                         c_p, c_f, c_p, c_p, c_p, c_sz, c_p]),
    "wsovod_b200_roi_pool_bwd": (c_int, [c_p, c_p, c_p, c_i64, c_i64, c_i64, c_i64, c_i64, c_int, c_int,
                                         c_p, c_p]),
    "wsovod_b200_roi_loop_pool_workspace": (c_sz, [c_i64, c_i64, c_int, c_int]),
    "wsovod_b200_roi_loop_pool_fwd": (c_int, [c_p, c_i64, c_i64, c_i64, c_i64, c_p, c_i64, c_f, c_int,
                                              c_int, c_p, c_f, c_p, c_p, c_p, c_sz, c_p]),
    "wsovod_b200_roi_loop_pool_bwd": (c_int, [c_p, c_p, c_p, c_i64, c_i64, c_i64, c_i64, c_i64, c_int,
                                              c_int, c_p, c_p]),
    "wsovod_b200_roi_loop_pool_dtype_workspace": (c_sz, [c_i64, c_i64, c_int, c_int]),
    "wsovod_b200_roi_loop_pool_dtype_fwd": (c_int, [c_int, c_p, c_i64, c_i64, c_i64, c_i64, c_p, c_i64, c_f, c_int,
                                                    c_int, c_p, c_p, c_p, c_sz, c_p]),
    "wsovod_b200_roi_loop_pool_dtype_bwd": (c_int, [c_int, c_p, c_p, c_p, c_i64, c_i64, c_i64, c_i64, c_i64, c_int,
                                                    c_int, c_p, c_p]),
    "wsovod_b200_roi_align_workspace": (c_sz, [c_i64, c_i64, c_int, c_int]),
    "wsovod_b200_roi_align_workspace_hw": (c_sz, [c_i64, c_i64, c_int, c_int, c_i64, c_i64]),
    "wsovod_b200_roi_align_fwd": (c_int, [c_p, c_i64, c_i64, c_i64, c_i64, c_p, c_i64, c_f, c_int, c_int,
                                          c_int, c_int, c_p, c_f, c_p, c_p, c_sz, c_p]),
    "wsovod_b200_roi_align_bwd": (c_int, [c_p, c_p, c_i64, c_i64, c_i64, c_i64, c_i64, c_f, c_int, c_int, c_int, c_int, c_p,
                                          c_p]),
    "wsovod_b200_align_workspace": (c_sz, [c_i64, c_i64, c_i64, c_int]),
    "wsovod_b200_align_fwd": (c_int, [c_p, c_p, c_i64, c_i64, c_i64, c_f, c_int, c_int, c_p, c_int,
                                      c_p, c_p, c_p, c_sz, c_p]),
    "wsovod_b200_align_bwd_workspace": (c_sz, [c_i64, c_i64, c_i64]),
    "wsovod_b200_align_bwd": (c_int, [c_p, c_p, c_p, c_i64, c_i64, c_i64, c_f, c_int, c_int, c_p, c_p,
                                      c_p, c_sz, c_p]),
    "wsovod_b200_mil_workspace": (c_sz, [c_i64, c_i64, c_i64]),
    "wsovod_b200_mil_fwd": (c_int, [c_p, c_p, c_p, c_i64, c_i64, c_i64, c_p, c_p, c_p, c_sz, c_p]),
    "wsovod_b200_mil_bwd": (c_int, [c_p, c_p, c_p, c_p, c_p, c_i64, c_i64, c_i64, c_p, c_p, c_p, c_sz,
                                    c_p]),
    "wsovod_b200_align_mil_fused_workspace": (c_sz, [c_i64, c_i64, c_i64, c_i64]),
    "wsovod_b200_align_mil_fused_fwd": (c_int, [c_p, c_p, c_p, c_p, c_i64, c_i64, c_i64, c_i64, c_f, c_int, c_p, c_p,
                                                c_p, c_p, c_p, c_sz, c_p]),
    "wsovod_b200_pgt_top1": (c_int, [c_p, c_i64, c_p, c_p, c_p, c_p, c_p, c_i64, c_i64, c_i64, c_i64,
                                     c_p, c_p, c_p, c_p, c_p, c_p, c_p]),
    "wsovod_b200_refine_assign": (c_int, [c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_i64, c_i64, c_i64,
                                          c_f, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p]),
    "wsovod_b200_refine_loss_workspace": (c_sz, [c_i64]),
    "wsovod_b200_refine_loss_fwd": (c_int, [c_p, c_i64, c_p, c_p, c_p, c_p, c_p, c_i64, c_i64, c_i64, c_f, c_f,
                                            c_f, c_f, c_f, c_p, c_p, c_p, c_sz, c_p]),
    "wsovod_b200_refine_loss_bwd": (c_int, [c_p, c_p, c_p, c_i64, c_p, c_p, c_p, c_p, c_p, c_p, c_i64, c_i64,
                                            c_i64, c_f, c_f, c_f, c_f, c_f, c_p, c_p, c_p]),
    "wsovod_b200_batched_nms_workspace": (c_sz, [c_i64, c_i64]),
    "wsovod_b200_batched_nms": (c_int, [c_p, c_p, c_p, c_i64, c_i64, c_d, c_int, c_p, c_p, c_p, c_sz,
                                        c_p]),
    "wsovod_b200_detections_workspace": (c_sz, [c_i64, c_i64, c_i64, c_i64]),
    "wsovod_b200_detections": (c_int, [c_p, c_p, c_p, c_p, c_i64, c_i64, c_i64, c_i64, c_f, c_d, c_i64, c_int,
                                       c_p, c_p, c_p, c_p, c_p, c_p, c_sz, c_p]),
    "wsovod_b200_csc_workspace": (c_sz, [c_i64, c_i64, c_i64]),
    "wsovod_b200_csc_fwd": (c_int, [c_p, c_p, c_p, c_p, c_i64, c_i64, c_i64, c_i64, c_i64, c_f, c_int, c_f, c_p, c_p, c_sz,
                                    c_p]),
    "wsovod_b200_infer_host_arena": (c_sz, [c_i64, c_i64, c_i64, c_i64, c_i64, c_i64, c_i64, c_int,
                                            c_i64, c_int]),
    "wsovod_b200_infer_host": (c_int, [c_p, c_i64, c_i64, c_i64, c_i64, c_p, c_p, c_i64, c_p, c_p, c_p,
                                       c_p, c_i64, c_i64, c_f, c_int, c_f, c_f, c_d, c_i64, c_int, c_int,
                                       c_int, c_p, c_p, c_p, c_p, c_p, c_p, c_sz, c_p, c_p, c_p]),
}


def lib():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: build it with `python -m wsovod_b200.build` "
                "(wsovod_b200 has no CPU or PyTorch fallback)")
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)          # AttributeError if the .so lacks a declared symbol
            fn.restype = res
            fn.argtypes = args
        got = L.wsovod_b200_abi_version()
        if got != 1:
            raise RuntimeError(f"libwsovod_b200 ABI version {got}, expected 1")
        _lib = L
    return _lib


def check(rc, what):
    if rc != 0:
        msg = lib().wsovod_b200_strerror(rc).decode()
        raise RuntimeError(f"wsovod_b200::{what} failed ({rc}): {msg}")


def launch_count():
    return int(lib().wsovod_b200_launch_count())


F32, F16, F64 = 0, 1, 2
TUNE_POOL_PATH, TUNE_POOL_GROUP, TUNE_ALIGN_PAIR = 0, 1, 2
POOL_AUTO, POOL_SCAN, POOL_BLOCKMAX = 0, 1, 2


def tune(key, value):
    """wsovod_b200_tune: process-wide switch between bit-identical kernels (tests / benches); returns the old value"""
    return int(lib().wsovod_b200_tune(int(key), int(value)))
