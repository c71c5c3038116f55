"""ctypes binding of libkdsl.so (the C ABI in include/kdsl.h).

There is no CPU path: if the shared object is missing this raises, and every compute entry
point fails with KDSL_ERR_CUDA when no CUDA device is visible.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
SO_PATH = os.environ.get("KDSL_LIB") or os.path.join(CSRC, "libkdsl.so")   # KDSL_LIB: developer builds (make libkdsl_ticks.so)

KDSL_OK = 0
KDSL_ERR_INVALID_ARGUMENT = -1
KDSL_ERR_CUDA = -2
KDSL_ERR_SINGULAR = -3
KDSL_ERR_STATE = -4
FLAG_SINGULAR, FLAG_NONFINITE, FLAG_BAD_SITE = 1, 2, 4
(ACC_WALKER_SWEEPS, ACC_SUM_ACC, ACC_SUM_OL, ACC_SUM_OL2, ACC_N_OL, ACC_N_REACH, ACC_N_REFRESH,
 ACC_N_SINGULAR) = range(8)
N_ACC = 8
TIMER_NAMES = ("propose", "update", "refresh_gather", "refresh_inverse", "refresh_gemm", "measure")
N_TIMERS = len(TIMER_NAMES)

# every symbol include/kdsl.h declares (tests check the library exports exactly these)
SYMBOLS = (
    "kdsl_version", "kdsl_last_error", "kdsl_device_count", "kdsl_create", "kdsl_create_c128", "kdsl_is_complex", "kdsl_destroy",
    "kdsl_set_config", "kdsl_get_config", "kdsl_set_rng", "kdsl_get_rng", "kdsl_set_sweeps",
    "kdsl_get_sweeps", "kdsl_refresh", "kdsl_sweep", "kdsl_replay", "kdsl_measure", "kdsl_last_OL",
    "kdsl_accumulators", "kdsl_reset_accumulators", "kdsl_get_W", "kdsl_set_W", "kdsl_update_W",
    "kdsl_get_Z", "kdsl_get_flags", "kdsl_set_profiling", "kdsl_timers", "kdsl_reset_timers",
    "kdsl_set_option", "kdsl_synchronize", "kdsl_info", "kdsl_event_record", "kdsl_event_elapsed", "kdsl_bench_fp64_dmma",
    "kdsl_bench_fp64_dmma_sustained", "kdsl_set_observables", "kdsl_get_observables",
    "kdsl_comm_version", "kdsl_comm_unique_id", "kdsl_comm_init_rank", "kdsl_comm_init_all", "kdsl_comm_info",
    "kdsl_comm_destroy", "kdsl_accumulators_allreduce", "kdsl_group_accumulators_allreduce",
)
COMM_ID_BYTES = 128


class KdslError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libkdsl error {code}: {msg}")
        self.code = code


class SingularException(KdslError):
    """LinearAlgebra.SingularException analogue"""


def build(force: bool = False) -> str:
    """nvcc-compile libkdsl.so in-tree for sm_100a (cross-compiles without a GPU)."""
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh"))]
    srcs.append(os.path.join(_HERE, "..", "include", "kdsl.h"))
    stale = (not os.path.exists(SO_PATH)) or any(os.path.getmtime(s) > os.path.getmtime(SO_PATH) for s in srcs)
    if force or stale:
        subprocess.check_call(["make", "-C", CSRC, "-B", "libkdsl.so"])
    return SO_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise ImportError(
                f"{SO_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). There is no CPU fallback."
            )
        L = C.CDLL(SO_PATH)
        L.kdsl_last_error.restype = C.c_char_p
        vp, i32, i64 = C.c_void_p, C.c_int, C.c_int64
        L.kdsl_create.argtypes = [C.POINTER(vp), i32, i32, i32, i32, i32, vp, vp, vp, i32]
        L.kdsl_create_c128.argtypes = [C.POINTER(vp), i32, i32, i32, i32, i32, vp, vp, vp, i32]
        L.kdsl_is_complex.argtypes = [vp, vp]
        L.kdsl_destroy.argtypes = [vp]
        L.kdsl_set_config.argtypes = [vp, vp, vp]
        L.kdsl_get_config.argtypes = [vp, vp, vp]
        L.kdsl_set_rng.argtypes = [vp, vp]
        L.kdsl_get_rng.argtypes = [vp, vp]
        L.kdsl_set_sweeps.argtypes = [vp, i64]
        L.kdsl_get_sweeps.argtypes = [vp, C.POINTER(i64)]
        L.kdsl_refresh.argtypes = [vp, C.POINTER(i32)]
        L.kdsl_sweep.argtypes = [vp, i64, i64]
        L.kdsl_replay.argtypes = [vp, i64, i64, vp, vp, vp]
        L.kdsl_measure.argtypes = [vp, vp]
        L.kdsl_last_OL.argtypes = [vp, vp, vp]
        L.kdsl_accumulators.argtypes = [vp, vp, vp, vp]
        L.kdsl_reset_accumulators.argtypes = [vp]
        L.kdsl_get_W.argtypes = [vp, i32, i32, vp]
        L.kdsl_set_W.argtypes = [vp, i32, i32, vp]
        L.kdsl_update_W.argtypes = [vp, i32, vp, vp, vp, vp, vp]
        L.kdsl_get_Z.argtypes = [vp, vp, vp]
        L.kdsl_get_flags.argtypes = [vp, vp]
        L.kdsl_set_profiling.argtypes = [vp, i32]
        L.kdsl_timers.argtypes = [vp, vp, vp, vp]
        L.kdsl_reset_timers.argtypes = [vp]
        L.kdsl_set_option.argtypes = [vp, C.c_char_p, i64]
        L.kdsl_synchronize.argtypes = [vp]
        L.kdsl_info.argtypes = [vp, vp]
        L.kdsl_device_count.argtypes = [C.POINTER(i32)]
        L.kdsl_event_record.argtypes = [vp, i32]
        L.kdsl_bench_fp64_dmma.argtypes = [vp, C.POINTER(C.c_double)]
        L.kdsl_bench_fp64_dmma_sustained.argtypes = [vp, C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.kdsl_event_elapsed.argtypes = [vp, i32, i32, C.POINTER(C.c_double)]
        L.kdsl_set_observables.argtypes = [vp, i32, vp, vp]
        L.kdsl_get_observables.argtypes = [vp, vp, i32]
        L.kdsl_comm_version.argtypes = [C.POINTER(i32)]
        L.kdsl_comm_unique_id.argtypes = [vp]
        L.kdsl_comm_init_rank.argtypes = [vp, i32, i32, vp]
        L.kdsl_comm_init_all.argtypes = [i32, C.POINTER(vp)]
        L.kdsl_comm_info.argtypes = [vp, C.POINTER(i32), C.POINTER(i32)]
        L.kdsl_comm_destroy.argtypes = [vp]
        L.kdsl_accumulators_allreduce.argtypes = [vp, vp]
        L.kdsl_group_accumulators_allreduce.argtypes = [i32, C.POINTER(vp), vp]
        _lib = L
    return _lib


def check(rc: int) -> int:
    if rc < 0:
        msg = lib().kdsl_last_error().decode("utf-8", "replace")
        if rc == KDSL_ERR_SINGULAR:
            raise SingularException(rc, msg)
        raise KdslError(rc, msg)
    return rc


def device_count() -> int:
    n = C.c_int(0)
    lib().kdsl_device_count(C.byref(n))
    return n.value
