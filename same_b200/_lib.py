"""ctypes binding of libsame_b200.so (include/same_b200.h).

There is no CPU fallback: if the shared library is missing or no CUDA device is
visible, every entry point raises.  Build with `python -m same_b200.build`.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsame_b200.so")

OK, E_CUDA, E_ARG, E_STATE, E_NOPAIRS, E_LIMIT = 0, -1, -2, -3, -4, -5
MAX_KNN = 256

# array ids (SAME_ARR_*)
WIN_A, WIN_R, KEEP_A, KEEP_R, PAIRS, COST = 1, 2, 3, 4, 5, 6
REF_GROUP_NODE, REF_GROUP_PTR, REF_GROUP_IDX, REF_GROUP_LIMIT, ROW_PTR = 7, 8, 9, 10, 11
TRI_IN, TRI_IN_SRC, TRI_CLASS, TRI_BAND, TRI, TRI_SRC = 12, 13, 14, 15, 16, 17
TRI_WEIGHT, TRI_SIGN, TRI_BOUNDS, TRI_ARGV, UNCONSTRAINED = 18, 19, 20, 21, 22
MATCH_J, MATCH_P, TRI_MASK, AREA_BEFORE, AREA_AFTER, FLIPPED = 23, 24, 25, 26, 27, 28
START_X, START_UNMATCHED = 29, 30
PAIR_J = 31
PAIR_J16 = 35
STAT_KNN_EVALUATIONS = 1
NODE_TRI_PTR, NODE_TRI_LEN, NODE_TRI_IDX = 32, 33, 34

#: numpy dtype and trailing shape of every retrievable array
ARRAY_SPEC = {
    WIN_A: (np.int32, ()), WIN_R: (np.int32, ()), KEEP_A: (np.int32, ()), KEEP_R: (np.int32, ()),
    PAIRS: (np.int32, (2,)), COST: (np.float64, ()),
    REF_GROUP_NODE: (np.int32, ()), REF_GROUP_PTR: (np.int32, ()), REF_GROUP_IDX: (np.int32, ()),
    REF_GROUP_LIMIT: (np.int32, ()), ROW_PTR: (np.int32, ()),
    TRI_IN: (np.int32, (3,)), TRI_IN_SRC: (np.int32, ()), TRI_CLASS: (np.uint8, ()), TRI_BAND: (np.int32, ()),
    TRI: (np.int32, (3,)), TRI_SRC: (np.int32, ()), TRI_WEIGHT: (np.float64, ()), TRI_SIGN: (np.int8, ()),
    TRI_BOUNDS: (np.float64, (4,)), TRI_ARGV: (np.int32, (4,)), UNCONSTRAINED: (np.int32, ()),
    MATCH_J: (np.int32, ()), MATCH_P: (np.int32, ()), TRI_MASK: (np.int32, ()),
    AREA_BEFORE: (np.float64, ()), AREA_AFTER: (np.float64, ()), FLIPPED: (np.uint8, ()),
    START_X: (np.uint8, ()), START_UNMATCHED: (np.uint8, ()), PAIR_J: (np.int32, ()), PAIR_J16: (np.uint16, ()), NODE_TRI_PTR: (np.int32, ()), NODE_TRI_LEN: (np.int32, ()), NODE_TRI_IDX: (np.int32, ()),
}

TRI_DROP_RADIUS, TRI_DROP_ANGLE, TRI_SAME_TYPE, TRI_KEEP = 0, 1, 2, 3

#: every symbol include/same_b200.h declares (tests check the library exports all of them)
SYMBOLS = [
    "same_last_error", "same_abi_version", "same_device_count", "same_section_create", "same_section_destroy",
    "same_section_bbox", "same_section_count_rects", "same_section_set_triangles", "same_batch_create",
    "same_batch_destroy", "same_batch_num_windows", "same_batch_candidates", "same_batch_triangles_remap",
    "same_batch_triangles_set", "same_batch_tri_classify", "same_batch_tri_override", "same_batch_tri_finalize",
    "same_batch_groups", "same_batch_separation", "same_batch_postsolve", "same_batch_offsets", "same_batch_length",
    "same_batch_get", "same_elem_size", "same_batch_sync", "same_batch_stream", "same_launch_count",
    "same_profile_enable", "same_profile_report", "same_batch_get_many", "same_batch_get_many_async", "same_pinned_alloc", "same_pinned_free",
    "same_postsolve_arrays", "same_batch_mip_start", "same_greedy_select", "same_collapse_select", "same_segment_mean", "same_measure_fp64_peak", "same_section_wait_uploads",
    "same_stream_create", "same_stream_destroy", "same_mempool_stats", "same_mempool_reserve", "same_set_host_wait", "same_batch_uncertain", "same_batch_stat", "same_debug_guard",
]


class SameError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libsame_b200 error {code}: {msg}")
        self.code = code


_lib = None


def load():
    """Load the CUDA library or raise — never falls back to a CPU path."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: the CUDA extension is not built (run `python -m same_b200.build`). "
            "same_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    vp, i64, i32, dbl = C.c_void_p, C.c_int64, C.c_int, C.c_double
    lib.same_last_error.restype = C.c_char_p
    lib.same_abi_version.restype = i32
    lib.same_device_count.argtypes = [C.POINTER(C.c_int)]
    lib.same_section_create.argtypes = [i32, vp, i64, i64, i32, vp, vp, vp, vp, vp, vp, vp, vp, C.POINTER(vp)]
    lib.same_section_destroy.argtypes = [vp]
    lib.same_section_wait_uploads.argtypes = [vp]
    lib.same_section_bbox.argtypes = [vp, vp]
    lib.same_section_count_rects.argtypes = [vp, i64, vp, vp, vp]
    lib.same_section_set_triangles.argtypes = [vp, vp, vp, i64]
    lib.same_batch_create.argtypes = [vp, i64, vp, C.POINTER(vp)]
    lib.same_batch_destroy.argtypes = [vp]
    lib.same_batch_num_windows.argtypes = [vp]
    lib.same_batch_num_windows.restype = i64
    lib.same_batch_candidates.argtypes = [vp, dbl, i32, i32, dbl]
    lib.same_batch_triangles_remap.argtypes = [vp]
    lib.same_batch_triangles_set.argtypes = [vp, vp, vp]
    lib.same_batch_tri_classify.argtypes = [vp, dbl, i32, dbl, i32, C.POINTER(i64)]
    lib.same_batch_tri_override.argtypes = [vp, i64, vp, vp]
    lib.same_batch_tri_finalize.argtypes = [vp, i32, i32, i32]
    lib.same_batch_groups.argtypes = [vp, i32, i32]
    lib.same_batch_separation.argtypes = [vp, i64, i64, vp, i64, vp, vp, vp]
    lib.same_batch_postsolve.argtypes = [vp, i64, i64, vp]
    lib.same_batch_stat.argtypes = [vp, i32, C.POINTER(i64)]
    lib.same_batch_uncertain.argtypes = [vp, i32, i64, C.POINTER(i64), vp]
    lib.same_batch_offsets.argtypes = [vp, i32, vp]
    lib.same_batch_length.argtypes = [vp, i32, C.POINTER(i64)]
    lib.same_batch_get.argtypes = [vp, i32, i64, i64, vp]
    lib.same_elem_size.argtypes = [i32]
    lib.same_elem_size.restype = i64
    lib.same_batch_sync.argtypes = [vp]
    lib.same_batch_stream.argtypes = [vp]
    lib.same_batch_stream.restype = vp
    lib.same_launch_count.restype = i64
    lib.same_batch_get_many.argtypes = [vp, i64, vp, vp, vp, vp]
    lib.same_batch_get_many_async.argtypes = [vp, i64, vp, vp, vp, vp]
    lib.same_pinned_alloc.argtypes = [i64, C.POINTER(vp)]
    lib.same_pinned_free.argtypes = [vp]
    lib.same_postsolve_arrays.argtypes = [i32, i64, vp, i64, vp, i64, vp, vp, vp, vp, vp, vp]
    lib.same_batch_mip_start.argtypes = [vp, dbl, C.POINTER(C.c_int32)]
    lib.same_greedy_select.argtypes = [i32, i64, i32, vp, vp, vp, i64, vp, vp, C.POINTER(C.c_int32)]
    lib.same_collapse_select.argtypes = [i32, i64, vp, vp, vp, i64, vp, dbl, vp, vp, C.POINTER(C.c_int32)]
    lib.same_segment_mean.argtypes = [i32, i64, i64, vp, i64, vp, vp, vp]
    lib.same_measure_fp64_peak.argtypes = [i32, C.POINTER(C.c_double)]
    lib.same_profile_enable.argtypes = [i32]
    lib.same_stream_create.argtypes = [i32, C.POINTER(vp)]
    lib.same_stream_destroy.argtypes = [i32, vp]
    lib.same_mempool_reserve.argtypes = [i32, i64]
    lib.same_set_host_wait.argtypes = [i32]
    lib.same_debug_guard.argtypes = [i32, C.POINTER(i64), C.POINTER(i64)]
    lib.same_mempool_stats.argtypes = [i32, C.POINTER(i64), C.POINTER(i64)]
    lib.same_profile_report.argtypes = [C.c_char_p, i64]
    lib.same_profile_report.restype = i64
    assert lib.same_abi_version() == 1
    _lib = lib
    return lib


def check(rc):
    if rc != OK:
        raise SameError(rc, load().same_last_error().decode("utf-8", "replace"))


def ptr(a):
    """void* of a numpy array (None -> NULL) or a raw device address (int)."""
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    return a.ctypes.data_as(C.c_void_p)


def fp64_peak_tflops(device: int = 0) -> float:
    v = C.c_double(0.0)
    check(load().same_measure_fp64_peak(device, C.byref(v)))
    return v.value


def launch_count() -> int:
    return int(load().same_launch_count())


def mempool_stats(device=0):
    """-> (reserved, used) bytes of the stream-ordered device memory pool the library allocates from."""
    r, u = C.c_int64(0), C.c_int64(0)
    check(load().same_mempool_stats(int(device), C.byref(r), C.byref(u)))
    return r.value, u.value


def set_host_wait(yield_core: bool):
    """How host threads wait for the GPU: spin (False, default, lowest latency) or sleep on a blocking event (True)."""
    check(load().same_set_host_wait(int(bool(yield_core))))


def debug_guard(enable=None):
    """The library's own memory checker (same_debug_guard): enable True / False switches guard mode for buffers allocated from now
    on, None leaves it, "selftest" overruns a guarded probe buffer on purpose (the first count must rise by one);
    -> (buffers found with a corrupted canary zone, buffers checked) since the library was loaded."""
    c, k = C.c_int64(0), C.c_int64(0)
    check(load().same_debug_guard(-1 if enable is None else (2 if enable == "selftest" else int(bool(enable))), C.byref(c), C.byref(k)))
    return int(c.value), int(k.value)


def mempool_reserve(device, nbytes):
    """Grow the device memory pool by `nbytes` of headroom now (same_mempool_reserve)."""
    check(load().same_mempool_reserve(int(device), int(nbytes)))


def profile_enable(on: bool):
    check(load().same_profile_enable(int(bool(on))))


def profile_report():
    """-> {kernel name: (launches, total_ms)} since profiling was enabled / last reported."""
    buf = C.create_string_buffer(1 << 16)
    n = load().same_profile_report(buf, len(buf))
    if n < 0:
        check(int(n))
    out = {}
    for line in buf.value.decode().splitlines():
        name, cnt, ms = line.split("\t")
        out[name] = (int(cnt), float(ms))
    return out
