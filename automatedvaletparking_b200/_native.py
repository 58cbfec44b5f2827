"""ctypes binding of the C ABI (include/avp_b200.h) exported by libavp_b200.so.

The library is built in-tree by `__graft_entry__.build()` (nvcc, sm_100a).  There is no
CPU fallback: a missing library or a machine without a CUDA device raises.
"""
import ctypes
import os

from .hostcfg import AvpConfig, AvpPlanSummary

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("AVP_B200_LIB") or os.path.join(_HERE, "libavp_b200.so")     # env override: A/B builds (development aid)

c_dp = ctypes.POINTER(ctypes.c_double)
c_ip = ctypes.POINTER(ctypes.c_int32)
c_u8p = ctypes.POINTER(ctypes.c_uint8)
c_lp = ctypes.POINTER(ctypes.c_int64)
c_fp = ctypes.POINTER(ctypes.c_float)
c_vp = ctypes.c_void_p

EXPORTS = [
    "avp_create", "avp_destroy", "avp_last_error", "avp_launch_count", "avp_scenarios_upload", "avp_rasterise",
    "avp_fetch_map", "avp_collision_check", "avp_check_start_goal", "avp_corridor", "avp_expand_pure", "avp_rs_optimal", "avp_plan_batch",
    "avp_plan_batch_resident", "avp_fetch_results", "avp_plan_configure", "avp_result_device_buffer",
    "avp_split_paths", "avp_split_path", "avp_trace_fgh", "avp_fetch_pop_fgh", "avp_fetch_hvalues", "avp_fetch_hq_log", "avp_device_info", "avp_set_watchdog", "avp_fetch_debug", "avp_timer_start", "avp_timer_stop", "avp_last_search_ms", "avp_last_search_passes", "avp_last_narrow_ms", "avp_fetch_profile", "avp_fetch_warp_profile", "avp_dijkstra_query",
]

_lib = None


class NativeLibraryMissing(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NativeLibraryMissing(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). This package has no CPU fallback.")
    L = ctypes.CDLL(LIB_PATH)
    P = ctypes.POINTER
    L.avp_create.restype = ctypes.c_int
    L.avp_create.argtypes = [ctypes.c_int, P(AvpConfig), P(c_vp)]
    L.avp_destroy.argtypes = [c_vp]
    L.avp_last_error.restype = ctypes.c_char_p
    L.avp_last_error.argtypes = [c_vp]
    L.avp_launch_count.restype = ctypes.c_int64
    L.avp_launch_count.argtypes = [c_vp]
    L.avp_scenarios_upload.argtypes = [c_vp, ctypes.c_int, c_dp, c_ip, c_ip, c_ip, c_dp, c_dp]
    L.avp_rasterise.argtypes = [c_vp]
    L.avp_fetch_map.argtypes = [c_vp, ctypes.c_int, c_ip, c_dp, c_u8p, ctypes.c_int64]
    L.avp_collision_check.argtypes = [c_vp, ctypes.c_int, ctypes.c_int, c_dp, c_u8p]
    L.avp_check_start_goal.argtypes = [c_vp, c_u8p]
    L.avp_corridor.argtypes = [c_vp, ctypes.c_int, ctypes.c_int, c_dp, ctypes.c_double, c_dp, c_ip]
    L.avp_expand_pure.argtypes = [c_vp, ctypes.c_int, c_dp, c_dp, c_ip, c_dp]
    L.avp_rs_optimal.argtypes = [c_vp, ctypes.c_int, c_dp, ctypes.c_double, ctypes.c_double, ctypes.c_int, ctypes.c_int,
                                 c_dp, ctypes.c_char_p, c_ip, c_dp, ctypes.c_int, c_dp, c_dp, c_dp, c_ip, c_ip]
    L.avp_plan_batch.argtypes = [c_vp, c_vp, c_dp, ctypes.c_int, c_ip, ctypes.c_int]
    L.avp_plan_batch_resident.argtypes = [c_vp, c_fp]
    L.avp_fetch_results.argtypes = [c_vp, c_vp, c_dp, ctypes.c_int, c_ip, ctypes.c_int]
    L.avp_plan_configure.argtypes = [c_vp, ctypes.c_int, ctypes.c_int]
    L.avp_result_device_buffer.argtypes = [c_vp, P(c_vp), P(c_vp), c_lp, c_lp]
    L.avp_split_paths.argtypes = [c_vp, ctypes.c_int, ctypes.c_int, c_dp, c_ip, c_ip]
    L.avp_split_path.argtypes = [c_vp, ctypes.c_int, ctypes.c_int, c_dp, ctypes.c_int, ctypes.c_int, c_dp, c_ip, c_ip]
    L.avp_trace_fgh.argtypes = [c_vp, ctypes.c_int]
    L.avp_fetch_pop_fgh.argtypes = [c_vp, c_dp, ctypes.c_int]
    L.avp_fetch_hvalues.argtypes = [c_vp, ctypes.c_int, c_ip, ctypes.c_int64, c_lp]
    L.avp_fetch_hq_log.argtypes = [c_vp, ctypes.c_int, c_ip, ctypes.c_int]
    L.avp_device_info.argtypes = [c_vp, c_ip, c_ip, c_ip]
    L.avp_set_watchdog.argtypes = [c_vp, ctypes.c_longlong]
    L.avp_fetch_debug.argtypes = [c_vp, c_ip]
    L.avp_timer_start.argtypes = [c_vp]
    L.avp_timer_stop.argtypes = [c_vp, c_fp]
    L.avp_last_search_ms.argtypes = [c_vp, c_fp]
    L.avp_fetch_profile.argtypes = [c_vp, c_lp]
    L.avp_fetch_warp_profile.argtypes = [c_vp, c_lp]
    L.avp_dijkstra_query.argtypes = [c_vp, ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_double, c_ip, c_ip, c_ip]
    L.avp_last_search_passes.argtypes = [c_vp, c_fp, c_fp, c_ip]
    L.avp_last_narrow_ms.argtypes = [c_vp, c_fp]
    for name in EXPORTS:
        getattr(L, name)  # AttributeError here = the header and the library disagree
    _lib = L
    return L
