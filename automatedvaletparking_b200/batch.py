"""Batched planning on one GPU: the host side of the C ABI for many scenarios.

    with DevicePlanner(config) as dp:
        dp.load(scenarios)            # upload + rasterise (costmap.Map for every scenario)
        res = dp.plan()               # PathPlanner.a_star_plan for every scenario, on the device

`res` is a PlanResults: a structured numpy array of per-scenario summaries
(hostcfg.SUMMARY_DTYPE), padded (n, cap_path, 3) paths and (n, cap_pops) pop-index traces.
"""
import ctypes
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np

from . import _native
from .hostcfg import SUMMARY_DTYPE, STATUS_NAMES, make_avp_config
from . import scenarios as scn


class AvpError(RuntimeError):
    pass


@dataclass
class PlanResults:
    summaries: np.ndarray          # (n,) SUMMARY_DTYPE
    paths: np.ndarray              # (n, cap_path, 3) f64, rows beyond n_final are unspecified
    pops: Optional[np.ndarray]     # (n, cap_pops) i32 or None

    def path(self, i: int) -> np.ndarray:
        return self.paths[i, :min(int(self.summaries["n_final"][i]), self.paths.shape[1])]

    def astar_path(self, i: int) -> np.ndarray:
        return self.paths[i, :int(self.summaries["n_astar"][i])]

    def pop_indices(self, i: int) -> np.ndarray:
        return self.pops[i, :min(int(self.summaries["n_pops"][i]), self.pops.shape[1])]

    def status_name(self, i: int) -> str:
        return STATUS_NAMES.get(int(self.summaries["status"][i]), "?")

    @property
    def successors(self) -> int:
        """collision-checked successor evaluations = sum of global_index (hybrid_a_star.py:239)"""
        return int(self.summaries["global_index"].astype(np.int64).sum())


def _pi_2_pi(theta: float) -> float:
    """rs_curve.pi_2_pi (rs_curve.py:649-656)"""
    import math
    while theta > math.pi:
        theta -= 2.0 * math.pi
    while theta < -math.pi:
        theta += 2.0 * math.pi
    return theta


def _dp(a):
    return a.ctypes.data_as(_native.c_dp)


def _ip(a):
    return a.ctypes.data_as(_native.c_ip)


class DevicePlanner:
    def __init__(self, config: dict = None, vehicle=None, device: int = 0, max_pops: int = 20000):
        self._L = _native.lib()
        self.cfg = make_avp_config(config, vehicle, max_pops=max_pops)
        h = ctypes.c_void_p()
        rc = self._L.avp_create(int(device), ctypes.byref(self.cfg), ctypes.byref(h))
        if rc != 0 or not h:
            raise AvpError(f"avp_create failed (rc={rc}): no usable CUDA device {device}? this package has no CPU path")
        self._h = h
        self.device = int(device)
        self.plan_epoch = 0            # bumped by every whole search and every upload: they overwrite the per-cell arrays of the single-step Dijkstra
        self.n = 0
        self.batch = None
        self.h2d_bytes = 0
        self.d2h_bytes = 0

    # -- lifetime
    def close(self):
        if getattr(self, "_h", None):
            self._L.avp_destroy(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc, what):
        if rc != 0:
            raise AvpError(f"{what} failed (rc={rc}): {self._L.avp_last_error(self._h).decode()}")

    # -- scenarios / maps
    def load(self, scenarios, rasterise: bool = True):
        b = scenarios if isinstance(scenarios, scn.ScenarioBatch) else scn.pack(scenarios)
        self.batch = b
        self.plan_epoch += 1
        self._ck(self._L.avp_scenarios_upload(self._h, len(b), _dp(b.poses), _ip(b.obs_off), _ip(b.nv), _ip(b.vert_off),
                                              _dp(b.verts), _dp(b.boundary) if b.boundary is not None else None),
                 "avp_scenarios_upload")
        self.n = len(b)
        self.h2d_bytes += b.nbytes()
        if rasterise:
            self._ck(self._L.avp_rasterise(self._h), "avp_rasterise")

    def map_info(self, s: int):
        dims = np.zeros(4, dtype=np.int32)
        geom = np.zeros(6)
        self._ck(self._L.avp_fetch_map(self._h, s, _ip(dims), _dp(geom), None, 0), "avp_fetch_map")
        return dict(nx=int(dims[0]), ny=int(dims[1]), n_obs=int(dims[2]), raster_error=int(dims[3]),
                    boundary=geom[:4].copy(), dx=float(geom[4]), dy=float(geom[5]))

    def cost_map(self, s: int) -> np.ndarray:
        info = self.map_info(s)
        out = np.zeros(info["nx"] * info["ny"], dtype=np.uint8)
        dims = np.zeros(4, dtype=np.int32)
        geom = np.zeros(6)
        self._ck(self._L.avp_fetch_map(self._h, s, _ip(dims), _dp(geom), out.ctypes.data_as(_native.c_u8p), out.size),
                 "avp_fetch_map")
        return out.reshape(info["nx"], info["ny"])

    # -- leaf entry points
    def check(self, s: int, poses) -> np.ndarray:
        p = np.ascontiguousarray(np.asarray(poses, dtype=np.float64).reshape(-1, 3))
        out = np.zeros(p.shape[0], dtype=np.uint8)
        self._ck(self._L.avp_collision_check(self._h, s, p.shape[0], _dp(p), out.ctypes.data_as(_native.c_u8p)),
                 "avp_collision_check")
        return out.astype(bool)

    def corridor(self, s: int, poses, expand_dis: float):
        """compute_collision_H distances per path point: (m, 4) = x_max, y_max, x_min, y_min and a status vector
        (1 = heading outside [-pi, pi], where the reference raises)"""
        p = np.ascontiguousarray(np.asarray(poses, dtype=np.float64).reshape(-1, 3))
        out = np.zeros((p.shape[0], 4))
        st = np.zeros(p.shape[0], dtype=np.int32)
        self._ck(self._L.avp_corridor(self._h, s, p.shape[0], _dp(p), float(expand_dis), _dp(out), _ip(st)), "avp_corridor")
        return out, st

    def start_goal_collisions(self):
        """distance_checker.check at every loaded scenario's start and goal pose (scenario recipes); one launch for the batch.
        The search wraps headings with pi_2_pi (hybrid_a_star.py:105,109); the check sees the same values."""
        out = np.zeros(2 * self.n, dtype=np.uint8)
        self._ck(self._L.avp_check_start_goal(self._h, out.ctypes.data_as(_native.c_u8p)), "avp_check_start_goal")
        o = out.reshape(self.n, 2).astype(bool)
        return o[:, 0].copy(), o[:, 1].copy()

    def expand_pure(self, s: int, parent):
        n = 2 * self.cfg.steering_angle_num
        p = np.array(parent, dtype=np.float64)
        pose = np.zeros((n, 3))
        flags = np.zeros(n, dtype=np.int32)
        rsl = np.zeros(n)
        self._ck(self._L.avp_expand_pure(self._h, s, _dp(p), _dp(pose), _ip(flags), _dp(rsl)), "avp_expand_pure")
        return pose, flags, rsl

    def rs_optimal(self, q, maxc: float, step_size: float = 0.5, xy_np: int = 1, phi_np: int = 1, cap_pts: int = 256):
        q = np.ascontiguousarray(np.asarray(q, dtype=np.float64).reshape(-1, 6))
        m = q.shape[0]
        lengths = np.zeros((m, 5))
        ct = ctypes.create_string_buffer(8 * m)
        nseg = np.zeros(m, dtype=np.int32)
        L = np.zeros(m)
        x, y, yaw = np.zeros((m, cap_pts)), np.zeros((m, cap_pts)), np.zeros((m, cap_pts))
        d = np.zeros((m, cap_pts), dtype=np.int32)
        npts = np.zeros(m, dtype=np.int32)
        self._ck(self._L.avp_rs_optimal(self._h, m, _dp(q), float(maxc), float(step_size), int(xy_np), int(phi_np),
                                        _dp(lengths), ct, _ip(nseg), _dp(L), cap_pts, _dp(x), _dp(y), _dp(yaw), _ip(d),
                                        _ip(npts)), "avp_rs_optimal")
        names = [ct.raw[8 * i:8 * i + 8].split(b"\0")[0].decode() for i in range(m)]
        return dict(nseg=nseg, lengths=lengths, ctypes=names, L=L, x=x, y=y, yaw=yaw, directions=d, n_pts=npts)

    # -- whole searches
    def plan(self, cap_path: int = 512, cap_pops: int = 0) -> PlanResults:
        n = self.n
        sums = np.zeros(n, dtype=SUMMARY_DTYPE)
        paths = np.zeros((n, cap_path, 3))
        pops = np.zeros((n, cap_pops), dtype=np.int32) if cap_pops > 0 else None
        self.plan_epoch += 1
        self._ck(self._L.avp_plan_batch(self._h, sums.ctypes.data_as(ctypes.c_void_p), _dp(paths), cap_path,
                                        _ip(pops) if pops is not None else None, cap_pops), "avp_plan_batch")
        self.d2h_bytes += sums.nbytes + paths.nbytes + (pops.nbytes if pops is not None else 0)
        return PlanResults(sums, paths, pops)

    def plan_resident(self, cap_path: int = 512, cap_pops: int = 0) -> float:
        """Run the search with results left on the device; returns the CUDA-event time in ms."""
        self._ck(self._L.avp_plan_configure(self._h, cap_path, cap_pops), "avp_plan_configure")
        ms = ctypes.c_float()
        self.plan_epoch += 1
        self._ck(self._L.avp_plan_batch_resident(self._h, ctypes.byref(ms)), "avp_plan_batch_resident")
        return float(ms.value)

    def fetch(self, cap_path: int = 512, cap_pops: int = 0) -> PlanResults:
        n = self.n
        sums = np.zeros(n, dtype=SUMMARY_DTYPE)
        paths = np.zeros((n, cap_path, 3))
        pops = np.zeros((n, cap_pops), dtype=np.int32) if cap_pops > 0 else None
        self._ck(self._L.avp_fetch_results(self._h, sums.ctypes.data_as(ctypes.c_void_p), _dp(paths), cap_path,
                                           _ip(pops) if pops is not None else None, cap_pops), "avp_fetch_results")
        return PlanResults(sums, paths, pops)

    def split_paths(self, cap_pts: int = None, cap_seg: int = 32):
        """PathPlanner.split_path (path_planner.py:112-192) for every finished plan of the last plan()/plan_resident(): one launch.
        -> dict(status (n,), n_seg (n,), change_gear (n,), n_pts (n,), seg_len (n, cap_seg), pts (n, cap_pts, 3));
        status: 0 ok, 1 no gear change (the reference raises IndexError), 2 the scenario has no path, 3 capacity."""
        n = self.n
        cap_pts = int(cap_pts or 768)
        pts = np.zeros((n, cap_pts, 3)); seg = np.zeros((n, cap_seg), dtype=np.int32); info = np.zeros((n, 4), dtype=np.int32)
        self._ck(self._L.avp_split_paths(self._h, cap_pts, cap_seg, _dp(pts), _ip(seg), _ip(info)), "avp_split_paths")
        self.d2h_bytes += pts.nbytes + seg.nbytes + info.nbytes
        return dict(status=info[:, 0].copy(), n_seg=info[:, 1].copy(), change_gear=info[:, 2].copy(), n_pts=info[:, 3].copy(), seg_len=seg, pts=pts)

    def split_path(self, s: int, path, cap_seg: int = 64):
        """PathPlanner.split_path(final_path) for one caller-supplied path against scenario s's raster
        -> (status, list of segments as (len, 3) arrays, change_gear)"""
        p = np.ascontiguousarray(np.asarray(path, dtype=np.float64).reshape(-1, 3))
        cap_pts = len(p) + cap_seg * (1 + 2 * int(self.cfg.extended_num)) + 8
        pts = np.zeros((cap_pts, 3)); seg = np.zeros(cap_seg, dtype=np.int32); info = np.zeros(4, dtype=np.int32)
        self._ck(self._L.avp_split_path(self._h, s, len(p), _dp(p), cap_pts, cap_seg, _dp(pts), _ip(seg), _ip(info)), "avp_split_path")
        st, ns, cg, npt = (int(v) for v in info)
        segs, o = [], 0
        for k in range(min(ns, cap_seg)):
            segs.append(pts[o:o + int(seg[k])].copy()); o += int(seg[k])
        return st, segs, cg

    def trace_fgh(self, on: bool = True):
        """record f, g, h of every popped node in the following plans (parity aid; needs cap_pops > 0)"""
        self._ck(self._L.avp_trace_fgh(self._h, 1 if on else 0), "avp_trace_fgh")

    def pop_fgh(self, cap_pops: int) -> np.ndarray:
        out = np.zeros((self.n, cap_pops, 3))
        self._ck(self._L.avp_fetch_pop_fgh(self._h, _dp(out), cap_pops), "avp_fetch_pop_fgh")
        return out

    def result_device_pointers(self):
        s, p = ctypes.c_void_p(), ctypes.c_void_p()
        n, cp = ctypes.c_int64(), ctypes.c_int64()
        self._ck(self._L.avp_result_device_buffer(self._h, ctypes.byref(s), ctypes.byref(p), ctypes.byref(n), ctypes.byref(cp)),
                 "avp_result_device_buffer")
        return s.value, p.value, n.value, cp.value

    def hvalues(self, s: int) -> np.ndarray:
        n_ids = ctypes.c_int64()
        self._ck(self._L.avp_fetch_hvalues(self._h, s, None, 0, ctypes.byref(n_ids)), "avp_fetch_hvalues")
        out = np.zeros(n_ids.value, dtype=np.int32)
        self._ck(self._L.avp_fetch_hvalues(self._h, s, _ip(out), out.size, ctypes.byref(n_ids)), "avp_fetch_hvalues")
        return out

    def hq_log(self, s: int, n_hq: int) -> np.ndarray:
        out = np.zeros((256, 3), dtype=np.int32)
        self._ck(self._L.avp_fetch_hq_log(self._h, s, _ip(out), 256), "avp_fetch_hq_log")
        return out[:min(n_hq, 256)]

    def rasterise(self):
        self._ck(self._L.avp_rasterise(self._h), "avp_rasterise")

    def timer_start(self):
        self._ck(self._L.avp_timer_start(self._h), "avp_timer_start")

    def timer_stop(self) -> float:
        ms = ctypes.c_float()
        self._ck(self._L.avp_timer_stop(self._h, ctypes.byref(ms)), "avp_timer_stop")
        return float(ms.value)

    def last_search_ms(self) -> float:
        ms = ctypes.c_float()
        self._ck(self._L.avp_last_search_ms(self._h, ctypes.byref(ms)), "avp_last_search_ms")
        return float(ms.value)

    def last_search_passes(self):
        a, b, n = ctypes.c_float(), ctypes.c_float(), ctypes.c_int32()
        self._ck(self._L.avp_last_search_passes(self._h, ctypes.byref(a), ctypes.byref(b), ctypes.byref(n)), "avp_last_search_passes")
        return float(a.value), float(b.value), int(n.value) % 100000, int(n.value) // 100000

    def last_narrow_ms(self) -> float:
        """CUDA-event time of the narrow first search launch of the last plan (0: one launch did everything)"""
        ms = ctypes.c_float()
        self._ck(self._L.avp_last_narrow_ms(self._h, ctypes.byref(ms)), "avp_last_narrow_ms")
        return float(ms.value)

    def dijkstra_query(self, s: int, x: float, y: float, reset: bool = False):
        """Dijkstra.compute_path(x, y) of scenario s -> (distance, len(closedlist), terminate_grid_id)"""
        d, c, t = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32()
        self._ck(self._L.avp_dijkstra_query(self._h, s, 1 if reset else 0, float(x), float(y), ctypes.byref(d), ctypes.byref(c), ctypes.byref(t)),
                 "avp_dijkstra_query")
        return d.value, c.value, t.value

    def phase_profile(self) -> np.ndarray:
        out = np.zeros((self.n, 16), dtype=np.int64)
        self._ck(self._L.avp_fetch_profile(self._h, out.ctypes.data_as(_native.c_lp)), "avp_fetch_profile")
        return out

    def warp_profile(self) -> np.ndarray:
        out = np.zeros((self.n, 16, 24), dtype=np.int64)
        self._ck(self._L.avp_fetch_warp_profile(self._h, out.ctypes.data_as(_native.c_lp)), "avp_fetch_warp_profile")
        return out

    def set_watchdog(self, cycles: int):
        self._ck(self._L.avp_set_watchdog(self._h, int(cycles)), "avp_set_watchdog")

    @property
    def launches(self) -> int:
        return int(self._L.avp_launch_count(self._h))

    def device_info(self):
        a, b, c = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32()
        self._L.avp_device_info(self._h, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c))
        return dict(n_sm=a.value, slots=b.value, block=c.value)
