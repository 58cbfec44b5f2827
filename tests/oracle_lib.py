"""ctypes wrapper of the CPU oracle (oracle/avp_oracle.c).  Test infrastructure only.

Builds oracle/_build/libavp_oracle.so with `make -C oracle` when missing or stale.
"""
import ctypes
import os
import subprocess

import numpy as np

from automatedvaletparking_b200.hostcfg import AvpConfig, AvpPlanSummary, SUMMARY_DTYPE, make_avp_config
from automatedvaletparking_b200 import scenarios as scn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
SO = os.path.join(ORACLE_DIR, "_build", "libavp_oracle.so")

c_dp = ctypes.POINTER(ctypes.c_double)
c_ip = ctypes.POINTER(ctypes.c_int32)
c_lp = ctypes.POINTER(ctypes.c_int64)


class PlanOut(ctypes.Structure):
    _fields_ = [
        ("sum", AvpPlanSummary),
        ("pops", c_ip), ("cap_pops", ctypes.c_int),
        ("pop_state", c_dp), ("pop_fgh", c_dp),
        ("final_path", c_dp), ("cap_path", ctypes.c_int),
        ("rs_x", c_dp), ("rs_y", c_dp), ("rs_yaw", c_dp), ("rs_dir", c_ip), ("cap_rs", ctypes.c_int),
        ("hq_log", c_lp), ("cap_hq", ctypes.c_int),
        ("hval_out", c_ip), ("hval_cap", ctypes.c_long),
    ]


def build(force=False):
    src = os.path.join(ORACLE_DIR, "avp_oracle.c")
    hdr = os.path.join(ROOT, "include", "avp_b200.h")
    stale = (not os.path.exists(SO)) or os.path.getmtime(SO) < max(os.path.getmtime(src), os.path.getmtime(hdr))
    if force or stale:
        subprocess.run(["make", "-C", ORACLE_DIR], check=True, capture_output=True)
    return SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = ctypes.CDLL(build())
        L.orc_map_build.restype = ctypes.c_void_p
        L.orc_map_build.argtypes = [c_dp, ctypes.c_int, c_ip, c_dp, ctypes.c_double, c_dp]
        L.orc_map_free.argtypes = [ctypes.c_void_p]
        L.orc_map_info.argtypes = [ctypes.c_void_p, c_ip, c_dp, c_ip, c_ip]
        L.orc_map_cost.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        L.orc_map_positions.argtypes = [ctypes.c_void_p, c_dp, c_dp]
        L.orc_convert_position_to_index.restype = ctypes.c_long
        L.orc_convert_position_to_index.argtypes = [ctypes.c_void_p, ctypes.c_double, ctypes.c_double]
        L.orc_check.restype = ctypes.c_int
        L.orc_check.argtypes = [ctypes.c_void_p, ctypes.POINTER(AvpConfig), ctypes.c_double, ctypes.c_double, ctypes.c_double]
        L.orc_vehicle_corners.argtypes = [ctypes.POINTER(AvpConfig), ctypes.c_double, ctypes.c_double, ctypes.c_double, c_dp]
        L.orc_pi_2_pi.restype = ctypes.c_double
        L.orc_pi_2_pi.argtypes = [ctypes.c_double]
        L.orc_py_hypot.restype = ctypes.c_double
        L.orc_py_hypot.argtypes = [ctypes.c_double, ctypes.c_double]
        L.orc_rs_words.restype = ctypes.c_int
        L.orc_rs_words.argtypes = [c_dp, c_dp, ctypes.c_double, ctypes.c_int, ctypes.c_int, c_ip, c_dp, ctypes.c_char_p, c_dp]
        L.orc_rs_optimal.restype = ctypes.c_int
        L.orc_rs_optimal.argtypes = [c_dp, c_dp, ctypes.c_double, ctypes.c_double, ctypes.c_int, ctypes.c_int, c_ip, c_dp, ctypes.c_char_p, c_dp,
                                     ctypes.c_int, c_dp, c_dp, c_dp, c_ip, c_ip]
        L.orc_dij_new.restype = ctypes.c_void_p
        L.orc_dij_new.argtypes = [ctypes.c_void_p]
        L.orc_dij_free.argtypes = [ctypes.c_void_p]
        L.orc_dij_compute_path.restype = ctypes.c_int
        L.orc_dij_compute_path.argtypes = [ctypes.c_void_p, ctypes.c_double, ctypes.c_double]
        L.orc_dij_closed_len.restype = ctypes.c_long
        L.orc_dij_closed_len.argtypes = [ctypes.c_void_p]
        L.orc_dij_n_ids.restype = ctypes.c_long
        L.orc_dij_n_ids.argtypes = [ctypes.c_void_p]
        L.orc_dij_hvalues.argtypes = [ctypes.c_void_p, c_ip]
        L.orc_split_path.argtypes = [ctypes.c_void_p, ctypes.POINTER(AvpConfig), ctypes.c_int, c_dp, c_dp, ctypes.c_int, c_ip, ctypes.c_int, c_ip]
        L.orc_plan.restype = ctypes.c_int
        L.orc_plan.argtypes = [ctypes.c_void_p, ctypes.POINTER(AvpConfig), ctypes.POINTER(PlanOut)]
        L.orc_corridor.argtypes = [ctypes.c_void_p, ctypes.POINTER(AvpConfig), ctypes.c_double, ctypes.c_int, c_dp, c_dp, c_ip]
        L.orc_expand_pure.argtypes = [ctypes.c_void_p, ctypes.POINTER(AvpConfig), c_dp, c_dp, c_ip, c_dp]
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(c_dp)


def _ip(a):
    return a.ctypes.data_as(c_ip)


class OracleMap:
    """Map(file, discrete_size) restated (costmap.py:159-261)."""

    def __init__(self, scenario: scn.Scenario, discrete_size: float = 0.1):
        self.scenario = scenario
        pose = np.array(scenario.pose, dtype=np.float64)
        nv = np.array([len(o) for o in scenario.obs], dtype=np.int32)
        verts = (np.concatenate([np.asarray(o, dtype=np.float64).reshape(-1, 2) for o in scenario.obs], axis=0)
                 if len(scenario.obs) else np.zeros((0, 2)))
        verts = np.ascontiguousarray(verts)
        bo = None
        if scenario.boundary is not None:
            bo = np.array(scenario.boundary, dtype=np.float64)
        self._h = lib().orc_map_build(_dp(pose), len(nv), _ip(nv), _dp(verts), float(discrete_size),
                                      _dp(bo) if bo is not None else None)
        dims = np.zeros(2, dtype=np.int32)
        geom = np.zeros(6)
        n_obs = ctypes.c_int32()
        err = ctypes.c_int32()
        lib().orc_map_info(self._h, _ip(dims), _dp(geom), ctypes.byref(n_obs), ctypes.byref(err))
        self.nx, self.ny = int(dims[0]), int(dims[1])
        self.boundary = geom[:4].copy()
        self.dx, self.dy = float(geom[4]), float(geom[5])
        self.n_obs = n_obs.value
        self.raster_error = err.value

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_map_free(self._h)
            self._h = None

    def cost_map(self):
        out = np.zeros(self.nx * self.ny, dtype=np.uint8)
        lib().orc_map_cost(self._h, out.ctypes.data_as(ctypes.c_void_p))
        return out.reshape(self.nx, self.ny)

    def positions(self):
        xs, ys = np.zeros(self.nx), np.zeros(self.ny)
        lib().orc_map_positions(self._h, _dp(xs), _dp(ys))
        return xs, ys

    def index(self, x, y):
        return lib().orc_convert_position_to_index(self._h, float(x), float(y))

    def check(self, cfg, x, y, theta):
        return bool(lib().orc_check(self._h, ctypes.byref(cfg), float(x), float(y), float(theta)))

    def corridor(self, cfg, poses, expand_dis):
        """compute_collision_H distances (x_max, y_max, x_min, y_min) per pose + status (1: heading outside [-pi, pi])"""
        p = np.ascontiguousarray(np.asarray(poses, dtype=np.float64).reshape(-1, 3))
        out = np.zeros((p.shape[0], 4))
        st = np.zeros(p.shape[0], dtype=np.int32)
        lib().orc_corridor(self._h, ctypes.byref(cfg), float(expand_dis), p.shape[0], _dp(p), _dp(out), _ip(st))
        return out, st

    def split_path(self, cfg, path, cap_pts=None, cap_seg=64):
        """PathPlanner.split_path(final_path) -> dict(status, segments: list of (len, 3) arrays, change_gear, out_final_path)"""
        p = np.ascontiguousarray(np.asarray(path, dtype=np.float64).reshape(-1, 3))
        cap_pts = cap_pts or (len(p) + 64 * (1 + 2 * int(cfg.extended_num)) + 8)
        out = np.zeros((cap_pts, 3)); seg = np.zeros(cap_seg, dtype=np.int32); info = np.zeros(4, dtype=np.int32)
        lib().orc_split_path(self._h, ctypes.byref(cfg), len(p), _dp(p), _dp(out), cap_pts, _ip(seg), cap_seg, _ip(info))
        st, ns, cg, npt = (int(v) for v in info)
        segs, o = [], 0
        for k in range(min(ns, cap_seg)):
            segs.append(out[o:o + int(seg[k])].copy()); o += int(seg[k])
        return dict(status=st, segments=segs, change_gear=cg, out_final_path=out[:min(npt, cap_pts)].copy(), seg_len=seg[:ns].copy())

    def expand_pure(self, cfg, parent):
        n = 2 * cfg.steering_angle_num
        p = np.array(parent, dtype=np.float64)
        pose = np.zeros((n, 3))
        flags = np.zeros(n, dtype=np.int32)
        rsl = np.zeros(n)
        lib().orc_expand_pure(self._h, ctypes.byref(cfg), _dp(p), _dp(pose), _ip(flags), _dp(rsl))
        return pose, flags, rsl


class OracleDijkstra:
    def __init__(self, m: OracleMap):
        self.m = m
        self._h = lib().orc_dij_new(m._h)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_dij_free(self._h)
            self._h = None

    def compute_path(self, x, y):
        return lib().orc_dij_compute_path(self._h, float(x), float(y)), lib().orc_dij_closed_len(self._h)

    def hvalues(self):
        n = lib().orc_dij_n_ids(self._h)
        out = np.zeros(n, dtype=np.int32)
        lib().orc_dij_hvalues(self._h, _ip(out))
        return out


def vehicle_corners(cfg, x, y, th):
    out = np.zeros(10)
    lib().orc_vehicle_corners(ctypes.byref(cfg), float(x), float(y), float(th), _dp(out))
    return out.reshape(5, 2)


def rs_words(q0, q1, maxc, xy_np=1, phi_np=1):
    q0 = np.array(q0, dtype=np.float64)
    q1 = np.array(q1, dtype=np.float64)
    nseg = np.zeros(64, dtype=np.int32)
    lengths = np.zeros((64, 5))
    ct = ctypes.create_string_buffer(64 * 8)
    L = np.zeros(64)
    n = lib().orc_rs_words(_dp(q0), _dp(q1), float(maxc), int(xy_np), int(phi_np), _ip(nseg), _dp(lengths), ct, _dp(L))
    degenerate = n < 0
    if degenerate:
        n = -1 - n
    words = []
    for i in range(n):
        name = ct.raw[8 * i:8 * i + 8].split(b"\0")[0].decode()
        words.append((name, lengths[i, :nseg[i]].copy(), float(L[i])))
    return words, degenerate


def rs_optimal(q0, q1, maxc, step=0.5, cap=2048, xy_np=1, phi_np=1):
    q0 = np.array(q0, dtype=np.float64)
    q1 = np.array(q1, dtype=np.float64)
    nseg = ctypes.c_int32()
    lengths = np.zeros(5)
    ct = ctypes.create_string_buffer(8)
    L = ctypes.c_double()
    x, y, yaw = np.zeros(cap), np.zeros(cap), np.zeros(cap)
    d = np.zeros(cap, dtype=np.int32)
    npts = ctypes.c_int32()
    rc = lib().orc_rs_optimal(_dp(q0), _dp(q1), float(maxc), float(step), int(xy_np), int(phi_np), ctypes.byref(nseg), _dp(lengths), ct,
                              ctypes.byref(L), cap, _dp(x), _dp(y), _dp(yaw), _ip(d), ctypes.byref(npts))
    n = npts.value
    return dict(rc=rc, ctypes=ct.value.decode(), lengths=lengths[:nseg.value].copy(), L=L.value,
                x=x[:n].copy(), y=y[:n].copy(), yaw=yaw[:n].copy(), directions=d[:n].copy())


def plan(m: OracleMap, cfg, cap_pops=None, cap_path=4096, want_hval=False):
    cap_pops = cap_pops or max(1, cfg.max_pops)
    out = PlanOut()
    pops = np.zeros(cap_pops, dtype=np.int32)
    pop_state = np.zeros((cap_pops, 3))
    pop_fgh = np.zeros((cap_pops, 3))
    fp = np.zeros((cap_path, 3))
    rs_x, rs_y, rs_yaw = np.zeros(2048), np.zeros(2048), np.zeros(2048)
    rs_dir = np.zeros(2048, dtype=np.int32)
    hq = np.zeros((4096, 3), dtype=np.int64)
    out.pops, out.cap_pops = _ip(pops), cap_pops
    out.pop_state, out.pop_fgh = _dp(pop_state), _dp(pop_fgh)
    out.final_path, out.cap_path = _dp(fp), cap_path
    out.rs_x, out.rs_y, out.rs_yaw, out.rs_dir, out.cap_rs = _dp(rs_x), _dp(rs_y), _dp(rs_yaw), _ip(rs_dir), 2048
    out.hq_log, out.cap_hq = hq.ctypes.data_as(c_lp), 4096
    hval = None
    if want_hval:
        hval = np.full((m.nx + 4) * (m.ny + 4) + 64, -1, dtype=np.int32)
        out.hval_out, out.hval_cap = _ip(hval), hval.size
    lib().orc_plan(m._h, ctypes.byref(cfg), ctypes.byref(out))
    s = out.sum
    npop = min(s.n_pops, cap_pops)
    return dict(status=s.status, n_pops=s.n_pops, global_index=s.global_index, n_closed=s.n_closed, n_open=s.n_open,
                n_astar=s.n_astar, n_rs=s.n_rs, n_final=s.n_final, rs_ctypes=s.rs_ctypes.decode(),
                rs_lengths=np.array(s.rs_lengths[:s.rs_nseg]), rs_L=s.rs_L, n_hq=s.n_hq, h_closed=s.h_closed,
                n_hcalls=s.n_hcalls, pops=pops[:npop].copy(), pop_state=pop_state[:npop].copy(),
                pop_fgh=pop_fgh[:npop].copy(), final_path=fp[:min(s.n_final, cap_path)].copy(),
                rs_x=rs_x[:s.n_rs].copy(), rs_y=rs_y[:s.n_rs].copy(), rs_yaw=rs_yaw[:s.n_rs].copy(),
                rs_dir=rs_dir[:s.n_rs].copy(), hq=hq[:min(s.n_hq, 4096)].copy(), last_index=s.last_index, hval=hval)
