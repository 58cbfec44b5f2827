"""Host-side configuration: config dict + Vehicle constants -> the C ABI's avp_config.

Everything that the reference evaluates with numpy/CPython *before* the hot loop is
evaluated here with the same numpy/CPython expressions, so the device only performs IEEE
add/mul/div on the results:
  steer      = np.linspace(-max_steer, max_steer, n)        (hybrid_a_star.py:81-83)
  tan_steer  = np.tan(steer)                                 (hybrid_a_star.py:147)
  n_substeps = math.ceil(dt / trajectory_dt)                 (hybrid_a_star.py:185)
  min_radius_turn = lw / np.tan(max_steer) + lb / 2          (costmap.py:62-63)
"""
import ctypes
import math

import numpy as np

AVP_MAX_STEER = 16
AVP_MAX_RS_SEG = 5


class AvpConfig(ctypes.Structure):
    _fields_ = [
        ("steering_angle_num", ctypes.c_int32), ("n_substeps", ctypes.c_int32),
        ("dt", ctypes.c_double), ("ddt", ctypes.c_double), ("map_discrete_size", ctypes.c_double),
        ("flag_radius", ctypes.c_double), ("extended_num", ctypes.c_int32), ("collision_mode", ctypes.c_int32),
        ("cost_gear", ctypes.c_double), ("cost_heading_change", ctypes.c_double), ("cost_scale", ctypes.c_double),
        ("safe_side_dis", ctypes.c_double), ("safe_fr_dis", ctypes.c_double),
        ("lw", ctypes.c_double), ("lf", ctypes.c_double), ("lr", ctypes.c_double), ("lb", ctypes.c_double),
        ("max_steering_angle", ctypes.c_double), ("max_v", ctypes.c_double), ("min_radius_turn", ctypes.c_double),
        ("steer", ctypes.c_double * AVP_MAX_STEER), ("tan_steer", ctypes.c_double * AVP_MAX_STEER),
        ("max_pops", ctypes.c_int32), ("reserved", ctypes.c_int32),
    ]


class AvpPlanSummary(ctypes.Structure):
    _fields_ = [
        ("status", ctypes.c_int32), ("n_pops", ctypes.c_int32), ("global_index", ctypes.c_int32),
        ("n_closed", ctypes.c_int32), ("n_open", ctypes.c_int32), ("n_astar", ctypes.c_int32),
        ("n_rs", ctypes.c_int32), ("n_final", ctypes.c_int32), ("rs_nseg", ctypes.c_int32),
        ("last_index", ctypes.c_int32), ("n_hq", ctypes.c_int32), ("h_closed", ctypes.c_int32),
        ("nx", ctypes.c_int32), ("ny", ctypes.c_int32), ("n_obs", ctypes.c_int32), ("n_hcalls", ctypes.c_int32),
        ("rs_L", ctypes.c_double), ("rs_lengths", ctypes.c_double * AVP_MAX_RS_SEG),
        ("rs_ctypes", ctypes.c_char * 8), ("origin", ctypes.c_double * 2), ("pitch", ctypes.c_double * 2),
        ("boundary", ctypes.c_double * 4), ("last_pose", ctypes.c_double * 3),
    ]


STATUS_NAMES = {0: "OK", 1: "OPEN_EXHAUSTED", 2: "OPEN_EXHAUSTED_RS", 3: "H_UNREACHABLE",
                4: "RS_DEGENERATE", 5: "CAPACITY", 6: "RASTER_AMBIGUOUS"}

SUMMARY_DTYPE = np.dtype([
    ("status", "<i4"), ("n_pops", "<i4"), ("global_index", "<i4"), ("n_closed", "<i4"), ("n_open", "<i4"),
    ("n_astar", "<i4"), ("n_rs", "<i4"), ("n_final", "<i4"), ("rs_nseg", "<i4"), ("last_index", "<i4"),
    ("n_hq", "<i4"), ("h_closed", "<i4"), ("nx", "<i4"), ("ny", "<i4"), ("n_obs", "<i4"), ("n_hcalls", "<i4"),
    ("rs_L", "<f8"), ("rs_lengths", "<f8", (AVP_MAX_RS_SEG,)), ("rs_ctypes", "S8"),
    ("origin", "<f8", (2,)), ("pitch", "<f8", (2,)), ("boundary", "<f8", (4,)), ("last_pose", "<f8", (3,)),
])
assert SUMMARY_DTYPE.itemsize == ctypes.sizeof(AvpPlanSummary)


class VehicleConstants:
    """Vehicle() constants (costmap.py:52-63)."""

    def __init__(self):
        self.lw = 2.8
        self.lf = 0.96
        self.lr = 0.929
        self.lb = 1.942
        self.max_steering_angle = 0.75
        self.max_angular_velocity = 0.5
        self.max_acc = 1
        self.max_v = 2.5
        self.min_v = -2.5
        self.min_radius_turn = self.lw / np.tan(self.max_steering_angle) + self.lb / 2


def default_config() -> dict:
    from .config.read_config import read_config
    return read_config("config")


def make_avp_config(config: dict = None, vehicle=None, max_pops: int = 20000) -> AvpConfig:
    config = default_config() if config is None else config
    v = VehicleConstants() if vehicle is None else vehicle
    c = AvpConfig()
    n = int(config['steering_angle_num'])
    if not 1 <= n <= AVP_MAX_STEER:
        raise ValueError("steering_angle_num out of range")
    c.steering_angle_num = n
    c.n_substeps = int(math.ceil(config['dt'] / config['trajectory_dt']))
    c.dt = float(config['dt'])
    c.ddt = float(config['trajectory_dt'])
    c.map_discrete_size = float(config['map_discrete_size'])
    c.flag_radius = float(config['flag_radius'])
    c.extended_num = int(config['extended_num'])
    c.collision_mode = 1 if config['collision_check'] == 'circle' else 0
    c.cost_gear = float(config['cost_gear'])
    c.cost_heading_change = float(config['cost_heading_change'])
    c.cost_scale = float(config['cost_scale'])
    c.safe_side_dis = float(config['safe_side_dis'])
    c.safe_fr_dis = float(config['safe_fr_dis'])
    c.lw, c.lf, c.lr, c.lb = float(v.lw), float(v.lf), float(v.lr), float(v.lb)
    c.max_steering_angle = float(v.max_steering_angle)
    c.max_v = float(v.max_v)
    c.min_radius_turn = float(v.min_radius_turn)
    steer = np.linspace(-v.max_steering_angle, v.max_steering_angle, n)
    for i in range(n):
        c.steer[i] = float(steer[i])
        c.tan_steer[i] = float(np.tan(steer[i]))
    c.max_pops = int(max_pops)
    return c
