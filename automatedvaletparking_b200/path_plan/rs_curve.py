"""Drop-in for path_plan/rs_curve.py: PATH, calc_optimal_path, pi_2_pi, M (rs_curve.py:86-680).
calc_optimal_path evaluates the 46 word instances, the set_path de-duplication, the last-minimum
selection and the 0.5 m course of the selected word on the GPU (avp_rs_optimal)."""
import math

import numpy as np

STEP_SIZE = 0.5
MAX_LENGTH = 1000.0
PI = math.pi

_dev = None


class PATH:
    def __init__(self, lengths, ctypes, L, x, y, yaw, directions):
        self.lengths = lengths
        self.ctypes = ctypes
        self.L = L
        self.x = x
        self.y = y
        self.yaw = yaw
        self.directions = directions


def _device():
    global _dev
    if _dev is None:
        from ..batch import DevicePlanner
        _dev = DevicePlanner()
    return _dev


def _is_np(v) -> bool:
    return isinstance(v, np.floating)


def calc_optimal_path(sx, sy, syaw, gx, gy, gyaw, maxc, step_size=STEP_SIZE) -> PATH:
    # CPython's sum() treats numpy scalars and exact floats differently (see include/avp_b200.h):
    # x, y are numpy scalars iff any of their inputs is; phi = gyaw - syaw likewise
    xy_np = int(_is_np(maxc) or _is_np(sx) or _is_np(sy) or _is_np(gx) or _is_np(gy))
    phi_np = int(_is_np(syaw) or _is_np(gyaw))
    r = _device().rs_optimal([[sx, sy, syaw, gx, gy, gyaw]], float(maxc), float(step_size), xy_np, phi_np, cap_pts=2048)
    n = int(r['nseg'][0])
    if n == -2:
        raise AssertionError("path.L >= 0.01")           # rs_curve.py:153
    if n < 0:
        raise IndexError("list index out of range")      # paths[0] on an empty list (rs_curve.py:103)
    k = int(r['n_pts'][0])
    return PATH([float(v) for v in r['lengths'][0][:n]], list(r['ctypes'][0]), float(r['L'][0]), [float(v) for v in r['x'][0][:k]],
                [float(v) for v in r['y'][0][:k]], [float(v) for v in r['yaw'][0][:k]], [int(v) for v in r['directions'][0][:k]])


def pi_2_pi(theta):
    while theta > PI:
        theta -= 2.0 * PI
    while theta < -PI:
        theta += 2.0 * PI
    return theta


def M(theta):
    phi = theta % (2.0 * PI)
    if phi < -PI:
        phi += 2.0 * PI
    if phi > PI:
        phi -= 2.0 * PI
    return phi
