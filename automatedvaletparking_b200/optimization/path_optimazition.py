"""Drop-in for the obstacle-raster part of optimization/path_optimazition.py.

`path_opti.compute_collision_H()` keeps the reference's name, inputs (`self.original_path`) and return value
(`H_collision_matrix`, `slack_H_collision_matrix`, path_optimazition.py:221-658); the scan of the raster
cells around every path point runs on the GPU (avp_corridor).  The QP itself (formate_matrix / get_result,
cvxopt) is downstream of the hot path and not part of this package: those methods raise.
"""
from typing import List

import numpy as np

from ..map.costmap import Map, Vehicle


def corridor_distances(park_map: Map, config: dict, path, vehicle: Vehicle = None):
    """(n, 4) array x_max, y_max, x_min, y_min per path point; raises like the reference for |theta| > pi."""
    dev = _device_for(park_map, config, vehicle)
    d, status = dev.corridor(0, [[float(p[0]), float(p[1]), float(p[2])] for p in path], float(config['expand_dis']))
    if status.any():
        # path_optimazition.py:343-350: no branch assigns `case`
        raise UnboundLocalError("cannot access local variable 'case' where it is not associated with a value")
    return d


def _device_for(park_map: Map, config: dict, vehicle: Vehicle = None):
    base = park_map._device
    if config is None or (base.cfg.safe_side_dis == float(config['safe_side_dis']) and base.cfg.safe_fr_dis == float(config['safe_fr_dis'])):
        return base
    key = (float(config['safe_side_dis']), float(config['safe_fr_dis']))
    cache = park_map.__dict__.setdefault('_corridor_devices', {})
    if key not in cache:                          # a context with this caller's vehicle inflation
        from ..batch import DevicePlanner
        want = dict(config)
        want['map_discrete_size'] = park_map.discrete_size
        cache[key] = DevicePlanner(want, vehicle)
        cache[key].load([park_map.scenario])
    return cache[key]


class path_opti:
    def __init__(self, park_map: Map, vehicle: Vehicle, config: dict) -> None:
        self.original_path = None
        self.map = park_map
        self.vehicle = vehicle
        self.matrix_dict = dict()
        self.expand_dis = config['expand_dis']  # m
        self.config = config

    def compute_collision_H(self):
        """[E;-E] X <= [H_max;-H_min] (path_optimazition.py:221-658)"""
        path = self.original_path
        points_n = len(path)
        d = corridor_distances(self.map, self.config, path, self.vehicle)
        H_max, H_min = [], []
        for p, (x_max, y_max, x_min, y_min) in zip(path, d):
            x, y = p[0], p[1]
            H_max.append(x_max + x)                       # :651-654
            H_max.append(y_max + y)
            H_min.append(x - x_min)
            H_min.append(y - y_min)
        H_max_matrix = np.array(H_max).reshape(len(H_max), 1)
        H_min_matrix = np.array(H_min).reshape(len(H_min), 1)
        H_collision_matrix = np.vstack((H_max_matrix, -H_min_matrix))
        slack_H_collision_matrix = np.vstack((H_max_matrix, 999 * np.ones((points_n - 2, 1)),
                                              -H_min_matrix, np.zeros((points_n - 2, 1))))
        return H_collision_matrix, slack_H_collision_matrix

    def formate_matrix(self, path: list):
        raise NotImplementedError("the QP assembly / cvxopt solve (path_optimazition.py:33-219) is outside the hot path; "
                                  "use the reference class and delegate compute_collision_H to this one (INTEGRATION.md)")

    def get_result(self, path) -> List[List]:
        raise NotImplementedError("see formate_matrix")
