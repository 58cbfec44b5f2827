"""Drop-in for the obstacle-raster part of optimization/ocp_optimization.py:
`ocp_optimization.compute_collision_H(path)` (ocp_optimization.py:36-480) returns the reference's four lists
X_max, Y_max, X_min, Y_min; the raster scan runs on the GPU (avp_corridor).  The Pyomo / ipopt model
(`solution`, :482-676) is outside the hot path and raises."""
from ..map.costmap import Map, Vehicle
from ..path_plan import rs_curve
from .path_optimazition import corridor_distances


class ocp_optimization:
    def __init__(self, park_map: Map, vehicle: Vehicle, config: dict) -> None:
        self.map = park_map
        self.vehicle = vehicle
        self.config = config
        self.expand_dis = config['expand_dis']

    def compute_collision_H(self, path):
        wrapped = [[p[0], p[1], rs_curve.pi_2_pi(p[2])] for p in path]           # ocp_optimization.py:130
        d = corridor_distances(self.map, self.config, wrapped, Vehicle())          # :57 builds a fresh Vehicle()
        X_max, Y_max, X_min, Y_min = [], [], [], []
        for p, (x_max, y_max, x_min, y_min) in zip(path, d):
            x, y = p[0], p[1]
            X_max.append(x_max + x)                                                 # :466-469
            Y_max.append(y_max + y)
            X_min.append(x - x_min)
            Y_min.append(y - y_min)
        return X_max, Y_max, X_min, Y_min

    def solution(self, path: list):
        raise NotImplementedError("the Pyomo / ipopt OCP (ocp_optimization.py:482-676) is outside the hot path (DESIGN.md §7)")
