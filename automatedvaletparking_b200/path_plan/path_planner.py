"""Drop-in for path_plan/path_planner.py: PathPlanner (path_planner.py:25-192).

a_star_plan() runs the WHOLE search on the GPU (avp_plan_batch: raster, lazy Dijkstra heuristic,
successor expansion, rs shots) and returns the reference's tuple (final_path, astar_path, PATH);
path_planning() adds split_path (gear-change splitting with collision-checked extension points,
path_planner.py:112-192), also computed on the device (avp_split_path).  A search the reference cannot finish raises
the same exception class the reference raises (AttributeError for an exhausted open list)."""
import copy
from typing import Dict, List, Tuple

import numpy as np

from .hybrid_a_star import hybrid_a_star
from ..map.costmap import Vehicle, Map
from ..collision_check import collision_check
from .rs_curve import PATH
from . import rs_curve


class PlanningCapacityError(RuntimeError):
    """the search exceeded max_pops (the reference would keep running)"""


class PathPlanner:
    def __init__(self, config: dict = None, map: Map = None, vehicle: Vehicle = None, max_pops: int = 20000, verbose: bool = False) -> None:
        self.config = config
        self.map = map
        self.vehicle = vehicle
        self.verbose = verbose
        self.max_pops = max_pops
        if config['collision_check'] == 'circle':
            self.collision_checker = collision_check.two_circle_checker(map=map, vehicle=vehicle, config=config)
        else:
            self.collision_checker = collision_check.distance_checker(map=map, vehicle=vehicle, config=config)
        self._planner = None
        self._ctx = None
        self.last_summary = None
        self.pop_indices = None

    def _context(self):
        """ONE device context for this planner's whole searches, shared with the collision checker when the derived
        avp_config matches (same config, vehicle and max_pops) -- not one per a_star_plan call"""
        if self._ctx is None:
            from ..batch import DevicePlanner
            from ..hostcfg import make_avp_config
            chk = self.collision_checker._device()
            want = dict(self.config)
            want['map_discrete_size'] = self.map.discrete_size
            if bytes(make_avp_config(want, self.vehicle, max_pops=self.max_pops)) == bytes(chk.cfg):
                self._ctx = chk
            else:
                self._ctx = DevicePlanner(want, self.vehicle, device=getattr(chk, "device", 0), max_pops=self.max_pops)
        return self._ctx

    @property
    def planner(self) -> hybrid_a_star:
        """the reference builds hybrid_a_star eagerly (path_planner.py:42-43); here the step-wise object is
        only materialised when somebody asks for it (a_star_plan does not need it)"""
        if self._planner is None:
            self._planner = hybrid_a_star(config=self.config, park_map=self.map, vehicle=self.vehicle)
        return self._planner

    def path_planning(self) -> Tuple[List[List], Dict, List[List[List]]]:
        final_path, astar_path, rs_path = self.a_star_plan()
        split_path_list, change_gear = self.split_path(final_path)
        path_info = {'astar_path': astar_path, 'rs_path': rs_path, 'change_gear': change_gear}
        out_final_path = sum(split_path_list, [])
        return out_final_path, path_info, split_path_list

    def a_star_plan(self) -> Tuple[List[List], List[List], PATH]:
        dev = self._context()
        dev.load([self.map.scenario])
        res = dev.plan(cap_path=4096, cap_pops=self.max_pops)
        s = res.summaries[0]
        self.last_summary = s
        self.pop_indices = res.pop_indices(0).copy()
        if self.verbose:                                   # the reference prints every pop (path_planner.py:72-76)
            for idx in self.pop_indices:
                print('---------------')
                print('current node index:', int(idx))
                print('---------------')
        status = int(s['status'])
        if status == 1:
            raise AttributeError("'NoneType' object has no attribute 'x'")        # path_planner.py:104 on an exhausted open list
        if status == 3:
            raise RuntimeError("heuristic target unreachable: the reference blocks forever (compute_h.py:77)")
        if status == 4:
            raise AssertionError("path.L >= 0.01")                               # rs_curve.py:153
        if status == 5:
            raise PlanningCapacityError(f"no path within max_pops={self.max_pops}")
        if status == 6:
            raise TypeError("only size-1 arrays can be converted to Python scalars")
        path = res.path(0)
        n_astar, n_rs = int(s['n_astar']), int(s['n_rs'])
        a_star_path = [[np.float64(p[0]), np.float64(p[1]), np.float64(p[2])] for p in path[:n_astar]]
        final_path = copy.deepcopy(a_star_path) + [[float(p[0]), float(p[1]), float(p[2])] for p in path[n_astar:]]
        # the PATH object of the successful shot: word, lengths and course.  Point 0 of the course is the popped node's own pose
        # (rs_curve.py:118-132) -- not the last finish_path row, which is re-rolled with 3 * ddt != dt and can differ in the last bit
        lp = [float(v) for v in s['last_pose']]
        nseg = int(s['rs_nseg'])
        rs_x = [lp[0]] + [float(p[0]) for p in path[n_astar:]]
        rs_y = [lp[1]] + [float(p[1]) for p in path[n_astar:]]
        rs_yaw = [lp[2]] + [float(p[2]) for p in path[n_astar:]]
        full = rs_curve.calc_optimal_path(np.float64(lp[0]), np.float64(lp[1]), np.float64(lp[2]) if int(s['last_index']) != 0 else lp[2],
                                          self.map.case.xf, self.map.case.yf, rs_curve.pi_2_pi(self.map.case.thetaf), 1 / self.vehicle.min_radius_turn)
        rs_path = PATH([float(v) for v in s['rs_lengths'][:nseg]], list(s['rs_ctypes'].decode()), float(s['rs_L']), rs_x, rs_y, rs_yaw,
                       full.directions if len(full.directions) == n_rs else [0] * n_rs)
        return final_path, a_star_path, rs_path

    def split_path(self, final_path: List[List]) -> Tuple[List[List[List]], int]:
        """path_planner.py:112-192 on the device (avp_split_path: gear-change detection with scipy's cosine semantics, collision-checked
        extension points, segment assembly); returns the reference's (split_path_list, change_gear).  A path without a gear
        change raises IndexError as the reference does (:181, split_path[-1] on an empty list)."""
        dev = self.collision_checker._device()           # the context whose config / vehicle / inflation the checker uses
        status, segs, change_gear = dev.split_path(0, final_path)
        if status == 1:
            raise IndexError("list index out of range")
        if status != 0:
            raise RuntimeError(f"split_path: device status {status}")
        split = [[[float(r[0]), float(r[1]), float(r[2])] for r in seg] for seg in segs]
        return split, int(change_gear)
