"""Drop-in for the reference's map/costmap.py: Vehicle, Case, Map (costmap.py:51-329).

Map keeps the reference's attribute surface (case, boundary, cost_map, map_position,
_discrete_x/_y, discrete_size, grid_index_max, convert_position_to_index); the raster itself
(discrete_map + detect_obstacle_edge, costmap.py:178-261) is computed on the GPU through the C ABI
(avp_scenarios_upload + avp_rasterise + avp_fetch_map).
"""
import math

import numpy as np

from .. import scenarios as scn
from ..hostcfg import VehicleConstants


class Vehicle(VehicleConstants):
    """costmap.Vehicle (costmap.py:51-121)"""

    def create_polygon(self, x, y, theta):
        """right back, right front, left front, left back, right back (costmap.py:65-83)"""
        c, s = np.cos(theta), np.sin(theta)
        local = np.array([[-self.lr, -self.lb / 2, 1], [self.lf + self.lw, -self.lb / 2, 1], [self.lf + self.lw, self.lb / 2, 1],
                          [-self.lr, self.lb / 2, 1], [-self.lr, -self.lb / 2, 1]])
        pose = np.array([[c, -s, x], [s, c, y], [0, 0, 1]])
        return local.dot(pose.transpose())[:, 0:2]

    def create_anticlockpoint(self, x, y, theta, config: dict = None):
        """inflated footprint corners, shape (5, 2, 1) like the reference (costmap.py:85-121)"""
        rot_t = np.array([[np.cos(theta), np.sin(theta)], [-np.sin(theta), np.cos(theta)]]).transpose()
        sd, fr = config['safe_side_dis'], config['safe_fr_dis']
        corners_local = [np.array([[-self.lr - fr], [-self.lb / 2 - sd]]), np.array([[self.lw + self.lf + fr], [-self.lb / 2 - sd]]),
                         np.array([[self.lw + self.lf + fr], [self.lb / 2 + sd]]), np.array([[-self.lr - fr], [self.lb / 2 + sd]])]
        origin = np.array([[x], [y]])
        pts = [rot_t.dot(p) + origin for p in corners_local]
        pts.append(pts[0])
        return np.array([[p[0], p[1]] for p in pts])


class Case:
    """costmap.Case (costmap.py:124-156)"""

    def __init__(self):
        self.x0, self.y0, self.theta0 = 0, 0, 0
        self.xf, self.yf, self.thetaf = 0, 0, 0
        self.xmin, self.xmax = 0, 0
        self.ymin, self.ymax = 0, 0
        self.obs_num = 0
        self.obs = np.array([])
        self.vehicle = Vehicle()

    @staticmethod
    def from_scenario(s: scn.Scenario):
        case = Case()
        case.x0, case.y0, case.theta0 = s.x0, s.y0, s.theta0
        case.xf, case.yf, case.thetaf = s.xf, s.yf, s.thetaf
        case.xmin = min(case.x0, case.xf) - 12
        case.xmax = max(case.x0, case.xf) + 12
        case.ymin = min(case.y0, case.yf) - 12
        case.ymax = max(case.y0, case.yf) + 12
        case.obs_num = len(s.obs)
        case.obs = [np.array(o, dtype=np.float64) for o in s.obs]
        case.scenario = s
        return case

    @staticmethod
    def read(file):
        return Case.from_scenario(scn.read_case_csv(file))


class Map:
    """costmap.Map(discrete_size=0.1, file=...) (costmap.py:159-329), rasterised on the GPU."""

    def __init__(self, discrete_size: np.float64 = 0.1, file: str = None, scenario: scn.Scenario = None, device: int = 0):
        from ..batch import DevicePlanner
        from ..hostcfg import default_config
        self.discrete_size = discrete_size
        self.grid_index = None
        if scenario is None:
            scenario = scn.read_case_csv(file)
        self.scenario = scenario
        self.case = Case.from_scenario(scenario)
        cfg = dict(default_config())
        cfg['map_discrete_size'] = discrete_size
        self._device = DevicePlanner(cfg, device=device)       # one single-scenario context per Map
        self._device.load([scenario])
        info = self._device.map_info(0)
        self.boundary = np.array(info['boundary'], dtype=np.float64)
        self._discrete_x = info['dx']
        self._discrete_y = info['dy']
        nx, ny = info['nx'], info['ny']
        self.cost_map = self._device.cost_map(0).astype(np.float64)           # indexed [ix][iy], values {0, 255}
        self.map_position = (np.linspace(self.boundary[0], self.boundary[1], nx), np.linspace(self.boundary[2], self.boundary[3], ny))
        self.grid_index_max = nx * ny
        if info['raster_error']:
            raise TypeError("only size-1 arrays can be converted to Python scalars")   # reference: costmap.py:260

    def convert_position_to_index(self, grid_x, grid_y):
        """costmap.py:319-329"""
        index_0 = math.floor((grid_x - self.boundary[0]) / self._discrete_x)
        index_1 = math.floor((self.boundary[3] - grid_y) / self._discrete_y) * (int((self.boundary[1] - self.boundary[0]) / self._discrete_x))
        return index_0 + index_1
