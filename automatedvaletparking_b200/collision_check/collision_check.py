"""Drop-in for collision_check/collision_check.py (collision_check.py:20-240): the checkers keep the
reference's constructor and `.check(node_x, node_y, theta) -> bool`; the test itself runs on the GPU
(avp_collision_check) against the map's raster."""
import numpy as np

from ..map.costmap import Map, Vehicle


class collision_checker:
    _mode = 'distance'

    def __init__(self, map: Map, vehicle: Vehicle = None, config: dict = None) -> None:
        self.map = map
        self.config = config
        self.vehicle = vehicle
        self._dev = None

    def _device(self):
        if self._dev is None:
            want = dict(self.config) if self.config is not None else None
            base = self.map._device
            same = want is None or (base.cfg.collision_mode == (1 if self._mode == 'circle' else 0)
                                    and base.cfg.safe_side_dis == float(want['safe_side_dis']) and base.cfg.safe_fr_dis == float(want['safe_fr_dis']))
            if same:
                self._dev = base
            else:                               # a context with this checker's mode / inflation
                from ..batch import DevicePlanner
                want['collision_check'] = self._mode
                want['map_discrete_size'] = self.map.discrete_size
                self._dev = DevicePlanner(want, self.vehicle)
                self._dev.load([self.map.scenario])
        return self._dev

    def check_many(self, poses) -> np.ndarray:
        return self._device().check(0, poses)

    def check(self, node_x, node_y, theta) -> bool:
        return bool(self._device().check(0, [[float(node_x), float(node_y), float(theta)]])[0])


class two_circle_checker(collision_checker):
    """two discs along the vehicle axis (collision_check.py:80-137)"""
    _mode = 'circle'


class distance_checker(collision_checker):
    """raster cells strictly inside the inflated rectangle (collision_check.py:140-240)"""
    _mode = 'distance'
