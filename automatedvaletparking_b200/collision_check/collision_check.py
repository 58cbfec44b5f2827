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
        """the device context whose derived avp_config (config AND vehicle) equals this checker's; the Map's own
        context (default config, default Vehicle) is reused only when every field matches"""
        if self._dev is None:
            from ..batch import DevicePlanner
            from ..hostcfg import make_avp_config
            if self.config is None:
                if self._mode == 'circle':
                    raise ValueError("two_circle_checker needs a config (collision_check.py:88-92 reads config['safe_side_dis'])")
                self._dev = self.map._device          # distance_checker(config=None): the reference only fails later, in check()
                return self._dev
            want = dict(self.config)
            want['collision_check'] = self._mode
            want['map_discrete_size'] = self.map.discrete_size
            base = self.map._device
            wcfg = make_avp_config(want, self.vehicle, max_pops=base.cfg.max_pops)
            if bytes(wcfg) == bytes(base.cfg):
                self._dev = base
            else:                               # a context with this checker's mode / inflation / vehicle / rollout constants
                self._dev = DevicePlanner(want, self.vehicle, device=getattr(base, "device", 0), max_pops=base.cfg.max_pops)
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
