"""Drop-in for path_plan/compute_h.py: Grid, Dijkstra (compute_h.py:20-255).  compute_path is the
stateful, resumable goal-rooted Dijkstra of the reference, run on the GPU with an exact emulation of
CPython's heapq (avp_dijkstra_query)."""
import numpy as np

from ..map.costmap import Map


class Grid:
    def __init__(self, grid_id: int, grid_x=None, grid_y=None, distance: int = 0, father_id: int = 0) -> None:
        self.grid_id = grid_id
        self.grid_x = grid_x
        self.grid_y = grid_y
        self.distance = distance
        self.father_id = father_id

    def __lt__(self, other):
        if self.distance == other.distance:
            return self.grid_id < other.grid_id
        return self.distance < other.distance


class _ClosedList:
    """len() == len(Dijkstra.closedlist); iteration yields one Grid per visited id with the distance of
    its first closedlist entry (what calc_node_heuristic reads, hybrid_a_star.py:272-280)."""

    def __init__(self, owner):
        self._owner = owner

    def __len__(self):
        return self._owner._closed_len

    def lookup(self, grid_id):
        hv = self._owner._hvalues()
        return int(hv[grid_id]) if 0 <= grid_id < len(hv) and hv[grid_id] >= 0 else None

    def __iter__(self):
        hv = self._owner._hvalues()
        for gid in np.nonzero(hv >= 0)[0]:
            yield Grid(int(gid), distance=int(hv[gid]))


class Dijkstra:
    """Every Dijkstra object owns its queue, closed list and h table, like the reference's: the first object on a Map uses the
    Map's device context, any further one (and any object made after that context ran a whole search, which re-uses the
    per-cell arrays) gets a context of its own."""

    def __init__(self, map: Map) -> None:
        self.map = map
        self.final_point = (map.case.xf, map.case.yf, map.case.thetaf)
        self.find_terminate = False
        self._closed_len = 0
        self._fresh = True
        self._hv = None
        self.closedlist = _ClosedList(self)
        self.terminate_grid_id = None
        owner = getattr(map, "_dijkstra_owner", None)
        if owner is None or owner() is None:
            import weakref
            self._dev = map._device
            map._dijkstra_owner = weakref.ref(self)
        else:
            from ..batch import DevicePlanner
            from ..hostcfg import default_config
            cfg = dict(default_config())
            cfg['map_discrete_size'] = map.discrete_size
            self._dev = DevicePlanner(cfg, device=getattr(map._device, "device", 0))
            self._dev.load([map.scenario])
        self._epoch = self._dev.plan_epoch

    def _hvalues(self):
        if self._hv is None:
            self._hv = self._dev.hvalues(0)
        return self._hv

    def compute_path(self, node_x, node_y):
        if self._dev.plan_epoch != self._epoch and not self._fresh:
            raise RuntimeError("this Dijkstra object's device state was overwritten by a whole search on the same context "
                               "(PathPlanner.a_star_plan); create the Dijkstra object after planning, or plan on another Map")
        d, closed, term = self._dev.dijkstra_query(0, float(node_x), float(node_y), reset=self._fresh)
        self._epoch = self._dev.plan_epoch
        self._fresh = False
        self._hv = None
        self._closed_len = closed
        self.terminate_grid_id = term
        if d < 0:
            raise RuntimeError("heuristic target unreachable: the reference blocks forever in queue.get() (compute_h.py:77)")
        self.find_terminate = True
        return d, self.closedlist
