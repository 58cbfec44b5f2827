"""Scenario containers: Case CSV rows <-> flat arrays for the C ABI, and the seeded
scenario recipes of BASELINE.json's configs.

A scenario is what the reference's `Case.read` extracts from one CSV row
(costmap.py:134-156): start pose, goal pose, and a list of obstacle polygons.
"""
import csv
import io
import os
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

_CASES_JSON = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "benchmark_cases.json")
_cases_cache = None


@dataclass
class Scenario:
    x0: float
    y0: float
    theta0: float
    xf: float
    yf: float
    thetaf: float
    obs: List[np.ndarray] = field(default_factory=list)      # each (nv, 2) float64
    boundary: Optional[Sequence[float]] = None               # override (config C4), else floor(min/max -/+ 12)
    name: str = ""

    @property
    def pose(self):
        return (self.x0, self.y0, self.theta0, self.xf, self.yf, self.thetaf)


def parse_case_row(values: Sequence[float], name: str = "") -> Scenario:
    """One CSV row -> Scenario; same slicing as costmap.py:140-155."""
    v = [float(i) for i in values]
    obs_num = int(v[6])
    nvs = np.array(v[7:7 + obs_num], dtype=np.int32)
    starts = 7 + obs_num + (np.cumsum(nvs, dtype=np.int32) - nvs) * 2
    obs = [np.array(v[s:s + n * 2], dtype=np.float64).reshape((n, 2)) for s, n in zip(starts, nvs)]
    return Scenario(v[0], v[1], v[2], v[3], v[4], v[5], obs, None, name)


def read_case_csv(path: str) -> Scenario:
    with open(path, "r") as f:
        rows = list(csv.reader(f))
    return parse_case_row(rows[0], os.path.splitext(os.path.basename(path))[0])


def case_row(s: Scenario) -> List[float]:
    row = [s.x0, s.y0, s.theta0, s.xf, s.yf, s.thetaf, float(len(s.obs))]
    row += [float(len(o)) for o in s.obs]
    for o in s.obs:
        row += [float(t) for t in np.asarray(o, dtype=np.float64).reshape(-1)]
    return row


def write_case_csv(s: Scenario, path: str) -> None:
    """Write with repr(float) so every consumer parses bit-identical doubles."""
    with open(path, "w") as f:
        f.write(",".join(repr(float(t)) for t in case_row(s)) + "\n")


def benchmark_case(n: int) -> Scenario:
    """The 20 TPCAP scenarios the reference ships as BenchmarkCases/Case{n}.csv (kept as the
    original text tokens in data/benchmark_cases.json so float() parses identical doubles)."""
    global _cases_cache
    if _cases_cache is None:
        import json
        with open(_CASES_JSON) as f:
            _cases_cache = json.load(f)["cases"]
    return parse_case_row(_cases_cache[f"Case{n}"], f"Case{n}")


@dataclass
class ScenarioBatch:
    """Flat arrays in the layout avp_scenarios_upload expects (include/avp_b200.h)."""
    poses: np.ndarray        # (n, 6) f64
    obs_off: np.ndarray      # (n+1,) i32
    nv: np.ndarray           # (n_obs_total,) i32
    vert_off: np.ndarray     # (n_obs_total+1,) i32
    verts: np.ndarray        # (n_vert_total, 2) f64
    boundary: Optional[np.ndarray]  # (n, 4) f64 or None
    names: List[str]

    def __len__(self):
        return self.poses.shape[0]

    def nbytes(self) -> int:
        b = self.poses.nbytes + self.obs_off.nbytes + self.nv.nbytes + self.vert_off.nbytes + self.verts.nbytes
        return b + (self.boundary.nbytes if self.boundary is not None else 0)


def pack(scenarios: Sequence[Scenario]) -> ScenarioBatch:
    n = len(scenarios)
    poses = np.array([s.pose for s in scenarios], dtype=np.float64).reshape(n, 6)
    obs_off = np.zeros(n + 1, dtype=np.int32)
    nv, verts = [], []
    for i, s in enumerate(scenarios):
        obs_off[i + 1] = obs_off[i] + len(s.obs)
        for o in s.obs:
            o = np.asarray(o, dtype=np.float64).reshape(-1, 2)
            nv.append(o.shape[0])
            verts.append(o)
    nv = np.array(nv, dtype=np.int32)
    vert_off = np.zeros(len(nv) + 1, dtype=np.int32)
    np.cumsum(nv, out=vert_off[1:])
    verts = np.concatenate(verts, axis=0) if verts else np.zeros((0, 2))
    has_b = [s.boundary is not None for s in scenarios]
    boundary = None
    if any(has_b):
        if not all(has_b):
            raise ValueError("boundary override must be given for all scenarios of a batch or none")
        boundary = np.array([s.boundary for s in scenarios], dtype=np.float64).reshape(n, 4)
    return ScenarioBatch(np.ascontiguousarray(poses), obs_off, nv, vert_off, np.ascontiguousarray(verts),
                         boundary, [s.name for s in scenarios])


# --------------------------------------------------------------------------- seeded recipes

def _axis_aligned(theta: float) -> bool:
    # the reference's slope/intercept rectangle test degenerates at multiples of pi/2
    # (collision_check.py:149-160): excluded by the recipe
    k = theta / (np.pi / 2)
    return float(k) == float(np.round(k)) and abs(k - np.round(k)) == 0.0


def perturb(base: Scenario, rng: np.random.Generator, checker=None, max_tries: int = 200) -> Scenario:
    """SURVEY §8d C2 recipe: start += U(-2,2) m / U(-pi/4,pi/4); goal += U(-0.5,0.5) m / U(-0.2,0.2) rad.
    `checker(scenario) -> (start_collides, goal_collides)` is optional (rejection sampling)."""
    for _ in range(max_tries):
        d = rng.uniform(-1.0, 1.0, size=6)
        s = Scenario(base.x0 + 2.0 * d[0], base.y0 + 2.0 * d[1], base.theta0 + (np.pi / 4) * d[2],
                     base.xf + 0.5 * d[3], base.yf + 0.5 * d[4], base.thetaf + 0.2 * d[5],
                     base.obs, None, base.name)
        s = Scenario(*(float(t) for t in s.pose), s.obs, None, base.name)
        if np.hypot(s.x0 - s.xf, s.y0 - s.yf) < 1.0:
            continue
        if _axis_aligned(s.theta0) or _axis_aligned(s.thetaf):
            continue
        if checker is not None:
            a, b = checker(s)
            if a or b:
                continue
        return s
    raise RuntimeError("could not draw a collision-free perturbation")


def perturbed_set(base: Scenario, count: int, seed: int, checker=None) -> List[Scenario]:
    rng = np.random.default_rng(seed)
    out = []
    for i in range(count):
        s = perturb(base, rng, checker)
        s.name = f"{base.name}_p{i}"
        out.append(s)
    return out


def perturbed_candidates(base: Scenario, m: int, seed: int) -> List[Scenario]:
    """m draws of the C2 recipe WITHOUT the collision rule (batch-friendly: the caller checks all
    candidates' start/goal poses at once, e.g. on the GPU, and keeps the first `count` free ones)."""
    out = perturbed_set(base, m, seed, checker=None)
    for i, s in enumerate(out):
        s.name = f"{base.name}_c{i}"
    return out


def keep_collision_free(cands: Sequence[Scenario], start_collides, goal_collides, count: int) -> List[Scenario]:
    out = [s for s, a, b in zip(cands, start_collides, goal_collides) if not (a or b)]
    if len(out) < count:
        raise RuntimeError(f"only {len(out)} of {len(cands)} candidates are collision free, need {count}")
    return out[:count]


def synthetic_map_obstacles(rng: np.random.Generator, n_poly: int = 256, extent: float = 20.0) -> List[np.ndarray]:
    """convex polygons with 3-8 vertices on a circle of radius U(0.10,0.35) m."""
    obs = []
    for _ in range(n_poly):
        nv = int(rng.integers(3, 9))
        r = rng.uniform(0.10, 0.35)
        cx, cy = rng.uniform(0.5, extent - 0.5, size=2)
        ang = np.sort(rng.uniform(0.0, 2 * np.pi, size=nv))
        obs.append(np.stack([cx + r * np.cos(ang), cy + r * np.sin(ang)], axis=1))
    return obs


def synthetic_set(n_maps: int, pairs_per_map: int, seed: int = 4, extent: float = 20.0, n_poly: int = 256,
                  checker=None) -> List[Scenario]:
    """boundary [0,extent]^2 override, random poses in [4,16]^2 x U(-pi,pi)."""
    rng = np.random.default_rng(seed)
    out = []
    for m in range(n_maps):
        obs = synthetic_map_obstacles(rng, n_poly, extent)
        k = 0
        tries = 0
        while k < pairs_per_map:
            tries += 1
            if tries > 100 * pairs_per_map:
                raise RuntimeError("synthetic map too cluttered")
            p = rng.uniform(0.0, 1.0, size=6)
            lo, hi = 0.2 * extent, 0.8 * extent
            s = Scenario(float(lo + (hi - lo) * p[0]), float(lo + (hi - lo) * p[1]), float(-np.pi + 2 * np.pi * p[2]),
                         float(lo + (hi - lo) * p[3]), float(lo + (hi - lo) * p[4]), float(-np.pi + 2 * np.pi * p[5]),
                         obs, (0.0, extent, 0.0, extent), f"syn{m}_{k}")
            if np.hypot(s.x0 - s.xf, s.y0 - s.yf) < 1.0 or _axis_aligned(s.theta0) or _axis_aligned(s.thetaf):
                continue
            if checker is not None:
                a, b = checker(s)
                if a or b:
                    continue
            out.append(s)
            k += 1
    return out


def synthetic_candidates(n_maps: int, cand_per_map: int, seed: int = 4, extent: float = 20.0, n_poly: int = 256) -> List[List[Scenario]]:
    """SURVEY §8d C4, batch-friendly: per map the obstacles and `cand_per_map` start/goal draws WITHOUT the collision rule (the
    caller checks all candidates at once, e.g. on the GPU, and keeps the first free ones of every map).  One list per map."""
    rng = np.random.default_rng(seed)
    out = []
    for m in range(n_maps):
        obs = synthetic_map_obstacles(rng, n_poly, extent)
        lo, hi = 0.2 * extent, 0.8 * extent
        cands = []
        while len(cands) < cand_per_map:
            p = rng.uniform(0.0, 1.0, size=6)
            s = Scenario(float(lo + (hi - lo) * p[0]), float(lo + (hi - lo) * p[1]), float(-np.pi + 2 * np.pi * p[2]),
                         float(lo + (hi - lo) * p[3]), float(lo + (hi - lo) * p[4]), float(-np.pi + 2 * np.pi * p[5]),
                         obs, (0.0, extent, 0.0, extent), f"syn{m}_c{len(cands)}")
            if np.hypot(s.x0 - s.xf, s.y0 - s.yf) < 1.0 or _axis_aligned(s.theta0) or _axis_aligned(s.thetaf):
                continue
            cands.append(s)
        out.append(cands)
    return out
