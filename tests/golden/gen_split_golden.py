"""Golden vectors for PathPlanner.split_path (path_planner.py:112-192) from the UNMODIFIED reference:
    python tests/golden/gen_split_golden.py      -> tests/golden/leaf_split.npz
The reference method is called on a PathPlanner whose __init__ is bypassed (it would run the eager Dijkstra, which
split_path does not use): split_path reads self.config, self.vehicle, self.planner.ddt and self.collision_checker only.
Paths: seeded random forward / reverse arc sequences on the maps of Cases 1, 4 and 13 (several gear changes, extension
points that collide and that do not), plus degenerate inputs: repeated points (scipy's cosine gives NaN), a path without a
gear change (IndexError), 2- and 3-point paths, near-perpendicular steps.  Also pins np.dot on 2-vectors = fma(a1, b1, a0*b0)."""
import os
import sys
import types
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

ref_shim.install()
warnings.filterwarnings("ignore")
from map import costmap  # noqa: E402
from path_plan import path_planner  # noqa: E402
from collision_check import collision_check  # noqa: E402

cfg = ref_shim.default_config()
rng = np.random.default_rng(21)


def planner_for(case):
    m = costmap.Map(file=os.path.join(ref_shim.REF_ROOT, cfg["Benchmark_path"], f"Case{case}.csv"), discrete_size=cfg["map_discrete_size"])
    v = costmap.Vehicle()
    pp = object.__new__(path_planner.PathPlanner)
    pp.config, pp.map, pp.vehicle = cfg, m, v
    pp.collision_checker = collision_check.distance_checker(map=m, vehicle=v, config=cfg)
    pp.planner = types.SimpleNamespace(ddt=cfg["trajectory_dt"])
    return pp, m


def random_path(m, n_arcs):
    c = m.case
    x, y, th = c.x0 + rng.uniform(-2, 2), c.y0 + rng.uniform(-2, 2), rng.uniform(-np.pi, np.pi)
    pts = [[x, y, th]]
    gear = 1.0
    for a in range(n_arcs):
        if a:
            gear = -gear if rng.uniform() < 0.75 else gear
        steer = rng.choice([-0.75, -0.375, 0.0, 0.375, 0.75])
        for k in range(int(rng.integers(2, 9))):
            th = th + 2.5 * np.tan(steer) / 2.8 * 0.2
            th = float((th + np.pi) % (2 * np.pi) - np.pi)
            x = x + gear * 2.5 * 0.2 * np.cos(th)
            y = y + gear * 2.5 * 0.2 * np.sin(th)
            pts.append([float(x), float(y), th])
    return pts


def run(pp, path):
    try:
        segs, cg = pp.split_path([list(p) for p in path])
        return 0, [len(s) for s in segs], int(cg), np.array(sum(segs, []), dtype=np.float64).reshape(-1, 3)
    except IndexError:
        return 1, [], 0, np.zeros((0, 3))


out = {}
k = 0
for case in (1, 4, 13):
    pp, m = planner_for(case)
    paths = [random_path(m, int(rng.integers(1, 7))) for _ in range(60)]
    base = random_path(m, 4)
    paths.append(base[:5] + [base[5]] * 3 + base[5:])                     # repeated point: NaN cosine
    paths.append(base[:2]); paths.append(base[:3]); paths.append(base[:1])
    paths.append([[0.0, 0.0, 0.3], [1.0, 0.0, 0.3], [1.0, 1e-17, 0.3], [1.0 + 1e-9, 1.0, 0.3], [0.5, 1.0, 0.3]])   # near-perpendicular steps
    paths.append([[c[0] + 1e9, c[1] - 1e9, c[2]] for c in base])           # far outside the raster: no collision anywhere
    for path in paths:
        st, lens, cg, flat = run(pp, path)
        out[f"p{k}_case"] = case
        out[f"p{k}_in"] = np.array(path, dtype=np.float64).reshape(-1, 3)
        out[f"p{k}_status"] = st
        out[f"p{k}_lens"] = np.array(lens, dtype=np.int32)
        out[f"p{k}_cg"] = cg
        out[f"p{k}_out"] = flat
        k += 1
out["n"] = k
# np.dot on 2-vectors (scipy's cosine): which rounding?
U = rng.normal(size=(4000, 2)); V = rng.normal(size=(4000, 2))
V[:2000] = np.stack([-U[:2000, 1], U[:2000, 0]], 1) * rng.normal(size=(2000, 1)) + rng.normal(size=(2000, 2)) * 1e-17
out["dot_u"], out["dot_v"] = U, V
out["dot_uv"] = np.array([np.dot(U[i], V[i]) for i in range(len(U))])
from scipy.spatial import distance  # noqa: E402
out["cosine"] = np.array([distance.cosine(tuple(U[i]), tuple(V[i])) for i in range(len(U))])
np.savez_compressed(os.path.join(HERE, "leaf_split.npz"), **out)
n_gc = [int(out[f"p{i}_cg"]) for i in range(k)]
print("paths", k, "gear changes hist", np.bincount(n_gc), "IndexError", sum(int(out[f"p{i}_status"]) for i in range(k)),
      "with extension points", sum(1 for i in range(k) if len(out[f"p{i}_out"]) > len(out[f"p{i}_in"]) + n_gc[i]))
