"""Known-answer vectors of the corridor extraction from the UNMODIFIED reference:
path_opti.compute_collision_H (optimization/path_optimazition.py:221-658) and
ocp_optimization.compute_collision_H (optimization/ocp_optimization.py:36-480).

    python tests/golden/gen_corridor_golden.py        -> tests/golden/leaf_corridor.npz

cvxopt / pyomo are not installed here; the two modules import them at module scope only, so empty stub
modules are enough (neither compute_collision_H touches them).  Inputs: the gear segments of the reference's
own path_planning() output for every BenchmarkCase it finishes (tests/golden/cases/*.npz), plus random poses.
"""
import os
import sys
import types
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
import ref_shim  # noqa: E402

ref_shim.install()
for name in ("cvxopt", "pyomo", "pyomo.environ", "pyomo.dae"):
    m = types.ModuleType(name)
    m.__path__ = []
    sys.modules.setdefault(name, m)
sys.modules["cvxopt"].matrix = None
sys.modules["cvxopt"].solvers = None
from gen_leaf_golden import ref_map, CFG  # noqa: E402
from map import costmap  # noqa: E402
from optimization import path_optimazition  # noqa: E402
from automatedvaletparking_b200 import scenarios as scn  # noqa: E402

warnings.simplefilter("ignore")


def ocp_class():
    try:
        from optimization import ocp_optimization
        return ocp_optimization.ocp_optimization
    except Exception as e:           # pyomo symbols used at import time
        print("ocp_optimization not importable with stubs:", e)
        return None


def main():
    rng = np.random.default_rng(11)
    out = {}
    veh = costmap.Vehicle()
    cases = [1, 2, 3, 4, 5, 6, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18]
    out["cases"] = np.array(cases, dtype=np.int32)
    out["expand_dis"] = np.float64(CFG["expand_dis"])
    OCP = ocp_class()
    for c in cases:
        s = scn.benchmark_case(c)
        m = ref_map(s)
        g = np.load(os.path.join(HERE, "cases", f"Case{c}.npz"))
        opt = path_optimazition.path_opti(park_map=m, vehicle=veh, config=CFG)
        path = g["out_final_path"]
        lens = g["split_lens"]
        off = np.concatenate([[0], np.cumsum(lens)])
        Hs, slacks = [], []
        for i in range(len(lens)):
            seg = [list(map(float, p)) for p in path[off[i]:off[i + 1]]]
            opt.original_path = seg
            H, slack = opt.compute_collision_H()
            Hs.append(np.asarray(H, dtype=np.float64).reshape(-1))
            slacks.append(np.asarray(slack, dtype=np.float64).reshape(-1))
        out[f"c{c}_H"] = np.concatenate(Hs)
        out[f"c{c}_slack"] = np.concatenate(slacks)
        out[f"c{c}_lens"] = lens.astype(np.int32)
        # random poses over the map (theta within [-pi, pi], incl. the quadrant boundaries and axis-aligned headings)
        nrand = 160 if c in (1, 4, 5, 13, 16) else 24
        b = m.boundary
        th = rng.uniform(-np.pi, np.pi, nrand)
        th[:6] = [0.0, np.pi / 2, -np.pi / 2, np.pi, -np.pi, -0.0]
        P = np.stack([rng.uniform(b[0] + 2, b[1] - 2, nrand), rng.uniform(b[2] + 2, b[3] - 2, nrand), th], 1)
        opt.original_path = [list(map(float, p)) for p in P]
        H, _ = opt.compute_collision_H()
        out[f"c{c}_rand_poses"] = P
        out[f"c{c}_rand_H"] = np.asarray(H, dtype=np.float64).reshape(-1)
        if OCP is not None and c in (1, 4, 13):
            o = OCP.__new__(OCP)
            o.map = m; o.vehicle = veh; o.config = CFG; o.expand_dis = CFG["expand_dis"]
            Q = P.copy(); Q[:, 2] += rng.choice([0.0, 2 * np.pi, -2 * np.pi], nrand)     # the OCP variant wraps theta itself
            xs = o.compute_collision_H(path=[list(map(float, p)) for p in Q])
            out[f"c{c}_ocp_poses"] = Q
            out[f"c{c}_ocp"] = np.array([[float(v) for v in col] for col in xs], dtype=np.float64)
        print("Case", c, "segments", len(lens), "points", int(lens.sum()), "rand", nrand, flush=True)
    np.savez_compressed(os.path.join(HERE, "leaf_corridor.npz"), **out)


if __name__ == "__main__":
    main()
