"""GPU tests of the drop-in modules (reference module paths / class names / signatures)
against traces of the UNMODIFIED reference (tests/golden/cases)."""
import os
import sys

import numpy as np
import pytest

from automatedvaletparking_b200 import scenarios as scn
from conftest import GOLDEN, ROOT

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dropin(native_built):
    """import the drop-ins under the reference's own module names (map, collision_check, path_plan, config) with
    INTEGRATION.md's sys.path setup: the shim directory first"""
    shim = os.path.join(ROOT, "automatedvaletparking_b200", "dropin")
    sys.path.insert(0, shim)
    try:
        for name in [n for n in sys.modules if n.split(".")[0] in ("map", "collision_check", "path_plan", "config")]:
            del sys.modules[name]
        from map import costmap
        from path_plan import path_planner, hybrid_a_star, rs_curve, compute_h
        from collision_check import collision_check
        from config import read_config
        assert costmap.__file__.startswith(shim) and path_planner.__file__.startswith(shim)
        yield dict(costmap=costmap, path_planner=path_planner, hybrid_a_star=hybrid_a_star, rs_curve=rs_curve, compute_h=compute_h,
                   collision_check=collision_check, config=read_config.read_config("config"))
    finally:
        sys.path.remove(shim)
        for name in [n for n in sys.modules if n.split(".")[0] in ("map", "collision_check", "path_plan", "config")]:
            del sys.modules[name]


def _case_csv(tmp_path, n):
    p = os.path.join(str(tmp_path), f"Case{n}.csv")
    scn.write_case_csv(scn.benchmark_case(n), p)
    return p


@pytest.mark.parametrize("case", [1, 4, 17])
def test_path_planning_matches_reference(dropin, tmp_path, case):
    g = np.load(os.path.join(GOLDEN, "cases", f"Case{case}.npz"))
    cfg = dropin["config"]
    m = dropin["costmap"].Map(file=_case_csv(tmp_path, case), discrete_size=cfg["map_discrete_size"])
    assert m.cost_map.shape == tuple(g["dims"]) and m.cost_map.dtype == np.float64
    ix, iy = np.where(m.cost_map == 255)
    assert np.array_equal(np.stack([ix, iy], 1).astype(np.uint16), g["obs_cells"])
    assert (m._discrete_x, m._discrete_y) == tuple(g["pitch"]) and np.array_equal(m.boundary, g["boundary"])
    v = dropin["costmap"].Vehicle()
    planner = dropin["path_planner"].PathPlanner(config=cfg, map=m, vehicle=v)
    out_final_path, info, split_list = planner.path_planning()
    assert np.array_equal(planner.pop_indices, g["pops"])                      # expanded-node indices
    assert np.array_equal(np.array(info["astar_path"], dtype=np.float64).reshape(-1, 3), g["astar_path"])
    rs = info["rs_path"]
    assert "".join(rs.ctypes) == str(g["rs_ctypes"]) and rs.L == float(g["rs_L"])
    assert np.array_equal(np.array(rs.lengths), g["rs_lengths"])
    assert np.array_equal(np.array(rs.x), g["rs_x"]) and np.array_equal(np.array(rs.yaw), g["rs_yaw"])
    assert np.array_equal(np.array(rs.directions), g["rs_dir"])
    assert np.array_equal(np.array(out_final_path, dtype=np.float64).reshape(-1, 3), g["out_final_path"])   # incl. split_path extension points
    assert [len(s) for s in split_list] == list(g["split_lens"]) and info["change_gear"] == int(g["change_gear"])


def test_case20_raises_like_the_reference(dropin, tmp_path):
    cfg = dropin["config"]
    m = dropin["costmap"].Map(file=_case_csv(tmp_path, 20), discrete_size=cfg["map_discrete_size"])
    planner = dropin["path_planner"].PathPlanner(config=cfg, map=m, vehicle=dropin["costmap"].Vehicle())
    with pytest.raises(AttributeError):                # path_planner.py:104 on an exhausted open list
        planner.path_planning()
    assert list(planner.pop_indices) == [0]


def test_stepwise_api_reproduces_reference_pops(dropin, tmp_path):
    """the reference's own search loop (path_planner.py:68-98) driven through the drop-in hybrid_a_star"""
    g = np.load(os.path.join(GOLDEN, "cases", "Case16.npz"))
    cfg = dropin["config"]
    m = dropin["costmap"].Map(file=_case_csv(tmp_path, 16), discrete_size=cfg["map_discrete_size"])
    astar = dropin["hybrid_a_star"].hybrid_a_star(config=cfg, park_map=m, vehicle=dropin["costmap"].Vehicle())
    assert len(astar.h_value_list) == int(g["H0"])
    pops, fgh = [], []
    reach = False
    while not astar.open_list.empty() and not reach:
        cur = astar.open_list.get()
        pops.append(cur.index)
        fgh.append((cur.f, cur.g, cur.h))
        rs_path, collision, info = astar.try_reach_goal(cur)
        if not collision and info["in_radius"]:
            break
        astar.expand_node(cur)
    assert pops == list(g["pops"])
    assert np.array_equal(np.array(fgh, dtype=np.float64), g["pop_fgh"])
    assert astar.global_index == int(g["global_index"])
    assert (len(astar.closed_list), len(astar.open_list.queue)) == (int(g["n_closed"]), int(g["n_open"]))
    assert len(astar.heuristic.closedlist) == int(g["H_end"])
    path = astar.finish_path(cur)
    assert np.array_equal(np.array(path, dtype=np.float64), g["astar_path"])


def test_leaf_dropins(dropin, tmp_path):
    cfg = dropin["config"]
    g = np.load(os.path.join(GOLDEN, "leaf_collision.npz"))
    m = dropin["costmap"].Map(file=_case_csv(tmp_path, 5), discrete_size=cfg["map_discrete_size"])
    v = dropin["costmap"].Vehicle()
    dc = dropin["collision_check"].distance_checker(map=m, vehicle=v, config=cfg)
    P = g["c5_poses"][:64]
    assert [dc.check(node_x=p[0], node_y=p[1], theta=p[2]) for p in P] == list(g["c5_distance"][:64])
    cc = dropin["collision_check"].two_circle_checker(map=m, vehicle=v, config=cfg)
    assert [cc.check(node_x=p[0], node_y=p[1], theta=p[2]) for p in P] == list(g["c5_circle"][:64])
    gg = np.load(os.path.join(GOLDEN, "leaf_geometry.npz"))
    got = np.array([v.create_anticlockpoint(x=np.float64(p[0]), y=np.float64(p[1]), theta=np.float64(p[2]), config=cfg).reshape(5, 2) for p in gg["poses"][:200]])
    assert got.shape == (200, 5, 2) and np.array_equal(got, gg["corners"][:200])
    r = np.load(os.path.join(GOLDEN, "leaf_rs.npz"))
    rs = dropin["rs_curve"]
    for i in range(20):
        q = r["q"][i]
        p = rs.calc_optimal_path(np.float64(q[0]), np.float64(q[1]), np.float64(q[2]), float(q[3]), float(q[4]), float(q[5]), np.float64(r["maxc"]))
        assert "".join(p.ctypes) == str(r["sel_ct"][i]) and p.L == r["sel_L"][i] and len(p.x) == int(r["npts"][i])
    d = np.load(os.path.join(GOLDEN, "leaf_dijkstra.npz"))
    m1 = dropin["costmap"].Map(file=_case_csv(tmp_path, 1), discrete_size=cfg["map_discrete_size"])
    dj = dropin["compute_h"].Dijkstra(m1)
    for (x, y), (gid, dist, ncl, called) in zip(d["c1_queries"], d["c1_res"]):
        assert m1.convert_position_to_index(x, y) == gid
        if called:
            got, cl = dj.compute_path(x, y)
            assert (got, len(cl)) == (dist, ncl)
        else:
            assert dj.closedlist.lookup(int(gid)) == dist
