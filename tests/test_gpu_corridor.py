"""GPU: the corridor extraction kernel (k_corridor through avp_corridor) against the oracle and against vectors of
the UNMODIFIED reference (tests/golden/leaf_corridor.npz): path_opti.compute_collision_H
(optimization/path_optimazition.py:221-658) and ocp_optimization.compute_collision_H (ocp_optimization.py:36-480).
Bit-exact fp64."""
import os

import numpy as np
import pytest

import oracle_lib as O
from automatedvaletparking_b200 import scenarios as scn
from conftest import GOLDEN
from test_oracle_corridor import h_vectors, same

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLDEN, "leaf_corridor.npz"))


def test_kernel_matches_reference_vectors_and_oracle(device_planner, gold, cfg):
    dp = device_planner
    cases = [int(c) for c in gold["cases"]]
    dp.load([scn.benchmark_case(c) for c in cases])
    e = float(gold["expand_dis"])
    for k, c in enumerate(cases):
        P = gold[f"c{c}_rand_poses"]
        d, st = dp.corridor(k, P, e)
        assert not st.any()
        H, _ = h_vectors(P, d)
        assert same(H, gold[f"c{c}_rand_H"]), f"Case{c}"
        od, _ = O.OracleMap(scn.benchmark_case(c)).corridor(cfg, P, e)
        assert same(d, od), f"Case{c}"


def test_kernel_large_random_batch_vs_oracle(device_planner, cfg):
    """20 000 poses on the densest map (Case19, 5 696 obstacle cells) and on a 5e9-offset map (Case13), other expand_dis"""
    dp = device_planner
    dp.load([scn.benchmark_case(19), scn.benchmark_case(13)])
    rng = np.random.default_rng(21)
    for k, c in enumerate((19, 13)):
        m = O.OracleMap(scn.benchmark_case(c))
        b = m.boundary
        n = 20000 if c == 19 else 4000
        P = np.stack([rng.uniform(b[0], b[1], n), rng.uniform(b[2], b[3], n), rng.uniform(-np.pi, np.pi, n)], 1)
        P[:8, 2] = [0.0, np.pi / 2, -np.pi / 2, np.pi, -np.pi, 4.0, -3.5, -0.0]
        for e in (0.8, 1.7):
            d, st = dp.corridor(k, P, e)
            od, ost = m.corridor(cfg, P, e)
            assert np.array_equal(st, ost) and st.sum() == 2
            assert same(d, od), (c, e)
            assert (d[st == 0] <= e).all()
    d0, s0 = dp.corridor(0, np.zeros((0, 3)), 0.8)                               # empty input
    assert d0.shape == (0, 4) and s0.shape == (0,)


def test_dropin_classes_match_reference(native_built, gold, tmp_path):
    """the reference's class / method names: path_opti.compute_collision_H(), ocp_optimization.compute_collision_H(path)"""
    from automatedvaletparking_b200.map import costmap
    from automatedvaletparking_b200.optimization.path_optimazition import path_opti
    from automatedvaletparking_b200.optimization.ocp_optimization import ocp_optimization
    from automatedvaletparking_b200.config import read_config
    config = read_config.read_config("config")
    assert float(config["expand_dis"]) == float(gold["expand_dis"])
    for c in (1, 13):
        p = os.path.join(str(tmp_path), f"Case{c}.csv")
        scn.write_case_csv(scn.benchmark_case(c), p)
        m = costmap.Map(file=p, discrete_size=config["map_discrete_size"])
        v = costmap.Vehicle()
        opt = path_opti(park_map=m, vehicle=v, config=config)
        path = np.load(os.path.join(GOLDEN, "cases", f"Case{c}.npz"))["out_final_path"]
        lens = gold[f"c{c}_lens"]
        off = np.concatenate([[0], np.cumsum(lens)])
        Hs, Ss = [], []
        for i in range(len(lens)):
            opt.original_path = [list(map(float, q)) for q in path[off[i]:off[i + 1]]]
            H, S = opt.compute_collision_H()
            assert H.shape == (4 * lens[i], 1) and S.shape == (4 * lens[i] + 2 * (lens[i] - 2), 1)
            Hs.append(H.reshape(-1)); Ss.append(S.reshape(-1))
        assert same(np.concatenate(Hs), gold[f"c{c}_H"]) and same(np.concatenate(Ss), gold[f"c{c}_slack"])
        ocp = ocp_optimization(park_map=m, vehicle=v, config=config)
        Q = gold[f"c{c}_ocp_poses"]
        X = ocp.compute_collision_H(path=[list(map(float, q)) for q in Q])
        assert same(np.array(X, dtype=np.float64), gold[f"c{c}_ocp"])
        with pytest.raises(UnboundLocalError):
            opt.original_path = [[float(Q[0, 0]), float(Q[0, 1]), 3.5]]
            opt.compute_collision_H()
