"""CPU: the oracle's corridor extraction (oracle/avp_oracle.c orc_corridor) against vectors of the UNMODIFIED
reference's path_opti.compute_collision_H / ocp_optimization.compute_collision_H
(tests/golden/leaf_corridor.npz, made by tests/golden/gen_corridor_golden.py).  Bit-exact."""
import os

import numpy as np
import pytest

import oracle_lib as O
from automatedvaletparking_b200 import scenarios as scn
from conftest import GOLDEN


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLDEN, "leaf_corridor.npz"))


def h_vectors(P, d):
    """H_collision_matrix / slack_H_collision_matrix of one path (path_optimazition.py:651-660)"""
    n = len(P)
    hmax = np.stack([d[:, 0] + P[:, 0], d[:, 1] + P[:, 1]], 1).reshape(-1)
    hmin = np.stack([P[:, 0] - d[:, 2], P[:, 1] - d[:, 3]], 1).reshape(-1)
    H = np.concatenate([hmax, -hmin])
    slack = np.concatenate([hmax, 999 * np.ones(max(n - 2, 0)), -hmin, np.zeros(max(n - 2, 0))])
    return H, slack


def same(a, b):
    return a.shape == b.shape and bool(((a == b) | (np.isnan(a) & np.isnan(b))).all())


def test_random_poses_all_cases(gold, cfg, native_built):
    e = float(gold["expand_dis"])
    for c in gold["cases"]:
        m = O.OracleMap(scn.benchmark_case(int(c)))
        P = gold[f"c{c}_rand_poses"]
        d, st = m.corridor(cfg, P, e)
        assert not st.any()
        H, _ = h_vectors(P, d)
        assert same(H, gold[f"c{c}_rand_H"]), f"Case{c}"
        assert (d <= e).all() and (d[~np.isnan(d)] >= 0).all()


def test_planned_path_segments(gold, cfg, native_built):
    """the gear segments of the reference's own path_planning() output, as main.py feeds them to the QP (main.py:68-75)"""
    e = float(gold["expand_dis"])
    for c in gold["cases"]:
        m = O.OracleMap(scn.benchmark_case(int(c)))
        path = np.load(os.path.join(GOLDEN, "cases", f"Case{c}.npz"))["out_final_path"]
        lens = gold[f"c{c}_lens"]
        off = np.concatenate([[0], np.cumsum(lens)])
        Hs, Ss = [], []
        for i in range(len(lens)):
            P = path[off[i]:off[i + 1]]
            d, st = m.corridor(cfg, P, e)
            assert not st.any()
            H, S = h_vectors(P, d)
            Hs.append(H); Ss.append(S)
        assert same(np.concatenate(Hs), gold[f"c{c}_H"]), f"Case{c}"
        assert same(np.concatenate(Ss), gold[f"c{c}_slack"]), f"Case{c}"


def test_ocp_variant_wraps_heading(gold, cfg, native_built):
    e = float(gold["expand_dis"])
    for c in (1, 4, 13):
        m = O.OracleMap(scn.benchmark_case(c))
        Q = gold[f"c{c}_ocp_poses"]
        W = Q.copy()
        W[:, 2] = [O.lib().orc_pi_2_pi(float(t)) for t in Q[:, 2]]           # ocp_optimization.py:130
        d, st = m.corridor(cfg, W, e)
        assert not st.any()
        ref = gold[f"c{c}_ocp"]                                              # X_max, Y_max, X_min, Y_min
        got = np.stack([d[:, 0] + Q[:, 0], d[:, 1] + Q[:, 1], Q[:, 0] - d[:, 2], Q[:, 1] - d[:, 3]])
        assert same(got, ref), f"Case{c}"


def test_heading_outside_range_is_flagged(cfg, native_built):
    m = O.OracleMap(scn.benchmark_case(1))
    d, st = m.corridor(cfg, [[-16.0, -13.5, 3.5], [-16.0, -13.5, 0.2], [-16.0, -13.5, -4.0]], 0.8)
    assert list(st) == [1, 0, 1] and np.isnan(d[0]).all() and not np.isnan(d[1]).any()
