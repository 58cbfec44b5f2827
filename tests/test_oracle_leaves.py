"""The CPU oracle against per-function vectors generated from the UNMODIFIED reference
(tests/golden/leaf_*.npz, made by tests/golden/gen_leaf_golden.py)."""
import os

import numpy as np

import oracle_lib as O
from automatedvaletparking_b200 import scenarios as scn
from conftest import GOLDEN


def _load(name):
    return np.load(os.path.join(GOLDEN, name))


def test_vehicle_corners_and_angle_wraps(cfg):
    g = _load("leaf_geometry.npz")
    got = np.array([O.vehicle_corners(cfg, *p) for p in g["poses"]])
    assert np.array_equal(got, g["corners"])            # incl. the BLAS fma pattern of the 2x2 dot
    L = O.lib()
    assert np.array_equal(np.array([L.orc_pi_2_pi(float(t)) for t in g["th"]]), g["pi_2_pi"])
    assert cfg.min_radius_turn == float(g["min_radius_turn"])
    m = O.OracleMap(scn.benchmark_case(1))
    assert np.array_equal(np.array([m.index(*p) for p in g["idx_pts"]], dtype=np.int64), g["idx"])


def test_collision_checkers(cfg):
    from automatedvaletparking_b200.hostcfg import make_avp_config
    g = _load("leaf_collision.npz")
    circ = make_avp_config()
    circ.collision_mode = 1
    for case in (1, 5, 13, 19):
        m = O.OracleMap(scn.benchmark_case(case))
        P = g[f"c{case}_poses"]
        got = np.array([m.check(cfg, *p) for p in P])
        assert np.array_equal(got, g[f"c{case}_distance"]), f"distance checker, Case{case}"
        assert 0.1 < got.mean() < 0.95        # the sample exercises both outcomes
        gotc = np.array([m.check(circ, *p) for p in P[:400]])
        assert np.array_equal(gotc, g[f"c{case}_circle"]), f"circle checker, Case{case}"


def test_reeds_shepp_words_selection_and_course():
    g = _load("leaf_rs.npz")
    q, maxc = g["q"], float(g["maxc"])
    w0 = 0
    c0 = 0
    for i in range(len(q)):
        words, deg = O.rs_words(q[i, :3], q[i, 3:], maxc, 1, 1)
        n = int(g["words_n"][i])
        assert not deg and len(words) == n
        for k, (ct, ln, L) in enumerate(words):
            assert ct == str(g["words_ct"][w0 + k])
            assert np.array_equal(ln, g["words_len"][w0 + k][:len(ln)])
            assert L == g["words_L"][w0 + k]               # CPython sum() typing (Neumaier vs naive)
        w0 += n
        if i < 1500:
            r = O.rs_optimal(q[i, :3], q[i, 3:], maxc)
            npt = int(g["npts"][i])
            assert r["rc"] == 0 and r["ctypes"] == str(g["sel_ct"][i])
            assert np.array_equal(r["lengths"], g["sel_len"][i][:len(r["lengths"])]) and r["L"] == g["sel_L"][i]
            assert len(r["x"]) == npt
            assert np.array_equal(r["x"], g["cx"][c0:c0 + npt]) and np.array_equal(r["y"], g["cy"][c0:c0 + npt])
            assert np.array_equal(r["yaw"], g["cyaw"][c0:c0 + npt]) and np.array_equal(r["directions"], g["cdir"][c0:c0 + npt])
        c0 += int(g["npts"][i])
    for i in range(len(g["root_L"])):       # root node: phi is a Python float
        r = O.rs_optimal(q[i, :3], q[i, 3:], maxc, xy_np=1, phi_np=0)
        assert r["ctypes"] == str(g["root_ct"][i]) and r["L"] == g["root_L"][i]


def test_dijkstra_query_sequences():
    g = _load("leaf_dijkstra.npz")
    for case in (1, 4):
        m = O.OracleMap(scn.benchmark_case(case))
        d = O.OracleDijkstra(m)
        for (x, y), (gid, dist, ncl, called) in zip(g[f"c{case}_queries"], g[f"c{case}_res"]):
            assert m.index(x, y) == gid
            if called:
                got, closed = d.compute_path(x, y)
                assert (got, closed) == (dist, ncl)
            else:
                assert d.hvalues()[gid] == dist
        hv = d.hvalues()
        ids = g[f"c{case}_h_ids"]
        assert np.array_equal(hv[ids], g[f"c{case}_h_val"])
        assert (hv >= 0).sum() == len(ids)               # visited-grid cells identical


def test_rasters_of_perturbed_and_synthetic_scenarios():
    g = _load("leaf_maps.npz")
    for i in range(int(g["n"])):
        row = g[f"m{i}_row"]
        s = scn.parse_case_row(row)
        b = g[f"m{i}_boundary"]
        if not np.isnan(b[0]):
            s.boundary = tuple(b)
        m = O.OracleMap(s)
        geom = g[f"m{i}_geom"]
        assert (m.nx, m.ny) == (int(geom[0]), int(geom[1]))
        assert np.array_equal(m.boundary, geom[2:6]) and (m.dx, m.dy) == (geom[6], geom[7])
        ix, iy = np.where(m.cost_map() == 255)
        assert np.array_equal(np.stack([ix, iy], 1).astype(np.uint16), g[f"m{i}_cells"])
        assert m.raster_error == 0


def test_numpy_trig_dispatches_to_libm():
    g = _load("leaf_numpy.npz")
    assert bool(g["cos_eq"]) and bool(g["sin_eq"]) and bool(g["scalar_eq"])
    assert np.array_equal(g["tan_steer"], g["math_tan"])
