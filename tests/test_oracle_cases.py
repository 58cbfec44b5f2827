"""The CPU oracle (oracle/avp_oracle.c) against full planner traces of the UNMODIFIED
reference (tests/golden/cases/*.npz, made by tests/golden/gen_case_golden.py).  Bit-exact:
raster, pop sequence, popped states, f/g/h, Dijkstra query log, path, rs word."""
import glob
import os

import numpy as np
import pytest

import oracle_lib as O
from automatedvaletparking_b200 import scenarios as scn
from conftest import GOLDEN

CASES = sorted(glob.glob(os.path.join(GOLDEN, "cases", "Case*.npz")), key=lambda p: int(os.path.basename(p)[4:-4]))
assert CASES, "no golden case files"


@pytest.mark.parametrize("path", CASES, ids=[os.path.basename(p)[:-4] for p in CASES])
def test_oracle_reproduces_reference_trace(path, cfg):
    g = np.load(path)
    sc = scn.benchmark_case(int(os.path.basename(path)[4:-4]))
    m = O.OracleMap(sc)
    # Map (costmap.py:159-261)
    assert (m.nx, m.ny) == tuple(g["dims"])
    assert np.array_equal(m.boundary, g["boundary"])
    assert (m.dx, m.dy) == (g["pitch"][0], g["pitch"][1])
    ix, iy = np.where(m.cost_map() == 255)
    assert np.array_equal(np.stack([ix, iy], 1).astype(np.uint16), g["obs_cells"])
    # search (path_planner.py:58-110)
    r = O.plan(m, cfg)
    status = str(g["status"])
    assert r["status"] == (0 if status == "ok" else 1)
    assert np.array_equal(r["pops"], g["pops"])                       # expanded-node indices
    assert np.array_equal(r["pop_state"], g["pop_state"])             # bit-exact poses
    assert np.array_equal(r["pop_fgh"], g["pop_fgh"])                 # bit-exact f, g, h
    assert np.array_equal(r["hq"], g["hq"][:len(r["hq"])])            # (target id, distance, |closedlist|) per compute_path
    assert r["n_hq"] == len(g["hq"])
    assert r["global_index"] == int(g["global_index"])
    assert (r["n_closed"], r["n_open"]) == (int(g["n_closed"]), int(g["n_open"]))
    assert r["h_closed"] == int(g["H_end"])
    assert r["n_hcalls"] == int(g["n_hcalls"])
    if status == "ok":
        final = np.concatenate([g["astar_path"], np.stack([g["rs_x"], g["rs_y"], g["rs_yaw"]], 1)[1:]], 0)
        assert np.array_equal(r["final_path"], final)
        assert r["n_astar"] == len(g["astar_path"])
        assert r["rs_ctypes"] == str(g["rs_ctypes"])
        assert np.array_equal(r["rs_lengths"], g["rs_lengths"])
        assert r["rs_L"] == float(g["rs_L"])
        assert np.array_equal(r["rs_dir"], g["rs_dir"])


def test_all_benchmark_cases_have_or_await_goldens():
    have = {int(os.path.basename(p)[4:-4]) for p in CASES}
    # Cases 7, 8, 19 need hours in the reference; everything else must be pinned
    assert have >= set(range(1, 21)) - {7, 8, 19}
