"""GPU: PathPlanner.split_path on the device (avp_split_path / avp_split_paths) against vectors of the
UNMODIFIED reference and against the oracle on whole batches."""
import os

import numpy as np
import pytest

import oracle_lib as O
from automatedvaletparking_b200 import scenarios as scn
from conftest import GOLDEN

pytestmark = pytest.mark.gpu


def test_split_one_path_matches_reference_vectors(device_planner):
    dp = device_planner
    g = np.load(os.path.join(GOLDEN, "leaf_split.npz"))
    cases = (1, 4, 13)
    dp.load([scn.benchmark_case(c) for c in cases])
    for k in range(int(g["n"])):
        s = cases.index(int(g[f"p{k}_case"]))
        st, segs, cg = dp.split_path(s, g[f"p{k}_in"])
        assert st == int(g[f"p{k}_status"]), k
        if st == 0:
            assert [len(x) for x in segs] == list(g[f"p{k}_lens"]) and cg == int(g[f"p{k}_cg"]), k
            assert np.array_equal(np.concatenate(segs, 0), g[f"p{k}_out"]), k


def test_split_batch_of_the_benchmark_cases_matches_reference(device_planner):
    dp = device_planner
    scs = [scn.benchmark_case(i) for i in range(1, 21)]
    dp.load(scs)
    res = dp.plan(cap_path=512, cap_pops=0)
    sp = dp.split_paths(cap_pts=640, cap_seg=16)
    n = 0
    for k, sc in enumerate(scs):
        g = np.load(os.path.join(GOLDEN, "cases", f"{sc.name}.npz"))
        if str(g["status"]) != "ok":
            assert int(sp["status"][k]) == 2            # no path: Cases 7, 8, 19 (capacity), 20 (open list exhausted)
            continue
        assert int(sp["status"][k]) == 0
        ns = int(sp["n_seg"][k])
        assert list(sp["seg_len"][k][:ns]) == list(g["split_lens"]) and int(sp["change_gear"][k]) == int(g["change_gear"])
        assert np.array_equal(sp["pts"][k][:int(sp["n_pts"][k])], g["out_final_path"]), sc.name
        n += 1
    assert n >= 16


def test_split_full_bench_batch_vs_oracle(device_planner, cfg):
    """BASELINE configs[1] at full size: split_path of every finished plan of the 1024-scenario batch, against the oracle's
    split_path of the same final paths (which the parity tests pin to the oracle's own plans)."""
    import bench
    dp = device_planner
    scs = bench.make_scenarios(0, 1024, dp)
    dp.load(scs)
    res = dp.plan(cap_path=512, cap_pops=0)
    sp = dp.split_paths(cap_pts=640, cap_seg=24)
    s = res.summaries
    n_ok = n_idx = 0
    for k, sc in enumerate(scs):
        if int(s["status"][k]) != 0:
            assert int(sp["status"][k]) == 2
            continue
        r = O.OracleMap(sc).split_path(cfg, res.path(k), cap_pts=640, cap_seg=24)
        assert int(sp["status"][k]) == r["status"], k
        if r["status"] == 0:
            ns = int(sp["n_seg"][k])
            assert list(sp["seg_len"][k][:ns]) == list(r["seg_len"]) and int(sp["change_gear"][k]) == r["change_gear"], k
            assert np.array_equal(sp["pts"][k][:int(sp["n_pts"][k])], r["out_final_path"]), k
            n_ok += 1
        else:
            n_idx += 1
    assert n_ok > 800
