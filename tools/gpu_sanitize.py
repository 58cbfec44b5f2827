"""A small batch through the whole hot path for compute-sanitizer runs (memcheck / synccheck / racecheck):
    AVP_QUANTUM=8 compute-sanitizer --tool racecheck python tools/gpu_sanitize.py
Rasterisation, eager Dijkstra kernel, the persistent search kernel with suspensions (quantum 8: every search is let go of and
resumed several times), split_path; results checked against the oracle."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oracle_lib as O
from automatedvaletparking_b200 import scenarios as scn
from automatedvaletparking_b200.batch import DevicePlanner

n_pert = int(sys.argv[1]) if len(sys.argv) > 1 else 3
scs = [scn.benchmark_case(c) for c in (1, 4, 16)] + scn.perturbed_set(scn.benchmark_case(1), n_pert, seed=7)
with DevicePlanner(max_pops=400) as dp:
    dp.load(scs)
    res = dp.plan(cap_path=256, cap_pops=400)
    sp = dp.split_paths(cap_pts=400, cap_seg=16)
    for k, sc in enumerate(scs):
        r = O.plan(O.OracleMap(sc), dp.cfg)
        assert int(res.summaries["status"][k]) == r["status"] and np.array_equal(res.pop_indices(k), r["pops"]), sc.name
    a, b, nsus, blk = dp.last_search_passes()
    print("sanitize batch ok:", len(scs), "scenarios, pops", res.summaries["n_pops"].tolist(), "suspensions", nsus, "block", blk, "launches", dp.launches)
