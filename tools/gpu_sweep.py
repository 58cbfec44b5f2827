"""Sweeps of the search launch's scheduling switches on a bench workload (run through gpurun).

    python tools/gpu_sweep.py c2 "" AVP_QUANTUM=128 AVP_QUANTUM=2048 AVP_SPREAD=0 AVP_PLAN_BLOCK=256 "AVP_PLAN_BLOCK=256 AVP_QUANTUM=128"

Every argument after the workload is one configuration (space-separated NAME=VALUE pairs, "" = defaults); the switches are
read by libavp_b200.so at plan time, so one process (one scenario build) serves them all.  Prints per configuration:
search ms (best of 2 after one warm-up), Dijkstra / search kernel split, suspensions, CTA width, successors."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import bench
from automatedvaletparking_b200.batch import DevicePlanner

os.environ.setdefault("AVP_HOST_TIMEOUT_S", "300")
name = sys.argv[1]
configs = sys.argv[2:] or [""]
dp = DevicePlanner(max_pops=20000)
scs, n_total, scaling, gids, _ = bench.make_workload(name, int(os.environ.get("SWEEP_RANK", "0")), 1, dp)
dp.load(scs)
base = None
touched = set()
for c in configs:
    for k in touched:
        os.environ.pop(k, None)
    for kv in c.split():
        k, v = kv.split("=")
        os.environ[k] = v
        touched.add(k)
    best = None
    for it in range(3):
        ms = dp.plan_resident(256, 0)
        a, b, nsus, blk = dp.last_search_passes()
        if it and (best is None or ms < best[0]):
            best = (ms, a, b, nsus, blk, dp.last_narrow_ms())
    res = dp.fetch(256, 0)
    s = res.summaries
    if base is None:
        base = s.copy()
    same = np.array_equal(base, s)
    hist = np.bincount(s["status"], minlength=7)
    print("%-4s %-44s search %8.1f ms (dijkstra %6.1f + plan %8.1f, narrow %6.1f) suspends %6d block %d successors %d M/s %.2f same_as_first %s hist %s" %
          (name, c or "(defaults)", best[0], best[1], best[2], best[5], best[3], best[4], res.successors, res.successors / best[0] / 1e3, same, hist.tolist()), flush=True)
dp.close()
