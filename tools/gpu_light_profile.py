"""Where do the two sides of the pipelined pop wait?  Needs the light profiling build:
    tools/build_variant.sh lp -DAVP_PROFILE_LIGHT;  AVP_B200_LIB=$PWD/automatedvaletparking_b200/libavp_b200_lp.so python tools/gpu_light_profile.py [workload]
Four shared-memory accumulators per scenario (avp_plan.cuh), cycles per pop."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import bench
from automatedvaletparking_b200.batch import DevicePlanner
os.environ.setdefault("AVP_HOST_TIMEOUT_S", "300")
name = sys.argv[1] if len(sys.argv) > 1 else "c2"
dp = DevicePlanner(max_pops=20000)
scs, n_total, scaling, gids, _ = bench.make_workload(name, 0, 1, dp)
dp.load(scs)
ms = dp.plan_resident(256, 0); ms = dp.plan_resident(256, 0)
res = dp.fetch(256, 0)
pr = dp.phase_profile().astype(np.float64)
s = res.summaries
print("search ms %.1f" % ms, dp.last_search_passes(), "successors", res.successors)
for nm, idx in (("long (20000 pops)", np.where(s["n_pops"] >= 20000)[0]), ("mid (1024..20000)", np.where((s["n_pops"] >= 1024) & (s["n_pops"] < 20000))[0]),
                ("short (< 1024)", np.where(s["n_pops"] < 1024)[0])):
    if len(idx) == 0:
        continue
    pops = np.maximum(pr[idx, 3], 1.0)
    print("%-18s n %4d pops/scenario %8.1f | per pop: evaluators wait at A %7.0f | serial section A->B %7.0f | commit warp waits at A %7.0f"
          " | commit warp: records+inserts %6.0f  commit loop %6.0f  Dijkstra %6.0f  heappop %6.0f" %
          (nm, len(idx), s["n_pops"][idx].mean(), (pr[idx, 0] / pops).mean(), (pr[idx, 1] / pops).mean(), (pr[idx, 2] / pops).mean(),
           (pr[idx, 4] / pops).mean(), (pr[idx, 5] / pops).mean(), (pr[idx, 6] / pops).mean(), (pr[idx, 7] / pops).mean()))
    print("%-18s pushes per pop %.2f, cycles per push %.0f (inside the commit loop)" % ("", (pr[idx, 9] / pops).mean(), pr[idx, 8].sum() / max(pr[idx, 9].sum(), 1.0)))
    tot = float(s["n_pops"][idx].sum())
    print("%-18s total pops %.0f" % ("", tot))
dp.close()

# wall-clock view: when was every scenario first taken, when did it finish, how long did it hold an SM (light build, prof[10..13])
t0 = pr[:, 11][pr[:, 11] > 0].min()
take = (pr[:, 11] - t0) / 1e6; fin = (pr[:, 12] - t0) / 1e6; on_ms = pr[:, 10] / 1.965e6; segs = pr[:, 13]
long_ = np.where(s["n_pops"] >= 20000)[0]; mid = np.where((s["n_pops"] >= 1024) & (s["n_pops"] < 20000))[0]
print("kernel span %.1f ms (first take-over to last finish)" % fin.max())
for nm, idx in (("long", long_), ("mid", mid)):
    if len(idx) == 0: continue
    print("%-5s n %3d | first taken at %.1f..%.1f ms | finished at %.1f..%.1f ms (mean %.1f) | held an SM %.1f..%.1f ms (mean %.1f) | waited (finish - take - on SM) mean %.1f max %.1f ms | run segments mean %.1f" %
          (nm, len(idx), take[idx].min(), take[idx].max(), fin[idx].min(), fin[idx].max(), fin[idx].mean(), on_ms[idx].min(), on_ms[idx].max(), on_ms[idx].mean(),
           (fin[idx] - take[idx] - on_ms[idx]).mean(), (fin[idx] - take[idx] - on_ms[idx]).max(), segs[idx].mean()))
order = np.argsort(-fin)[:8]
print("last finishers: " + "  ".join("#%d pops %d fin %.0f onSM %.0f take %.0f" % (i, s["n_pops"][i], fin[i], on_ms[i], take[i]) for i in order))
# what becomes of the 10 successors of a pop (long searches): pushed, closed on creation (collision), found again in the open list
# (key update or not), skipped (found in the closed list / out of bounds)
li = np.where(s["n_pops"] >= 20000)[0]
if len(li):
    G = s["global_index"][li].astype(np.float64); pops_ = s["n_pops"][li].astype(np.float64)
    pushed = s["n_open"][li] + pops_; coll = s["n_closed"][li] - pops_; upd = s["n_hcalls"][li] - pushed; skip = G - pushed - coll - upd
    print("successors per pop (long): pushed %.2f  collided %.2f  found open %.2f  skipped (closed / out of bounds) %.2f" %
          ((pushed / pops_).mean(), (coll / pops_).mean(), (upd / pops_).mean(), (skip / pops_).mean()))
