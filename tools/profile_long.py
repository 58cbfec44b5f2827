"""ncu workload of the LONG searches only: the bench workload (C2) is planned once with a reduced pop cap, the scenarios that hit
the cap are re-loaded alone (one per SM pair, no short searches, few Dijkstra resumes) and planned again -- the launch to
capture is the third k_plan launch of the process (`-k k_plan --launch-skip 2 --launch-count 1`: the first plan is a narrow and a
wide launch, the second one a single launch).
    python tools/profile_long.py [max_pops]"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
import bench
from automatedvaletparking_b200.batch import DevicePlanner
max_pops = int(sys.argv[1]) if len(sys.argv) > 1 else 6000
os.environ.setdefault('AVP_HOST_TIMEOUT_S', '900')
dp = DevicePlanner(max_pops=max_pops)
scs = bench.make_scenarios(0, 1024, dp)
dp.load(scs)
dp.plan_resident(256, 0)
res = dp.fetch(256, 0)
idx = np.where(res.summaries["n_pops"] >= max_pops)[0]
print('capped scenarios', len(idx), 'launches so far', dp.launches)
dp.load([scs[i] for i in idx])
ms = dp.plan_resident(256, 0)
res = dp.fetch(256, 0)
print('search ms', ms, 'passes', dp.last_search_passes(), 'successors', res.successors, 'pops', int(res.summaries["n_pops"].sum()))
