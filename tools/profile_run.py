"""Small fixed workload for ncu captures: BASELINE configs[1] recipe, fewer scenarios and a lower pop cap
so that one replayed launch of the search kernel stays short."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import bench
from automatedvaletparking_b200.batch import DevicePlanner
n = int(sys.argv[1]) if len(sys.argv) > 1 else 384
max_pops = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
os.environ.setdefault('AVP_HOST_TIMEOUT_S', '900')
dp = DevicePlanner(max_pops=max_pops)
scs = bench.make_scenarios(0, n, dp)
dp.load(scs)
ms = dp.plan_resident(256, 0)
res = dp.fetch(256, 0)
print('search ms', ms, 'passes', dp.last_search_passes(), 'successors', res.successors, 'launches', dp.launches)
