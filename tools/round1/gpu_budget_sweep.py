"""Pass-1 pop budget sweep on the bench workload (AVP_POP_BUDGET is read at every plan call)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import bench
from automatedvaletparking_b200.batch import DevicePlanner
os.environ.setdefault('AVP_HOST_TIMEOUT_S', '120')
dp = DevicePlanner(max_pops=20000)
scs = bench.make_scenarios(0, 1024, dp)
dp.load(scs)
for b in [int(x) for x in (sys.argv[1:] or ['1024', '512', '384', '256', '192', '128', '64'])]:
    os.environ['AVP_POP_BUDGET'] = str(b)
    dp.plan_resident(256, 0)
    ms = dp.plan_resident(256, 0)
    print('budget %5d: search ms %8.1f passes %s' % (b, ms, dp.last_search_passes()), flush=True)
