import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import numpy as np
import bench
from automatedvaletparking_b200.batch import DevicePlanner
os.environ.setdefault('AVP_HOST_TIMEOUT_S', '120')
dp = DevicePlanner(max_pops=20000)
scs = bench.make_scenarios(0, 1024, dp)
dp.load(scs)
ms = dp.plan_resident(256, 0); ms = dp.plan_resident(256, 0)
res = dp.fetch(256, 0)
pr = dp.phase_profile()
s = res.summaries
print('search ms', ms, 'passes', dp.last_search_passes())
names = ['init+dij0', 'looptop', 'pop+poses', 'lookup+coll+rs', 'select+plan', 'course+shot', 'commit', 'commit_prep']
long_ = np.where(s['n_pops'] >= 20000)[0]
short = np.where(s['n_pops'] < 192)[0]
for nm, idx in (('long(20000 pops)', long_), ('short', short)):
    if len(idx) == 0: continue
    p = pr[idx].astype(np.float64)
    pops = s['n_pops'][idx].astype(np.float64)
    print(nm, 'n', len(idx), 'mean pops', pops.mean(), 'total cycles/scenario %.3g' % p.sum(1).mean())
    for k, n in enumerate(names):
        print('   %-16s %10.0f cycles/pop  (%.1f%%)' % (n, (p[:, k] / pops).mean(), 100 * p[:, k].sum() / p.sum()))
if len(long_):
    p = pr[long_].astype(np.float64); pops = s['n_pops'][long_].astype(np.float64)
    print('long: push cycles/pop %.0f  pushes/pop %.2f  cycles/push %.0f  mean final heap pos %.0f  mean heap size %.0f' % ((p[:,8]/pops).mean(), (p[:,9]/pops).mean(), p[:,8].sum()/p[:,9].sum(), p[:,12].sum()/p[:,9].sum(), p[:,14].sum()/pops.sum()))
    print('long: dijkstra-in-commit cycles/pop %.0f  resumes/scenario %.1f  heappop cycles/pop %.0f' % ((p[:,10]/pops).mean(), p[:,11].mean(), (p[:,13]/pops).mean()))
print('h_closed mean', s['h_closed'].mean(), 'n_hq mean', s['n_hq'].mean())
