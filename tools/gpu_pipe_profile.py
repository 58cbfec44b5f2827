"""Per-phase cycle counters of the search kernel (k_plan) on the bench workload.  Needs the profiling build:
    tools/build_variant.sh prof -DAVP_PROFILE;  AVP_B200_LIB=$PWD/automatedvaletparking_b200/libavp_b200_prof.so python tools/gpu_pipe_profile.py
Commit warp (thread 0) and evaluators (thread 32) keep separate accumulators (avp_plan.cuh); the counters of all run segments
(quanta) of a scenario are summed."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import numpy as np
import bench
from automatedvaletparking_b200.batch import DevicePlanner
os.environ.setdefault('AVP_HOST_TIMEOUT_S', '120')
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
dp = DevicePlanner(max_pops=20000)
scs = bench.make_scenarios(0, n, dp)
dp.load(scs)
ms = dp.plan_resident(256, 0); ms = dp.plan_resident(256, 0)
res = dp.fetch(256, 0)
pr = dp.phase_profile()
s = res.summaries
print('search ms', ms, 'passes', dp.last_search_passes(), 'successors', res.successors)
long_ = np.where(s['n_pops'] >= 20000)[0]
mid = np.where((s['n_pops'] >= 1024) & (s['n_pops'] < 20000))[0]
names = {0: 'C init+dij0', 1: 'C accept+lookups+predict', 2: 'C records+commit', 3: 'C heappop', 4: 'C queue help + wait for E',
         7: 'E0 poses', 12: 'E1 course|rs|substep checks', 13: 'E2 course checks|select+combine', 15: 'E wait for C'}
for nm, idx in (('long(20000 pops)', long_), ('mid(1024..20000)', mid)):
    if len(idx) == 0:
        continue
    p = pr[idx].astype(np.float64)
    pops = s['n_pops'][idx].astype(np.float64)
    print(nm, 'n', len(idx), 'mean pops', pops.mean())
    for k in sorted(names):
        print('   %-28s %10.0f cycles/pop' % (names[k], (p[:, k] / pops).mean()))
    print('   hits/pop %.4f  misses/pop %.4f' % ((p[:, 5] / pops).mean(), (p[:, 6] / pops).mean()))
    print('   push cycles/pop %.0f pushes/pop %.2f | dijkstra cycles/pop %.0f resumes/scenario %.1f' %
          ((p[:, 8] / pops).mean(), (p[:, 9] / pops).mean(), (p[:, 10] / pops).mean(), p[:, 11].mean()))
    print('   C total/pop %.0f   E total/pop %.0f' % ((p[:, 1:5].sum(1) / pops).mean(), (p[:, [7, 12, 13, 15]].sum(1) / pops).mean()))

wp = dp.warp_profile()
if len(long_):
    w = wp[long_][:, :, :8].astype(np.float64) / s['n_pops'][long_].astype(np.float64)[:, None, None]
    w = w.mean(0)
    print('per-warp WORK cycles/pop (long scenarios; evaluators: E0 E1 E2 -; warp 0 = commit warp: lookups, commit, heappop, queue help)')
    for k in range(16):
        print('  warp %2d  %7.0f %7.0f %7.0f %7.0f' % (k, w[k, 0], w[k, 1], w[k, 2], w[k, 3]))

if os.environ.get('AVP_TRACE_POP') and len(long_):
    ev = ['afterA', 'arrB', 'afterB', 'arr1|Ccommit', 'aft1|Cpop', 'arr2', 'aft2', 'arrA']
    for sc_i in long_[:3]:
        t = wp[sc_i][:, 8:16].astype(np.int64)
        t0 = t[t > 0].min() if (t > 0).any() else 0
        print('timeline of pop %s, scenario %d (cycles since the first stamp): %s' % (os.environ['AVP_TRACE_POP'], sc_i, ' '.join(ev)))
        for k in range(16):
            print('  warp %2d ' % k + ' '.join('%7d' % (x - t0 if x > 0 else -1) for x in t[k]))

if len(long_):
    ic = wp[long_][:, :6, 16:24].reshape(len(long_), 48).astype(np.float64)
    pops = s['n_pops'][long_].astype(np.float64)
    m = (ic / pops[:, None]).mean(0)
    print('queue items, cycles/pop (long scenarios): course %.0f  rs queries %.0f  sub-step poses %.0f  lookahead %.0f' % (m[0], m[1], m[2], m[3]))
    print('  rs items   ' + ' '.join('%.0f' % x for x in m[4:23]))
    print('  chains     ' + ' '.join('%.0f' % x for x in m[23:33]))
    print('  selections (2 successors each) %.0f per pop, %.2f per pop, %.0f each | course checks %.0f per pop, %.2f per pop, %.0f each' %
          (m[40], m[42], m[40] / max(m[42], 1e-9), m[41], m[43], m[41] / max(m[43], 1e-9)))

# the scenarios that set the launch time: total cycles of the commit warp (init + phases 1..4), slowest first
tot = pr[:, 0:5].sum(1).astype(np.float64)
order = np.argsort(-tot)[:10]
print('slowest scenarios: scen pops Gcycles | per pop: lookups commit heappop help+wait | dijkstra/pop resumes pushes/pop init(Mcycles)')
for i in order:
    if s['n_pops'][i] < 1024:
        continue
    p = pr[i].astype(np.float64); n = float(s['n_pops'][i])
    print('  %5d %6d %6.3f | %7.0f %7.0f %7.0f %7.0f | %7.0f %6d %5.2f %8.1f' % (i, n, tot[i] / 1e9, p[1] / n, p[2] / n, p[3] / n, p[4] / n, p[10] / n, p[11], p[9] / n, p[0] / 1e6))

# did a scenario share its TPC (SM pair 2k, 2k+1) with another long scenario?  o[14] = smid << 40 | start clock
smid = (pr[:, 14] >> 40).astype(np.int64)
in_p2 = s['n_pops'] >= 1024
if len(long_):
    host = {}
    for i in np.where(in_p2)[0]:
        host.setdefault(int(smid[i]), []).append(int(i))
    rows = []
    for i in long_:
        sib = host.get(int(smid[i]) ^ 1, [])
        sib_pops = sum(int(s['n_pops'][j]) for j in sib)
        rows.append((tot[i] / 1e9, int(smid[i]), sib_pops))
    rows.sort()
    alone = [r[0] for r in rows if r[2] == 0]; part = [r[0] for r in rows if 0 < r[2] < 20000]; full = [r[0] for r in rows if r[2] >= 20000]
    print('long scenarios by TPC sibling load (Gcycles mean / max / n): sibling idle %.3f %.3f %d | sibling mid %.3f %.3f %d | sibling long %.3f %.3f %d' % (
        np.mean(alone) if alone else 0, max(alone) if alone else 0, len(alone), np.mean(part) if part else 0, max(part) if part else 0, len(part),
        np.mean(full) if full else 0, max(full) if full else 0, len(full)))
    print('  (Gcycles, smid, pops on the sibling SM): ' + ' '.join('%.2f/%d/%d' % r for r in rows))

if len(long_):
    sub = wp[long_][:, 8:10, 16:24].reshape(len(long_), 16).astype(np.float64) / s['n_pops'][long_].astype(np.float64)[:, None]
    print('serial section A->B of the commit warp, cycles/pop (long scenarios): ' +
          ' '.join('%s %.0f' % (nm, v) for nm, v in zip(['A..start', 'ctl', 'accept', 'lookup loads', 'skip+g+h', 'predict', 'target', 'queue init', 'tick'], sub.mean(0)[:9])))
