import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import numpy as np
from automatedvaletparking_b200 import scenarios as S
from automatedvaletparking_b200.batch import DevicePlanner
import oracle_lib as O
os.environ.setdefault('AVP_HOST_TIMEOUT_S', '20')
cases = [int(a) for a in sys.argv[1:]] or [1]
scs = [S.benchmark_case(i) for i in cases]
dp = DevicePlanner(max_pops=20000); dp.load(scs)
res = dp.plan(cap_path=512, cap_pops=20000)
for k, sc in enumerate(scs):
    m = O.OracleMap(sc); r = O.plan(m, dp.cfg)
    gp = res.pop_indices(k); op = r['pops']
    n = min(len(gp), len(op)); bad = [i for i in range(n) if gp[i] != op[i]]
    print('Case', cases[k], 'n', len(gp), len(op), 'first mismatch', bad[:5])
    if bad:
        i = bad[0]
        print(' gpu', gp[max(0,i-2):i+6]); print(' orc', op[max(0,i-2):i+6])
        fo = {int(a): f for a, f in zip(op, r['pop_fgh'])}
        for idx in list(gp[i:i+4]) + list(op[i:i+4]):
            print('   node', idx, 'oracle fgh', fo.get(int(idx)), [float(v).hex() for v in fo.get(int(idx), [])])
        # expand parent of both to compare rsL
        for idx in [int(gp[i]), int(op[i])]:
            par = (idx - 1) // 10 * 10  # global_index at creation; parent is the node popped at that expansion number
            e = (idx - 1) // 10
            pst = r['pop_state'][e] if e < len(r['pop_state']) else None
            # expansion e corresponds to pop e only if no pop ended early; ok for debugging
            gpz, gf, gl = dp.expand_pure(k, pst); opz, of, ol = m.expand_pure(dp.cfg, pst)
            j = (idx - 1) % 10
            print('   child', idx, 'of pop', e, 'gpu rsL', gl[j].hex(), 'orc rsL', ol[j].hex(), 'pose eq', np.array_equal(gpz[j], opz[j]))
