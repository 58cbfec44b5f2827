import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import numpy as np
from automatedvaletparking_b200 import scenarios as S
from automatedvaletparking_b200.batch import DevicePlanner
import oracle_lib as O
os.environ.setdefault('AVP_HOST_TIMEOUT_S', '20')
for cases in [[17], [20], [1], [4, 16, 3]]:
    scs = [S.benchmark_case(i) for i in cases]
    dp = DevicePlanner(max_pops=20000)
    dp.load(scs)
    t = time.time()
    try:
        res = dp.plan(cap_path=512, cap_pops=20000)
    except Exception as e:
        print(cases, 'FAILED', str(e)[:600]); sys.stdout.flush()
        os._exit(1)
    print(cases, 'plan %.3fs' % (time.time() - t), [(res.status_name(k), int(res.summaries['n_pops'][k]), int(res.summaries['h_closed'][k])) for k in range(len(cases))])
    for k, sc in enumerate(scs):
        r = O.plan(O.OracleMap(sc), dp.cfg)
        print('   Case', cases[k], 'pops_eq', np.array_equal(res.pop_indices(k), r['pops']), 'h_closed', int(res.summaries['h_closed'][k]), r['h_closed'], 'status', int(res.summaries['status'][k]), r['status'])
    dp.close()
