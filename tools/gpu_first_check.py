"""Quick GPU-vs-oracle comparison (development aid; the real tests are in tests/)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import numpy as np
import oracle_lib as O
from automatedvaletparking_b200 import scenarios as S
from automatedvaletparking_b200.batch import DevicePlanner

cases = [int(a) for a in sys.argv[1:]] or [1, 4, 16, 17, 20, 3, 2, 6, 14, 15, 18]
scs = [S.benchmark_case(i) for i in cases]
t = time.time()
dp = DevicePlanner(max_pops=20000)
print('device', dp.device_info(), 'create %.2fs' % (time.time() - t))
t = time.time(); dp.load(scs); print('load+raster %.3fs' % (time.time() - t))
cfg = dp.cfg
rng = np.random.default_rng(0)
for k, sc in enumerate(scs):
    m = O.OracleMap(sc)
    info = dp.map_info(k)
    cm = dp.cost_map(k)
    ok_map = (info['nx'], info['ny']) == (m.nx, m.ny) and info['dx'] == m.dx and info['dy'] == m.dy and np.array_equal(cm, m.cost_map())
    # collision
    P = np.stack([rng.uniform(m.boundary[0] + 5, m.boundary[1] - 5, 400), rng.uniform(m.boundary[2] + 5, m.boundary[3] - 5, 400), rng.uniform(-np.pi, np.pi, 400)], 1)
    g = dp.check(k, P); o = np.array([m.check(cfg, *p) for p in P])
    # expand pure
    par = np.array([sc.x0, sc.y0, O.lib().orc_pi_2_pi(sc.theta0)])
    gp, gf, gl = dp.expand_pure(k, par); op, of, ol = m.expand_pure(cfg, par)
    print(f'Case{cases[k]}: map {ok_map} n_obs {info["n_obs"]}/{m.n_obs} coll mismatch {int((g!=o).sum())}/400 (hits {int(o.sum())}) '
          f'expand pose_eq {np.array_equal(gp, op)} flags_eq {np.array_equal(gf, of)} rsL maxrel {np.nanmax(np.abs(gl-ol)/np.abs(ol)):.2e}')
t = time.time(); res = dp.plan(cap_path=512, cap_pops=20000); tp = time.time() - t
print('plan %.3fs' % tp, 'launches', dp.launches)
for k, sc in enumerate(scs):
    m = O.OracleMap(sc)
    t = time.time(); r = O.plan(m, cfg); to = time.time() - t
    s = res.summaries[k]
    keys = ['status', 'n_pops', 'global_index', 'n_closed', 'n_open', 'n_astar', 'n_rs', 'n_final', 'n_hq', 'h_closed', 'n_hcalls']
    diffs = {q: (int(s[q]), r[q]) for q in keys if int(s[q]) != r[q]}
    pops_eq = np.array_equal(res.pop_indices(k), r['pops'])
    fp = res.path(k)
    pd = np.abs(fp - r['final_path']).max() if fp.shape == r['final_path'].shape and fp.size else (0.0 if fp.shape == r['final_path'].shape else -1)
    hq = dp.hq_log(k, int(s['n_hq'])); hq_eq = np.array_equal(hq, r['hq'][:len(hq)].astype(np.int32))
    hv = dp.hvalues(k)
    print(f'Case{cases[k]}: status {res.status_name(k)} pops_eq {pops_eq} ({int(s["n_pops"])}) diffs {diffs} path maxabs {pd:.3e} rs {s["rs_ctypes"].decode()}/{r["rs_ctypes"]} L rel {abs(s["rs_L"]-r["rs_L"])/(abs(r["rs_L"])+1e-300):.1e} hq_eq {hq_eq} oracle {to:.2f}s')
