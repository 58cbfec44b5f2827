"""Per-function known-answer vectors from the UNMODIFIED reference (build container only).

    python tests/golden/gen_leaf_golden.py        -> tests/golden/leaf_*.npz

  leaf_geometry.npz  Vehicle.create_anticlockpoint corners, pi_2_pi, M, convert_position_to_index
  leaf_collision.npz distance_checker.check / two_circle_checker.check on Case maps, random poses
  leaf_rs.npz        rs_curve.generate_path (all retained words) + calc_optimal_path (word + course)
  leaf_dijkstra.npz  Dijkstra.compute_path query sequences (distance, len(closedlist), h table)
  leaf_maps.npz      Map rasters of perturbed / synthetic scenarios (dims, pitch, obstacle cells)
  leaf_numpy.npz     np.cos/np.sin/np.tan vs math.* (the dispatch the device sin/cos port relies on)
"""
import contextlib
import io
import math
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
import ref_shim  # noqa: E402

ref_shim.install()
from map import costmap  # noqa: E402
from collision_check import collision_check  # noqa: E402
from path_plan import rs_curve, compute_h  # noqa: E402
from automatedvaletparking_b200 import scenarios as scn  # noqa: E402

CFG = ref_shim.default_config()


def ref_map(s: scn.Scenario):
    """Build the reference Map for a Scenario (CSV round trip; boundary override per SURVEY §8d C4)."""
    with tempfile.NamedTemporaryFile('w', suffix='.csv', delete=False) as f:
        path = f.name
    scn.write_case_csv(s, path)
    if s.boundary is None:
        m = costmap.Map(file=path, discrete_size=CFG['map_discrete_size'])
    else:
        m = costmap.Map.__new__(costmap.Map)
        m.discrete_size = CFG['map_discrete_size']
        m.grid_index = None
        m.case = costmap.Case.read(path)
        m.boundary = np.array(s.boundary, dtype=np.float64)
        m._discrete_x = 0
        m._discrete_y = 0
        m.detect_obstacle_edge()
    os.unlink(path)
    return m


def gen_geometry(rng):
    v = costmap.Vehicle()
    poses = np.stack([rng.uniform(-30, 30, 3000), rng.uniform(-30, 30, 3000), rng.uniform(-3.2, 3.2, 3000)], 1)
    poses[:50, :2] += 5e9   # Cases 13-15 magnitudes
    corners = np.array([v.create_anticlockpoint(x=np.float64(p[0]), y=np.float64(p[1]), theta=np.float64(p[2]), config=CFG).reshape(5, 2)
                        for p in poses])
    th = rng.uniform(-20, 20, 5000)
    p2p = np.array([rs_curve.pi_2_pi(t) for t in th])
    M = np.array([rs_curve.M(t) for t in th])
    m = ref_map(scn.benchmark_case(1))
    pts = np.stack([rng.uniform(m.boundary[0], m.boundary[1], 5000), rng.uniform(m.boundary[2], m.boundary[3], 5000)], 1)
    idx = np.array([m.convert_position_to_index(p[0], p[1]) for p in pts], dtype=np.int64)
    np.savez_compressed(os.path.join(HERE, 'leaf_geometry.npz'), poses=poses, corners=corners, th=th, pi_2_pi=p2p, M=M,
                        idx_pts=pts, idx=idx, min_radius_turn=np.float64(v.min_radius_turn))


def gen_collision(rng):
    out = {}
    for case in (1, 5, 13, 19):
        s = scn.benchmark_case(case)
        m = ref_map(s)
        v = costmap.Vehicle()
        ox, oy = np.where(m.cost_map == 255)
        n = 1500
        # poses near obstacle cells so that ~half collide
        k = rng.integers(0, len(ox), n)
        poses = np.stack([m.map_position[0][ox[k]] + rng.uniform(-4, 4, n), m.map_position[1][oy[k]] + rng.uniform(-4, 4, n),
                          rng.uniform(-np.pi, np.pi, n)], 1)
        dc = collision_check.distance_checker(map=m, vehicle=v, config=CFG)
        cc = collision_check.two_circle_checker(map=m, vehicle=v, config=CFG)
        with np.errstate(all='ignore'):
            out[f'c{case}_poses'] = poses
            out[f'c{case}_distance'] = np.array([bool(dc.check(node_x=p[0], node_y=p[1], theta=p[2])) for p in poses])
            out[f'c{case}_circle'] = np.array([bool(cc.check(node_x=p[0], node_y=p[1], theta=p[2])) for p in poses[:400]])
    np.savez_compressed(os.path.join(HERE, 'leaf_collision.npz'), **out)


def gen_rs(rng):
    v = costmap.Vehicle()
    maxc = 1 / v.min_radius_turn          # np.float64, as on the planner path
    n = 4000
    q = np.stack([rng.uniform(-15, 15, n), rng.uniform(-15, 15, n), rng.uniform(-np.pi, np.pi, n),
                  rng.uniform(-15, 15, n), rng.uniform(-15, 15, n), rng.uniform(-np.pi, np.pi, n)], 1)
    q[:1000, 3:5] = q[:1000, 0:2] + rng.uniform(-4, 4, (1000, 2))     # close pairs: CCC / CCCC words win
    words_n = np.zeros(n, dtype=np.int32)
    words_ct, words_len, words_L = [], [], []
    sel_ct, sel_len, sel_L, sel_n = [], np.zeros((n, 5)), np.zeros(n), np.zeros(n, dtype=np.int32)
    npts = np.zeros(n, dtype=np.int32)
    cx, cy, cyaw, cdir = [], [], [], []
    for i in range(n):
        # children on the planner path carry np.float64 poses; the goal is a Python float triple
        q0 = [np.float64(q[i, 0]), np.float64(q[i, 1]), np.float64(q[i, 2])]
        q1 = [float(q[i, 3]), float(q[i, 4]), float(q[i, 5])]
        paths = rs_curve.generate_path(q0, q1, maxc)
        words_n[i] = len(paths)
        for p in paths:
            words_ct.append(''.join(p.ctypes))
            l = np.zeros(5)
            l[:len(p.lengths)] = p.lengths
            words_len.append(l)
            words_L.append(float(p.L))
        best = rs_curve.calc_optimal_path(q0[0], q0[1], q0[2], q1[0], q1[1], q1[2], maxc)
        sel_ct.append(''.join(best.ctypes))
        sel_n[i] = len(best.lengths)
        sel_len[i, :sel_n[i]] = best.lengths
        sel_L[i] = best.L
        npts[i] = len(best.x)
        cx += list(best.x)
        cy += list(best.y)
        cyaw += list(best.yaw)
        cdir += list(best.directions)
    # root-node typing: theta values are Python floats (phi is a float, x/y still numpy through maxc)
    nr = 500
    root_L = np.zeros(nr)
    root_ct = []
    for i in range(nr):
        best = rs_curve.calc_optimal_path(float(q[i, 0]), float(q[i, 1]), float(q[i, 2]), float(q[i, 3]), float(q[i, 4]),
                                          float(q[i, 5]), maxc)
        root_L[i] = best.L
        root_ct.append(''.join(best.ctypes))
    np.savez_compressed(os.path.join(HERE, 'leaf_rs.npz'), q=q, maxc=np.float64(maxc), words_n=words_n,
                        words_ct=np.array(words_ct), words_len=np.array(words_len), words_L=np.array(words_L),
                        sel_ct=np.array(sel_ct), sel_n=sel_n, sel_len=sel_len, sel_L=sel_L, npts=npts,
                        cx=np.array(cx), cy=np.array(cy), cyaw=np.array(cyaw), cdir=np.array(cdir, dtype=np.int32),
                        root_L=root_L, root_ct=np.array(root_ct))


def gen_dijkstra(rng):
    out = {}
    for case in (1, 4):
        s = scn.benchmark_case(case)
        m = ref_map(s)
        d = compute_h.Dijkstra(m)
        # queries at growing distance from the goal (the python Dijkstra is O(n^2): keep it small)
        qs = []
        for r in (1.0, 2.0, 3.0, 4.0, 5.0, 6.0):
            a = rng.uniform(0, 2 * np.pi)
            qs.append((s.xf + r * np.cos(a), s.yf + r * np.sin(a)))
        res = []
        for (x, y) in qs:
            gid = m.convert_position_to_index(x, y)
            hit = [g.distance for g in d.closedlist if g.grid_id == gid]
            if hit:                       # calc_node_heuristic would not call compute_path
                res.append((gid, hit[0], len(d.closedlist), 0))
                continue
            dist, cl = d.compute_path(x, y)
            res.append((gid, dist, len(cl), 1))
        first = {}
        for g in d.closedlist:
            first.setdefault(int(g.grid_id), int(g.distance))
        out[f'c{case}_queries'] = np.array(qs)
        out[f'c{case}_res'] = np.array(res, dtype=np.int64)
        out[f'c{case}_h_ids'] = np.array(sorted(first), dtype=np.int64)
        out[f'c{case}_h_val'] = np.array([first[k] for k in sorted(first)], dtype=np.int64)
    np.savez_compressed(os.path.join(HERE, 'leaf_dijkstra.npz'), **out)


def gen_maps(rng):
    out = {}
    scs = []
    for case in (1, 7, 13, 19, 20):
        scs += scn.perturbed_set(scn.benchmark_case(case), 4, seed=100 + case)
    scs += scn.synthetic_set(2, 2, seed=4)
    rows = []
    for i, s in enumerate(scs):
        m = ref_map(s)
        ix, iy = np.where(m.cost_map == 255)
        out[f'm{i}_cells'] = np.stack([ix, iy], 1).astype(np.uint16)
        out[f'm{i}_geom'] = np.array([m.cost_map.shape[0], m.cost_map.shape[1], *m.boundary, m._discrete_x, m._discrete_y])
        rows.append(scn.case_row(s) + ([] if s.boundary is None else []))
        out[f'm{i}_row'] = np.array(scn.case_row(s))
        out[f'm{i}_boundary'] = np.array(s.boundary if s.boundary is not None else [np.nan] * 4)
    out['n'] = np.int64(len(scs))
    np.savez_compressed(os.path.join(HERE, 'leaf_maps.npz'), **out)


def gen_numpy(rng):
    th = np.concatenate([rng.uniform(-7, 7, 2000000), rng.uniform(-100, 100, 200000)])
    ok_cos = bool(np.all(np.cos(th) == np.array([math.cos(t) for t in th])))
    ok_sin = bool(np.all(np.sin(th) == np.array([math.sin(t) for t in th])))
    ok_scalar = all(np.cos(np.float64(t)) == math.cos(t) and np.sin(np.float64(t)) == math.sin(t) for t in th[:200000])
    st = np.linspace(-0.75, 0.75, 5)
    np.savez_compressed(os.path.join(HERE, 'leaf_numpy.npz'), cos_eq=ok_cos, sin_eq=ok_sin, scalar_eq=ok_scalar,
                        steer=st, tan_steer=np.tan(st), math_tan=np.array([math.tan(t) for t in st]),
                        sample_th=th[:20000], sample_cos=np.cos(th[:20000]), sample_sin=np.sin(th[:20000]))
    print('np.cos==math.cos', ok_cos, 'np.sin==math.sin', ok_sin, 'scalar', ok_scalar)


if __name__ == '__main__':
    which = sys.argv[1:] or ['geometry', 'collision', 'rs', 'dijkstra', 'maps', 'numpy']
    for w in which:
        with contextlib.redirect_stdout(io.StringIO()) if w != 'numpy' else contextlib.nullcontext():
            globals()['gen_' + w](np.random.default_rng(abs(hash(w)) % 1000 + 7 if False else {'geometry': 11, 'collision': 12, 'rs': 13, 'dijkstra': 14, 'maps': 15, 'numpy': 16}[w]))
        print('wrote leaf_' + w, flush=True)
