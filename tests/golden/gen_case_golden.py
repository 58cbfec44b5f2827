"""Generate known-answer fixtures by running the UNMODIFIED reference planner.

Usage (in the build container only; /root/reference must exist):
    python tests/golden/gen_case_golden.py Case1 [Case2 ...]      # one process per case
    python tests/golden/gen_case_golden.py --csv path/to/x.csv --out name

Writes tests/golden/cases/<name>.npz.  Recording is done by wrapping bound
methods of the live objects at run time (the reference files are not edited):

  * every node taken from the open list (path_planner.py:70) -> pops / pop_state / pop_fgh
  * every Dijkstra.compute_path call (compute_h.py:198)       -> hq (target id, dist, |closedlist|)
  * every calc_node_heuristic / calc_node_cost result         -> sha256 trace digest + counts
  * the a_star_plan / split_path outputs                      -> astar_path, rs_*, final_path, split_*
"""
import argparse
import hashlib
import io
import os
import struct
import sys
import time
import contextlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

ref_shim.install()
from map import costmap  # noqa: E402
from path_plan import path_planner  # noqa: E402


class _Truncate(Exception):
    pass


def run_case(csv_path, out_path, max_pops=None):
    cfg = ref_shim.default_config()
    rec = {}
    t0 = time.time()
    park_map = costmap.Map(file=csv_path, discrete_size=cfg['map_discrete_size'])
    vehicle = costmap.Vehicle()
    rec['dims'] = np.array(park_map.cost_map.shape, dtype=np.int32)
    rec['boundary'] = np.array(park_map.boundary, dtype=np.float64)
    rec['pitch'] = np.array([park_map._discrete_x, park_map._discrete_y], dtype=np.float64)
    ix, iy = np.where(park_map.cost_map == 255)
    rec['obs_cells'] = np.stack([ix, iy], axis=1).astype(np.uint16)
    rec['case_pose'] = np.array([park_map.case.x0, park_map.case.y0, park_map.case.theta0,
                                 park_map.case.xf, park_map.case.yf, park_map.case.thetaf])
    t_map = time.time() - t0

    hq = []
    t1 = time.time()
    # wrap Dijkstra.compute_path at class level for the eager call in __init__
    from path_plan import compute_h
    orig_cp = compute_h.Dijkstra.compute_path

    def cp(self, node_x, node_y):
        d, cl = orig_cp(self, node_x, node_y)
        hq.append((int(self.terminate_grid_id), int(d), len(cl)))
        return d, cl

    compute_h.Dijkstra.compute_path = cp
    with contextlib.redirect_stdout(io.StringIO()):
        planner = path_planner.PathPlanner(config=cfg, map=park_map, vehicle=vehicle)
    t_init = time.time() - t1
    astar = planner.planner
    rec['H0'] = np.int64(len(astar.h_value_list))

    pops, pop_state, pop_fgh, pop_parent = [], [], [], []
    digest = hashlib.sha256()
    counts = {'h': 0, 'g': 0}

    orig_trg = astar.try_reach_goal

    def trg(node):
        if max_pops is not None and len(pops) >= max_pops:
            raise _Truncate()          # Cases 7, 8, 19 never finish in the reference: pin a prefix
        pops.append(int(node.index))
        pop_state.append((float(node.x), float(node.y), float(node.theta)))
        pop_fgh.append((float(node.f), float(node.g), float(node.h)))
        pop_parent.append(-1 if node.parent_index is None else int(node.parent_index))
        return orig_trg(node)

    astar.try_reach_goal = trg

    orig_h = astar.calc_node_heuristic

    def ch(node):
        h = orig_h(node)
        digest.update(struct.pack('<cqdddd', b'h', int(node.index), float(node.x), float(node.y),
                                  float(node.theta), float(h)))
        counts['h'] += 1
        return h

    astar.calc_node_heuristic = ch

    orig_g = astar.calc_node_cost

    def cg(node, father_theta, father_gear):
        g = orig_g(node, father_theta=father_theta, father_gear=father_gear)
        digest.update(struct.pack('<cqd', b'g', int(node.index), float(g)))
        counts['g'] += 1
        return g

    astar.calc_node_cost = cg

    status = 'ok'
    err = ''
    t2 = time.time()
    result = None
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            result = planner.path_planning()
    except AttributeError as e:  # open list exhausted (path_planner.py:104)
        status = 'open_exhausted'
        err = repr(e)
    except _Truncate:
        status = 'truncated'
    t_search = time.time() - t2

    rec['status'] = np.array(status)
    rec['err'] = np.array(err)
    rec['pops'] = np.array(pops, dtype=np.int32)
    rec['pop_state'] = np.array(pop_state, dtype=np.float64).reshape(-1, 3)
    rec['pop_fgh'] = np.array(pop_fgh, dtype=np.float64).reshape(-1, 3)
    rec['pop_parent'] = np.array(pop_parent, dtype=np.int32)
    rec['hq'] = np.array(hq, dtype=np.int64).reshape(-1, 3)
    rec['global_index'] = np.int64(astar.global_index)
    rec['n_closed'] = np.int64(len(astar.closed_list))
    rec['n_open'] = np.int64(len(astar.open_list.queue))
    rec['n_hcalls'] = np.int64(counts['h'])
    rec['n_gcalls'] = np.int64(counts['g'])
    rec['trace_sha256'] = np.array(digest.hexdigest())
    rec['H_end'] = np.int64(len(astar.heuristic.closedlist))
    rec['times'] = np.array([t_map, t_init, t_search])
    if result is not None:
        out_final_path, info, split_list = result
        rs = info['rs_path']
        rec['astar_path'] = np.array(info['astar_path'], dtype=np.float64).reshape(-1, 3)
        rec['rs_lengths'] = np.array(rs.lengths, dtype=np.float64)
        rec['rs_ctypes'] = np.array(''.join(rs.ctypes))
        rec['rs_L'] = np.float64(rs.L)
        rec['rs_x'] = np.array(rs.x, dtype=np.float64)
        rec['rs_y'] = np.array(rs.y, dtype=np.float64)
        rec['rs_yaw'] = np.array(rs.yaw, dtype=np.float64)
        rec['rs_dir'] = np.array(rs.directions, dtype=np.int32)
        rec['out_final_path'] = np.array(out_final_path, dtype=np.float64).reshape(-1, 3)
        rec['split_lens'] = np.array([len(s) for s in split_list], dtype=np.int32)
        rec['change_gear'] = np.int64(info['change_gear'])
    os.makedirs(os.path.dirname(out_path), exist_ok=True)
    np.savez_compressed(out_path, **rec)
    print(f"{os.path.basename(out_path)}: status={status} pops={len(pops)} G={astar.global_index} "
          f"t_map={t_map:.1f}s t_init={t_init:.1f}s t_search={t_search:.1f}s", flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('cases', nargs='*')
    ap.add_argument('--csv')
    ap.add_argument('--out')
    ap.add_argument('--max-pops', type=int, default=None)
    a = ap.parse_args()
    if a.csv:
        run_case(a.csv, os.path.join(HERE, 'cases', a.out + '.npz'), a.max_pops)
        return
    for c in a.cases:
        run_case(os.path.join(ref_shim.REF_ROOT, 'BenchmarkCases', c + '.csv'),
                 os.path.join(HERE, 'cases', c + '.npz'), a.max_pops)


if __name__ == '__main__':
    main()
