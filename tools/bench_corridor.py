"""Throughput of the corridor extraction kernel (k_corridor) next to the oracle on the host.
One JSON line: path points / s through the C ABI with HOST buffers (H2D + kernel + D2H inside the timed region,
CUDA events) and the kernel's algorithmic traffic: every path point streams the obstacle cells of the raster
columns under its inflated, expanded AABB (16 B per cell) and writes 36 B."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import numpy as np
import oracle_lib as O
from automatedvaletparking_b200 import scenarios as scn
from automatedvaletparking_b200.batch import DevicePlanner

case = int(sys.argv[1]) if len(sys.argv) > 1 else 19
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 20
dp = DevicePlanner()
sc = scn.benchmark_case(case)
dp.load([sc])
m = O.OracleMap(sc)
b = m.boundary
rng = np.random.default_rng(5)
P = np.stack([rng.uniform(b[0], b[1], n), rng.uniform(b[2], b[3], n), rng.uniform(-np.pi, np.pi, n)], 1)
for _ in range(3):
    dp.corridor(0, P, 0.8)
dp.timer_start()
K = 5
for _ in range(K):
    d, st = dp.corridor(0, P, 0.8)
ms = dp.timer_stop() / K
ns = 20000
t0 = time.perf_counter()
od, _ = m.corridor(dp.cfg, P[:ns], 0.8)
cpu = ns / (time.perf_counter() - t0)
assert ((d[:ns] == od) | (np.isnan(d[:ns]) & np.isnan(od))).all()
# algorithmic bytes: cells in the column range of the expanded AABB (half extent <= hypot(4.889, 2.142)/2 + 0.8 m), 16 B each
xs = np.sort(m.positions()[0][np.where(m.cost_map() == 255)[0]])
half = 0.5 * np.hypot(4.889, 2.142) + 0.8
cells = np.searchsorted(xs, P[:, 0] + half) - np.searchsorted(xs, P[:, 0] - half)
bytes_ = float(cells.sum()) * 16 + 36.0 * n
print(json.dumps({"metric": "corridor path points/sec (compute_collision_H)", "value": n / (ms * 1e-3), "unit": "points/s", "ms": ms, "points": n,
                  "workload": f"Case{case} raster ({m.n_obs} obstacle cells), {n} random poses, host buffers in/out",
                  "model_bytes": bytes_, "model_GBps": bytes_ / (ms * 1e-3) / 1e9,
                  "cpu_oracle_points_per_s_1thread": cpu, "parity": "first %d points bit-identical to the oracle" % ns}))
