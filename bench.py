#!/usr/bin/env python
"""bench.py -- collision-checked successor evaluations/s of the B200 hybrid-A* hot path.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched by torch.distributed.run)
    python bench.py --impl reference --gpus N --steps K --warmup W

Workload (BASELINE.json configs[1]): Case1's obstacle map, 1024 seeded random start/goal poses
per GPU (SURVEY §8d C2 recipe; every scenario gets its own raster).  One step = one pass of the
hot path over the batch: rasterise (costmap.py:178-261) + compute_h (lazily, inside the search)
+ hybrid-A* search with rs shots (path_planner.py:58-110).  Unit = one (scenario, expanded node,
primitive) slot of expand_node = sum of global_index (hybrid_a_star.py:239).

  value : inputs (poses + polygons) resident in HBM, results left in HBM; CUDA events.
  e2e   : the public API with HOST buffers: upload (H2D) + rasterise + search + fetch (D2H) [+ the
          NCCL all-gather of trajectories at N>1]; CUDA events around the whole sequence.
  roofline : the search kernel; achieved = sum((12*N_obs + 63.2) * successors) / CUDA-event kernel
          time (SURVEY §8d byte model) against MEASURED_PEAKS.json hbm_gbs.
  cpu_baseline : the oracle port (oracle/avp_oracle.c, the reference's algorithm in C) on the host
          cores, same scenarios (rank 0, N=1 only).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "collision-checked successor evaluations/sec"
UNIT = "successors/s"
N_SCEN = 1024
CAP_PATH = 256


def make_candidates(rank: int, n: int = N_SCEN):
    from automatedvaletparking_b200 import scenarios as scn
    return scn.perturbed_candidates(scn.benchmark_case(1), 8 * n, seed=1 + 1000 * rank)


def make_scenarios(rank: int, n: int = N_SCEN, dp=None):
    """SURVEY 8d C2 recipe incl. the rule that start and goal poses are collision free.  The check
    runs on the GPU (the product's own checker) when a DevicePlanner is given, else on the oracle."""
    from automatedvaletparking_b200 import scenarios as scn
    cands = make_candidates(rank, n)
    if dp is not None:
        dp.load(cands)
        a, b = dp.start_goal_collisions()
    else:
        import oracle_lib as O
        from automatedvaletparking_b200.hostcfg import make_avp_config
        from automatedvaletparking_b200.batch import _pi_2_pi
        cfg = make_avp_config()
        a, b = [], []
        for s in cands:
            m = O.OracleMap(s)
            a.append(m.check(cfg, s.x0, s.y0, _pi_2_pi(s.theta0)))
            b.append(m.check(cfg, s.xf, s.yf, _pi_2_pi(s.thetaf)))
            if len(a) - sum(x or y for x, y in zip(a, b)) >= n:
                break
        cands = cands[:len(a)]
    return scn.keep_collision_free(cands, a, b, n)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.p, self.first = index, [], None, 0

    def mark(self):
        """samples from here on belong to the timed region (nvidia-smi itself is started before the warm-up:
        its start-up takes driver locks for a few hundred ms and must not sit inside the timed steps)"""
        self.first = len(self.rows)

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.rows.append([t.strip() for t in line.split(",")])

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        rows = self.rows[self.first:] or self.rows
        sm = [float(r[0]) for r in rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for j, n in enumerate(names) if any(len(r) > 3 + j and r[3 + j].lower().startswith("active") for r in rows)]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


def oracle_run(scs, cfg, threads: int):
    """The reference's algorithm restated in C (oracle), one scenario per call, `threads` host threads."""
    import oracle_lib as O
    from concurrent.futures import ThreadPoolExecutor
    O.lib()

    def one(sc):
        m = O.OracleMap(sc)
        r = O.plan(m, cfg, cap_pops=1, cap_path=CAP_PATH)
        return r["global_index"], r["status"]

    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=threads) as ex:
        out = list(ex.map(one, scs))
    dt = time.perf_counter() - t0
    return sum(o[0] for o in out), dt, sum(1 for o in out if o[1] == 0)


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path (oracle port; the Python
    original cannot travel to the GPU box) on all host threads, same config/metric."""
    if rank != 0:
        return
    from automatedvaletparking_b200.hostcfg import make_avp_config
    cfg = make_avp_config(max_pops=20000)
    cores = os.cpu_count() or 1
    n_sample = min(N_SCEN, 32 * cores)
    scs = make_scenarios(0, n_sample)
    for _ in range(max(0, min(args.warmup, 1))):
        oracle_run(scs[:cores], cfg, cores)
    vals, times = [], []
    for _ in range(args.steps):
        succ, dt, _ = oracle_run(scs, cfg, cores)
        vals.append(succ / dt)
        times.append(dt)
    v = statistics.mean(vals)
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * statistics.mean(times), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "Case1 obstacle map, 1024 randomised start/goal poses per GPU (SURVEY 8d C2)",
                       "scenarios_per_step": n_sample, "note": "CPU port of the reference algorithm; a step is a bounded sample"},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"first {n_sample} of the {N_SCEN} scenarios of rank 0, {cores} threads, oracle/avp_oracle.c"},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--scenarios", type=int, default=N_SCEN)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import __graft_entry__ as graft
    graft.build()
    import torch
    from automatedvaletparking_b200 import scenarios as scn
    from automatedvaletparking_b200.batch import DevicePlanner
    from automatedvaletparking_b200 import distributed as avd

    use_dist = world > 1
    if use_dist:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")

    def barrier():
        if use_dist:
            dist.barrier()
        torch.cuda.synchronize()

    n = args.scenarios
    os.environ.setdefault("AVP_HOST_TIMEOUT_S", "600")
    dp = DevicePlanner(device=local_rank, max_pops=20000)
    scs = make_scenarios(rank, n, dp)
    batch = scn.pack(scs)
    dp.load(batch)                       # inputs resident: poses + polygons in HBM
    W = max(3, args.warmup)
    K = args.steps

    # ---------------- resident leg (value): rasterise + search, results stay on the device
    clocks = ClockSampler(local_rank)
    clocks.start()
    for _ in range(W):
        dp.rasterise()
        dp.plan_resident(CAP_PATH, 0)
    barrier()
    clocks.mark()
    l0 = dp.launches
    dp.timer_start()
    search_ms = []
    for _ in range(K):
        dp.rasterise()
        dp.plan_resident(CAP_PATH, 0)
        search_ms.append(dp.last_search_ms())
    ms_res = dp.timer_stop()
    barrier()
    clk = clocks.stop()
    launches = dp.launches - l0
    res = dp.fetch(CAP_PATH, 0)
    succ = res.successors
    s = res.summaries
    bytes_model = float(((12.0 * s["n_obs"].astype(np.float64) + 63.2) * s["global_index"].astype(np.float64)).sum())
    plans_ok = int((s["status"] == 0).sum())

    # ---------------- e2e leg: host buffers -> upload + rasterise + search + fetch (+ all-gather)
    for _ in range(2):
        dp.load(batch)
        dp.plan(CAP_PATH, 0)
    barrier()
    dp.h2d_bytes = 0
    dp.d2h_bytes = 0
    dp.timer_start()
    for _ in range(K):
        dp.load(batch)
        r2 = dp.plan(CAP_PATH, 0)
        if use_dist:
            avd.gather_results_device(dp, n, CAP_PATH)
            torch.cuda.synchronize()
    ms_e2e = dp.timer_stop()
    barrier()
    h2d, d2h = dp.h2d_bytes // K, dp.d2h_bytes // K
    assert np.array_equal(r2.summaries["global_index"], s["global_index"])

    # ---------------- reduce over ranks (max time, sum of units)
    tot_succ, t_res, t_e2e, t_search, tot_bytes, tot_ok = succ, ms_res, ms_e2e, statistics.mean(search_ms), bytes_model, plans_ok
    if use_dist:
        t = torch.tensor([ms_res, ms_e2e, statistics.mean(search_ms)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        u = torch.tensor([float(succ), bytes_model, float(plans_ok)], dtype=torch.float64, device="cuda")
        dist.all_reduce(u, op=dist.ReduceOp.SUM)
        t_res, t_e2e, t_search = t.tolist()
        tot_succ, tot_bytes, tot_ok = u.tolist()

    if rank == 0:
        peak, peak_kind = hbm_peak()
        value = tot_succ * K / (t_res * 1e-3)
        e2e_v = tot_succ * K / (t_e2e * 1e-3)
        achieved = (bytes_model / (statistics.mean(search_ms) * 1e-3)) / 1e9          # this rank's kernel, GB/s
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp)).get("k_search_dram_bytes_per_launch")
            except Exception:
                traffic = None
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": t_res / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic",
                "config": {"workload": "Case1 obstacle map, 1024 randomised start/goal poses per GPU (BASELINE configs[1], SURVEY 8d C2)",
                           "scenarios_per_gpu": n, "plans_ok": int(tot_ok), "plans_per_s": n * world * K / (t_res * 1e-3),
                           "successors_per_step": int(tot_succ), "parallelism": f"scenario-sharded x{world}",
                           "l2": "per-step working set (h tables + rasters, ~1.8 GB/GPU) exceeds the 126 MB L2; no explicit flush",
                           "search_kernel_ms": t_search},
                "e2e": {"value": e2e_v, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                        "ms_per_step": t_e2e / K},
                "gpu_launches": int(launches),
                "clocks": clk,
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                             "traffic": traffic, "peak_source": f"{peak_kind} (MEASURED_PEAKS.json hbm_gbs)" if peak_kind == "measured" else "fallback 6.65 TB/s",
                             "kernel": "k_search", "model": "sum((12*N_obs+63.2)*successors) per launch (SURVEY 8d)"}}
        if world == 1 and not args.no_cpu_baseline:
            from automatedvaletparking_b200.hostcfg import make_avp_config
            cores = os.cpu_count() or 1
            n_sample = min(n, 32 * cores)
            csucc, cdt, _ = oracle_run(scs[:n_sample], make_avp_config(max_pops=20000), cores)
            line["cpu_baseline"] = {"value": csucc / cdt, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"first {n_sample} scenarios of the step, {cores} threads, oracle/avp_oracle.c ({cdt:.1f} s)"}
        print(json.dumps(line), flush=True)
    dp.close()
    if use_dist:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
