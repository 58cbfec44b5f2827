#!/usr/bin/env python
"""bench.py -- collision-checked successor evaluations/s of the B200 hybrid-A* hot path.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched by torch.distributed.run)
    python bench.py --impl reference --gpus N --steps K --warmup W
    python bench.py --workload c3|c4 ...                     (the other BASELINE configs as headline; default c2)

Workloads (BASELINE.json configs, SURVEY §8d recipes; every scenario gets its own raster):
  c2  configs[1]: Case1's obstacle map, 1024 seeded random start/goal poses PER GPU (weak scaling, one seed per rank).
  c3  configs[2]: all 20 BenchmarkCases x 256 perturbed start/goal poses = 5120 scenarios, dealt to the ranks with
      distributed.shard_indices (strong scaling).
  c4  configs[3]: 64 synthetic 200x200 maps with 256 convex polygons x 64 start/goal pairs = 4096 scenarios (strong scaling).
The JSON line's headline is the --workload (default c2, the configuration BASELINE's metric is quoted on); the default run
also times c3 and c4 (fewer steps) and reports them under "workloads".

One step = one pass of the hot path over the batch: rasterise (costmap.py:178-261) + eager Dijkstra (hybrid_a_star.py:89-91)
+ hybrid-A* search with the lazy heuristic and rs shots (path_planner.py:58-110).  Unit = one (scenario, expanded node,
primitive) slot of expand_node = sum of global_index (hybrid_a_star.py:239).

  value : inputs (poses + polygons) resident in HBM, results left in HBM; CUDA events.
  e2e   : the public API with HOST buffers: upload (H2D) + rasterise + search + fetch (D2H) [+ the NCCL all-gather of
          trajectories at N>1]; CUDA events around the whole sequence.
  roofline : the search kernel (k_plan); achieved = sum((12*N_obs + 63.2) * successors) / CUDA-event kernel time
          (SURVEY §8d byte model) against MEASURED_PEAKS.json hbm_gbs.
  cpu_baseline : the oracle port (oracle/avp_oracle.c, the reference's algorithm in C) on the host cores, a bounded sample
          of the same scenarios (rank 0, N=1 only).
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
STATUS = {0: "ok", 1: "open_exhausted", 2: "open_exhausted_rs", 3: "h_unreachable", 4: "rs_degenerate", 5: "capped_max_pops", 6: "raster_ambiguous"}
WORKLOAD_TEXT = {
    "c2": "Case1 obstacle map, 1024 randomised start/goal poses per GPU (BASELINE configs[1], SURVEY 8d C2)",
    "c3": "all 20 BenchmarkCases x 256 perturbed start/goal poses = 5120 scenarios, sharded over the GPUs (BASELINE configs[2], SURVEY 8d C3)",
    "c4": "64 synthetic 200x200 maps x 256 convex polygons x 64 start/goal pairs = 4096 scenarios, sharded over the GPUs (BASELINE configs[3], SURVEY 8d C4); "
          "at this obstacle density no pose is collision free, every search ends at its first pop: the step is rasterisation + eager Dijkstra",
    "c4s": "as c4 with 24 polygons per map and collision-free start/goal poses (searches that run): 4096 scenarios, sharded over the GPUs",
}


# ------------------------------------------------------------------------------------------------ scenario recipes

def make_candidates(rank: int, n: int = N_SCEN):
    from automatedvaletparking_b200 import scenarios as scn
    return scn.perturbed_candidates(scn.benchmark_case(1), 8 * n, seed=1 + 1000 * rank)


def _collisions(cands, dp, cfg=None, need=None):
    """start / goal collision flags of the candidates: on the GPU (the product's own checker, one launch per 4096 candidates)
    when a DevicePlanner is given, else on the oracle (the reference arm; both agree bit for bit, tests/test_gpu_parity.py).
    With `need`, stops after the chunk in which the need-th collision-free candidate was seen; returns flags for the
    candidates looked at (a prefix)."""
    a, b, free = [], [], 0
    chunk = 4096 if dp is not None else 64               # bounded uploads: every candidate gets its own raster and h-table space
    if dp is None:
        import oracle_lib as O
        from automatedvaletparking_b200.batch import _pi_2_pi
    for k in range(0, len(cands), chunk):
        part = cands[k:k + chunk]
        if dp is not None:
            dp.load(part)
            ak, bk = dp.start_goal_collisions()
        else:
            ak, bk = np.zeros(len(part), dtype=bool), np.zeros(len(part), dtype=bool)
            for i, s in enumerate(part):
                m = O.OracleMap(s)
                ak[i] = m.check(cfg, s.x0, s.y0, _pi_2_pi(s.theta0))
                bk[i] = m.check(cfg, s.xf, s.yf, _pi_2_pi(s.thetaf))
        a.append(ak); b.append(bk)
        free += int((~(ak | bk)).sum())
        if need is not None and free >= need:
            break
    return np.concatenate(a), np.concatenate(b)


def make_scenarios(rank: int, n: int = N_SCEN, dp=None):
    """SURVEY 8d C2 recipe incl. the rule that start and goal poses are collision free."""
    from automatedvaletparking_b200 import scenarios as scn
    from automatedvaletparking_b200.hostcfg import make_avp_config
    cands = make_candidates(rank, n)
    a, b = _collisions(cands, dp, make_avp_config(), need=n)
    return scn.keep_collision_free(cands[:len(a)], a, b, n)


def make_c3(dp=None, per_case: int = 256, cases=range(1, 21)):
    """for each case `per_case` perturbations (seed 100 + case) with the C2 recipe incl. the collision rule: the first
    `per_case` collision-free draws of the case's seeded stream (Case8 accepts 0.3 % of the draws, Case20's base start pose collides)."""
    from automatedvaletparking_b200 import scenarios as scn
    from automatedvaletparking_b200.hostcfg import make_avp_config
    cfg = make_avp_config()
    out = []
    for c in cases:
        base = scn.benchmark_case(c)
        over = 8
        while True:
            cands = scn.perturbed_candidates(base, over * per_case, seed=100 + c)      # a longer draw of the same stream: same prefix
            a, b = _collisions(cands, dp, cfg, need=per_case)
            free = [s for s, x, y in zip(cands, a, b) if not (x or y)]
            if len(free) >= per_case or over >= 4096:
                break
            over *= 8
        if len(free) < per_case:
            raise RuntimeError(f"Case{c}: only {len(free)} collision-free perturbations")
        out += free[:per_case]
    return out


def make_c4(dp=None, n_maps: int = 64, pairs: int = 64, n_poly: int = 256, collision_free: bool = False):
    """synthetic 200x200 maps (seed 4) with `n_poly` convex polygons, `pairs` start/goal draws per map.
    With 256 polygons NO pose of the 4.9 m x 2.1 m inflated footprint is collision free (0 of 4000 random poses, oracle), so the
    collision rule of the recipe cannot be applied: the draws are taken as they come and every search ends at its first pop
    (status parity; the step is rasterisation + eager Dijkstra).  `collision_free=True` (used with fewer polygons: c4s) keeps
    the first `pairs` collision-free of 64*pairs draws per map."""
    from automatedvaletparking_b200 import scenarios as scn
    from automatedvaletparking_b200.hostcfg import make_avp_config
    if not collision_free:
        return scn.synthetic_set(n_maps, pairs, seed=4, n_poly=n_poly)
    cfg = make_avp_config()
    per = 64 * pairs
    maps = scn.synthetic_candidates(n_maps, per, seed=4, n_poly=n_poly)
    out = []
    for m in maps:
        a, b = _collisions(m, dp, cfg, need=pairs)
        free = [s for s, x, y in zip(m, a, b) if not (x or y)]
        if len(free) < pairs:
            raise RuntimeError(f"synthetic map: only {len(free)} of {per} draws are collision free")
        out += free[:pairs]
    return out


def make_workload(name: str, rank: int, world: int, dp, n_c2: int = N_SCEN):
    """-> (this rank's scenarios, total scenario count of the job, scaling, global ids of this rank's scenarios, shard keys)"""
    from automatedvaletparking_b200 import distributed as avd
    if name == "c2":
        scs = make_scenarios(rank, n_c2, dp)
        return scs, n_c2 * world, "weak", np.arange(rank * n_c2, (rank + 1) * n_c2), None
    full = make_c3(dp) if name == "c3" else make_c4(dp) if name == "c4" else make_c4(dp, n_poly=24, collision_free=True)
    keys = [avd.cost_proxy(s) for s in full]
    idx = avd.shard_indices(len(full), rank, world, keys)
    return [full[i] for i in idx], len(full), "strong", idx, keys


# ------------------------------------------------------------------------------------------------ helpers

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


def kernel_counters():
    """whole-launch counters of the search kernel from the latest committed ncu capture (profiles/kernel_counters.json)"""
    p = os.path.join(ROOT, "profiles", "kernel_counters.json")
    try:
        return json.load(open(p))
    except Exception:
        return {}


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


def cpu_sample(name: str, scs, cores: int):
    """a bounded sample of the rank-0 scenarios for the CPU arm: per-unit rate on a sample (about 10-30 s of CPU work)"""
    if name == "c2":
        k = min(len(scs), 32 * cores)
        return scs[:k], f"first {k} of the {len(scs)} scenarios of rank 0"
    step = max(1, len(scs) // (4 * cores))              # strided: every case / map of the workload is represented
    sample = scs[::step][:4 * cores]
    return sample, f"every {step}th of the {len(scs)} scenarios ({len(sample)} scenarios)"


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path (oracle port; the Python
    original cannot travel to the GPU box) on all host threads, same config/metric."""
    if rank != 0:
        return
    from automatedvaletparking_b200.hostcfg import make_avp_config
    cfg = make_avp_config(max_pops=20000)
    cores = os.cpu_count() or 1
    name = args.workload
    if name == "c2":
        scs = make_scenarios(0, min(N_SCEN, 32 * cores))
        what = f"first {len(scs)} of the {N_SCEN} scenarios of rank 0"
    elif name == "c3":
        scs = make_c3(None, per_case=max(1, (4 * cores) // 20))
        what = f"the first {len(scs) // 20} collision-free perturbations of each of the 20 cases"
    else:
        scs = make_c4(None, n_maps=max(1, min(64, cores // 2)), pairs=8)
        what = f"the first 8 pairs of {len(scs) // 8} maps"
    for _ in range(max(0, min(args.warmup, 1))):
        oracle_run(scs[:cores], cfg, cores)
    vals, times = [], []
    for _ in range(args.steps):
        succ, dt, _ = oracle_run(scs, cfg, cores)
        vals.append(succ / dt)
        times.append(dt)
    v = statistics.mean(vals)
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * statistics.mean(times), "higher_is_better": True,
            "scaling": "weak" if name == "c2" else "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD_TEXT[name], "scenarios_per_step": len(scs),
                       "note": "CPU port of the reference algorithm (oracle/avp_oracle.c, -O2); a step is a bounded sample of the workload; "
                               "the rate is per unit, one host whatever --gpus says"},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "per_thread": v / cores,
                             "sample": f"{what}, {cores} threads, oracle/avp_oracle.c"},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ the B200 arm

def run_workload(name, dp, rank, world, local_rank, W, K, n_c2, with_clocks, cpu_baseline):
    """time one workload; returns (dict for the JSON line, rank-0 scenarios)"""
    import torch
    from automatedvaletparking_b200 import scenarios as scn
    from automatedvaletparking_b200 import distributed as avd
    use_dist = world > 1
    if use_dist:
        import torch.distributed as dist

    def barrier():
        if use_dist:
            dist.barrier()
        torch.cuda.synchronize()

    scs, n_total, scaling, gids, shard_keys = make_workload(name, rank, world, dp, n_c2)
    n = len(scs)
    batch = scn.pack(scs)
    dp.load(batch)                       # inputs resident: poses + polygons in HBM

    # ---------------- resident leg (value): rasterise + search, results stay on the device
    clocks = ClockSampler(local_rank) if with_clocks else None
    if clocks:
        clocks.start()
    for _ in range(W):
        dp.rasterise()
        dp.plan_resident(CAP_PATH, 0)
    barrier()
    if clocks:
        clocks.mark()
    l0 = dp.launches
    dp.timer_start()
    search_ms, dij_ms, plan_ms = [], [], []
    for _ in range(K):
        dp.rasterise()
        dp.plan_resident(CAP_PATH, 0)
        search_ms.append(dp.last_search_ms())
        a, b, nsus, blk = dp.last_search_passes()
        dij_ms.append(a); plan_ms.append(b)
    ms_res = dp.timer_stop()
    barrier()
    clk = clocks.stop() if clocks else None
    launches = dp.launches - l0
    res = dp.fetch(CAP_PATH, 0)
    succ = res.successors
    s = res.summaries
    bytes_model = float(((12.0 * s["n_obs"].astype(np.float64) + 63.2) * s["global_index"].astype(np.float64)).sum())
    hist = np.bincount(s["status"], minlength=7).astype(np.float64)
    capped_succ = float(s["global_index"][s["status"] == 5].astype(np.int64).sum())

    # ---------------- e2e leg: host buffers -> upload + rasterise + search + fetch (+ all-gather)
    for _ in range(min(2, W)):
        dp.load(batch)
        dp.plan(CAP_PATH, 0)
    barrier()
    dp.h2d_bytes = 0
    dp.d2h_bytes = 0
    per_rank = (n_total + world - 1) // world
    gathered = None
    dp.timer_start()
    for _ in range(K):
        dp.load(batch)
        r2 = dp.plan(CAP_PATH, 0)
        if use_dist:
            gathered = avd.gather_results_device(dp, per_rank, CAP_PATH)
            torch.cuda.synchronize()
    ms_e2e = dp.timer_stop()
    barrier()
    h2d, d2h = dp.h2d_bytes // K, dp.d2h_bytes // K
    assert np.array_equal(r2.summaries["global_index"], s["global_index"])

    # ---------------- the job's records in scenario order and their checksum: identical for every N (strong scaling)
    if use_dist:
        gs, gp = gathered
        S = gs.cpu().numpy().view(s.dtype).reshape(world, per_rank)
        Pth = gp.cpu().numpy().reshape(world, per_rank, CAP_PATH, 3)
        if scaling == "strong":
            sums_all = avd.unshard([S[r] for r in range(world)], n_total, world, shard_keys)
            paths_all = avd.unshard([Pth[r] for r in range(world)], n_total, world, shard_keys)
        else:
            sums_all = S.reshape(-1)[:n_total]
            paths_all = Pth.reshape(-1, CAP_PATH, 3)[:n_total]
    else:
        sums_all = np.zeros(n_total, dtype=s.dtype)
        paths_all = np.zeros((n_total, CAP_PATH, 3))
        sums_all[gids] = r2.summaries
        paths_all[gids] = r2.paths
    checksum = avd.records_checksum(sums_all, paths_all)

    # ---------------- reduce over ranks (max time, sum of units)
    t_search = statistics.mean(search_ms)
    per_rank_ms = [ms_res / K]
    tot = np.array([float(succ), bytes_model, capped_succ] + list(hist), dtype=np.float64)
    t_res, t_e2e, t_search_max = ms_res, ms_e2e, t_search
    if use_dist:
        t = torch.tensor([ms_res, ms_e2e, t_search], dtype=torch.float64, device="cuda")
        allt = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allt, t)
        per_rank_ms = [float(x[0]) / K for x in allt]
        t_res, t_e2e, t_search_max = (max(float(x[i]) for x in allt) for i in range(3))
        u = torch.tensor(tot, dtype=torch.float64, device="cuda")
        dist.all_reduce(u, op=dist.ReduceOp.SUM)
        tot = u.cpu().numpy()
    tot_succ, tot_bytes, tot_capped = float(tot[0]), float(tot[1]), float(tot[2])
    tot_hist = tot[3:]
    out = None
    if rank == 0:
        peak, peak_kind = hbm_peak()
        kc = kernel_counters()
        achieved = (bytes_model / (statistics.mean(plan_ms) * 1e-3)) / 1e9          # this rank's search kernel, GB/s of the byte model
        out = {
            "value": tot_succ * K / (t_res * 1e-3), "ms_per_step": t_res / K, "scaling": scaling,
            "config": {"workload": WORKLOAD_TEXT[name], "scenarios_total": int(n_total), "scenarios_rank0": int(n),
                       "plans_ok": int(tot_hist[0]), "plans_per_s": n_total * K / (t_res * 1e-3),
                       "successors_per_step": int(tot_succ),
                       "status_histogram": {STATUS[i]: int(tot_hist[i]) for i in range(7) if tot_hist[i] > 0},
                       "capped_share_of_successors": tot_capped / max(tot_succ, 1.0),
                       "note_capped": "searches that reach max_pops = 20000 return no path (the reference would search on); their expansions are "
                                      "real work and count in the metric's numerator, plans_ok / plans_per_s count only completed plans",
                       "parallelism": f"scenario-sharded x{world}",
                       "l2": "per-step working set (h tables, rasters, node arrays: > 1.8 GB/GPU) exceeds the 126 MB L2; no explicit flush",
                       "search_kernel_ms": statistics.mean(plan_ms), "dijkstra_kernel_ms": statistics.mean(dij_ms),
                       "search_total_ms_max_rank": t_search_max, "suspensions_per_step": int(nsus), "cta_threads": int(blk),
                       "ms_per_step_per_rank": per_rank_ms, "straggler_rank": int(np.argmax(per_rank_ms)),
                       "records_checksum": checksum},
            "e2e": {"value": tot_succ * K / (t_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": t_e2e / K},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": kc.get(name, {}).get("dram_bytes_per_launch"),
                         "peak_source": f"{peak_kind} (MEASURED_PEAKS.json hbm_gbs)" if peak_kind == "measured" else "fallback 6.65 TB/s",
                         "kernel": "k_plan", "model": "sum((12*N_obs+63.2)*successors) per launch (SURVEY 8d)",
                         "fp64_pipe_pct": kc.get(name, {}).get("fp64_pipe_pct"), "issue_slots_pct": kc.get(name, {}).get("issue_slots_pct"),
                         "icc_hit_pct": kc.get(name, {}).get("icc_hit_pct"), "ipc_per_active_sm": kc.get(name, {}).get("ipc_per_active_sm"),
                         "counters_from": kc.get(name, {}).get("source"),
                         "reading": "the kernel is latency and instruction-fetch bound (strictly ordered pops of bit-exact fp64 libm chains from ~117 KB of "
                                    "hot code against a 32 KB instruction cache: icc hit rate 64 %, 0.7 instructions per cycle on a busy SM), not "
                                    "HBM bound: measured DRAM traffic is far below the algorithmic bytes of the model"},
        }
        if clk is not None:
            out["clocks"] = clk
        if world == 1 and cpu_baseline:
            from automatedvaletparking_b200.hostcfg import make_avp_config
            cores = os.cpu_count() or 1
            sample, what = cpu_sample(name, scs, cores)
            csucc, cdt, _ = oracle_run(sample, make_avp_config(max_pops=20000), cores)
            out["cpu_baseline"] = {"value": csucc / cdt, "unit": UNIT, "cores": cores, "kind": "port", "per_thread": csucc / cdt / cores,
                                   "sample": f"{what}, {cores} threads, oracle/avp_oracle.c ({cdt:.1f} s)"}
    return out


def latency_case1(dp_device: int):
    """BASELINE configs[0] through the drop-in call surface: Map + PathPlanner.path_planning() of Case1, wall-clock ms (median of 5
    after one warm-up), host buffers in, Python lists out -- the call a user of the reference makes."""
    import tempfile
    from automatedvaletparking_b200 import scenarios as scn
    from automatedvaletparking_b200.map import costmap
    from automatedvaletparking_b200.path_plan import path_planner
    from automatedvaletparking_b200.config import read_config
    cfg = read_config.read_config("config")
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "Case1.csv")
        scn.write_case_csv(scn.benchmark_case(1), p)
        t_map, t_plan = [], []
        for it in range(6):
            t0 = time.perf_counter()
            m = costmap.Map(file=p, discrete_size=cfg["map_discrete_size"], device=dp_device)
            t1 = time.perf_counter()
            planner = path_planner.PathPlanner(config=cfg, map=m, vehicle=costmap.Vehicle())
            out, info, split = planner.path_planning()
            t2 = time.perf_counter()
            if it:
                t_map.append(1e3 * (t1 - t0)); t_plan.append(1e3 * (t2 - t1))
            m._device.close()
    return {"workload": "BenchmarkCases/Case1.csv through costmap.Map + PathPlanner.path_planning() (BASELINE configs[0])",
            "map_ms": statistics.median(t_map), "path_planning_ms": statistics.median(t_plan), "path_points": len(out),
            "segments": len(split), "note": "wall clock incl. context creation per Map; the Python reference needs ~38 s for this call (SURVEY 6)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=["c2", "c3", "c4", "c4s"])
    ap.add_argument("--scenarios", type=int, default=N_SCEN, help="c2 only: scenarios per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the other workloads and the Case1 latency")
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
    from automatedvaletparking_b200.batch import DevicePlanner

    use_dist = world > 1
    if use_dist:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")

    os.environ.setdefault("AVP_HOST_TIMEOUT_S", "600")
    dp = DevicePlanner(device=local_rank, max_pops=20000)
    W = max(3, args.warmup)
    K = args.steps
    head = run_workload(args.workload, dp, rank, world, local_rank, W, K, args.scenarios, True, not args.no_cpu_baseline)
    extra = {}
    if not args.no_extra:
        for name in ("c2", "c3", "c4"):
            if name != args.workload:
                r = run_workload(name, dp, rank, world, local_rank, 3, max(1, min(K, 2)), args.scenarios, False, False)
                if r is not None:
                    extra[name] = {"metric": METRIC, "unit": UNIT, "steps": max(1, min(K, 2)), "warmup": 3, **r}
    dp.close()
    if rank == 0:
        line = {"metric": METRIC, "value": head["value"], "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": head["scaling"], "vs_baseline": None,
                "dtype": "f64", "data": "synthetic", "config": head["config"], "e2e": head["e2e"], "gpu_launches": head["gpu_launches"],
                "clocks": head.get("clocks"), "roofline": head["roofline"]}
        if "cpu_baseline" in head:
            line["cpu_baseline"] = head["cpu_baseline"]
        if extra:
            line["workloads"] = extra
        if not args.no_extra and world == 1:
            try:
                line["case1_dropin_latency"] = latency_case1(local_rank)
            except Exception as e:                       # never lose the headline over the latency probe
                line["case1_dropin_latency"] = {"error": repr(e)}
        print(json.dumps(line), flush=True)
    if use_dist:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
