"""GPU parity tests (run on the B200 box with `-m gpu`): the CUDA path, called through the C ABI
(include/avp_b200.h via ctypes), against the CPU oracle on the same seeded inputs and against
the committed traces of the UNMODIFIED reference (tests/golden/).

Bars (BASELINE.json north_star): expanded-node indices, visited grid cells, raster cells,
collision booleans: bit-exact.  Path states (x, y, theta): <= 1e-4; in fact every fp64 value
compared here is required to be bit-identical, because the device evaluates the same IEEE
operation sequence as the host libm the reference runs on (csrc/avp_sincos.h, avp_libm.h)."""
import glob
import os

import numpy as np
import pytest

import oracle_lib as O
from automatedvaletparking_b200 import scenarios as scn
from conftest import GOLDEN

pytestmark = pytest.mark.gpu

POSE_TOL = 1e-4     # north_star tolerance on (x, y, theta); the tests also assert bit-equality


def _oracle_cells(m):
    ix, iy = np.where(m.cost_map() == 255)
    return np.stack([ix, iy], 1)


def test_rasters_bit_exact_all_benchmark_cases(device_planner):
    dp = device_planner
    scs = [scn.benchmark_case(i) for i in range(1, 21)]
    dp.load(scs)
    for k, sc in enumerate(scs):
        m = O.OracleMap(sc)
        info = dp.map_info(k)
        assert (info["nx"], info["ny"], info["n_obs"]) == (m.nx, m.ny, m.n_obs), sc.name
        assert info["dx"] == m.dx and info["dy"] == m.dy and np.array_equal(info["boundary"], m.boundary)
        assert np.array_equal(dp.cost_map(k), m.cost_map()), sc.name
        g = os.path.join(GOLDEN, "cases", f"{sc.name}.npz")
        if os.path.exists(g):        # the reference's own raster
            ref = np.load(g)
            ix, iy = np.where(dp.cost_map(k) == 255)
            assert np.array_equal(np.stack([ix, iy], 1).astype(np.uint16), ref["obs_cells"])


def test_rasters_of_perturbed_and_synthetic_scenarios_match_reference(device_planner):
    dp = device_planner
    g = np.load(os.path.join(GOLDEN, "leaf_maps.npz"))
    per, syn = [], []
    for i in range(int(g["n"])):
        s = scn.parse_case_row(g[f"m{i}_row"])
        b = g[f"m{i}_boundary"]
        if not np.isnan(b[0]):
            s.boundary = tuple(b)
            syn.append((i, s))
        else:
            per.append((i, s))
    for group in (per, syn):
        dp.load([s for _, s in group])
        for k, (i, s) in enumerate(group):
            ix, iy = np.where(dp.cost_map(k) == 255)
            assert np.array_equal(np.stack([ix, iy], 1).astype(np.uint16), g[f"m{i}_cells"])
            info = dp.map_info(k)
            assert (info["dx"], info["dy"]) == (g[f"m{i}_geom"][6], g[f"m{i}_geom"][7])


def test_collision_checker_matches_reference_vectors(device_planner, native_built):
    from automatedvaletparking_b200.batch import DevicePlanner
    dp = device_planner
    g = np.load(os.path.join(GOLDEN, "leaf_collision.npz"))
    cases = (1, 5, 13, 19)
    dp.load([scn.benchmark_case(c) for c in cases])
    for k, c in enumerate(cases):
        got = dp.check(k, g[f"c{c}_poses"])
        assert np.array_equal(got, g[f"c{c}_distance"]), f"Case{c}"
    cfg = dict(__import__("automatedvaletparking_b200.hostcfg", fromlist=["x"]).default_config())
    cfg["collision_check"] = "circle"
    with DevicePlanner(cfg) as dc:
        dc.load([scn.benchmark_case(c) for c in cases])
        for k, c in enumerate(cases):
            got = dc.check(k, g[f"c{c}_poses"][:400])
            assert np.array_equal(got, g[f"c{c}_circle"]), f"circle Case{c}"


def test_collision_checker_axis_aligned_and_edge_poses_match_oracle(device_planner, cfg):
    """theta in {0, +-pi/2, +-pi}: the reference's slope form degenerates to inf/nan."""
    dp = device_planner
    sc = scn.benchmark_case(4)
    dp.load([sc])
    m = O.OracleMap(sc)
    rng = np.random.default_rng(3)
    P = np.stack([rng.uniform(m.boundary[0] + 3, m.boundary[1] - 3, 800), rng.uniform(m.boundary[2] + 3, m.boundary[3] - 3, 800),
                  rng.choice([0.0, np.pi / 2, -np.pi / 2, np.pi, -np.pi, 1e-300, np.pi / 4], 800)], 1)
    assert np.array_equal(dp.check(0, P), np.array([m.check(cfg, *p) for p in P]))
    far = np.array([[m.boundary[0] - 50.0, m.boundary[2] - 50.0, 0.3], [m.boundary[1] + 50.0, 0.0, 1.0]])
    assert not dp.check(0, far).any()          # empty AABB intersection


def test_expand_pure_bit_exact(device_planner, cfg):
    """successor poses (glibc-exact sin/cos), sub-step collision flags, rs lengths."""
    dp = device_planner
    cases = (1, 4, 13, 16)
    scs = [scn.benchmark_case(c) for c in cases]
    dp.load(scs)
    rng = np.random.default_rng(5)
    for k, sc in enumerate(scs):
        m = O.OracleMap(sc)
        for t in range(40):
            par = np.array([sc.x0 + rng.uniform(-3, 3), sc.y0 + rng.uniform(-3, 3), rng.uniform(-np.pi, np.pi)])
            gp, gf, gl = dp.expand_pure(k, par)
            op, of, ol = m.expand_pure(cfg, par)
            assert np.array_equal(gp, op) and np.array_equal(gf, of)
            assert np.array_equal(gl, ol, equal_nan=True)


def test_rs_optimal_matches_reference_vectors(device_planner):
    dp = device_planner
    g = np.load(os.path.join(GOLDEN, "leaf_rs.npz"))
    q, maxc = g["q"], float(g["maxc"])
    n = 2000
    r = dp.rs_optimal(q[:n], maxc, 0.5, 1, 1, cap_pts=256)
    off = np.concatenate([[0], np.cumsum(g["npts"])])
    for i in range(n):
        assert r["ctypes"][i] == str(g["sel_ct"][i])
        assert r["L"][i] == g["sel_L"][i]
        assert np.array_equal(r["lengths"][i][:r["nseg"][i]], g["sel_len"][i][:g["sel_n"][i]])
        npt = int(g["npts"][i])
        assert r["n_pts"][i] == npt
        assert np.array_equal(r["x"][i, :npt], g["cx"][off[i]:off[i] + npt])
        assert np.array_equal(r["y"][i, :npt], g["cy"][off[i]:off[i] + npt])
        assert np.array_equal(r["yaw"][i, :npt], g["cyaw"][off[i]:off[i] + npt])
        assert np.array_equal(r["directions"][i, :npt], g["cdir"][off[i]:off[i] + npt])
    rr = dp.rs_optimal(q[:500], maxc, 0.5, 1, 0, cap_pts=256)      # root-node typing
    assert list(rr["ctypes"]) == [str(c) for c in g["root_ct"]] and np.array_equal(rr["L"], g["root_L"])


def _compare_plan(dp, res, k, sc, cfg, ref=None):
    m = O.OracleMap(sc)
    r = O.plan(m, cfg) if ref is None else ref
    s = res.summaries[k]
    for key in ("status", "n_pops", "global_index", "n_closed", "n_open", "n_astar", "n_rs", "n_final", "n_hq", "h_closed", "n_hcalls"):
        assert int(s[key]) == r[key], (sc.name, key, int(s[key]), r[key])
    gp = res.pop_indices(k)                                                        # expanded-node indices (n_pops itself is compared above;
    assert np.array_equal(gp, r["pops"][:len(gp)]) and len(gp) == min(r["n_pops"], res.pops.shape[1]), sc.name   # the trace is capped at cap_pops)
    hq = dp.hq_log(k, int(s["n_hq"]))
    assert np.array_equal(hq, r["hq"][:len(hq)].astype(np.int32)), sc.name         # Dijkstra query trace
    if r["status"] in (0, 2):
        fp = res.path(k)
        assert fp.shape == r["final_path"].shape
        assert np.abs(fp - r["final_path"]).max() <= POSE_TOL
        assert np.array_equal(fp, r["final_path"]), sc.name
        assert s["rs_ctypes"].decode() == r["rs_ctypes"] and s["rs_L"] == r["rs_L"]
    return r


def test_plan_all_benchmark_cases_vs_oracle_and_reference(device_planner, cfg):
    dp = device_planner
    scs = [scn.benchmark_case(i) for i in range(1, 21)]
    dp.load(scs)
    res = dp.plan(cap_path=512, cap_pops=20000)
    for k, sc in enumerate(scs):
        r = _compare_plan(dp, res, k, sc, cfg)
        gpath = os.path.join(GOLDEN, "cases", f"{sc.name}.npz")
        if os.path.exists(gpath):                 # the reference's own trace
            g = np.load(gpath)
            n = len(g["pops"])
            assert np.array_equal(res.pop_indices(k)[:n], g["pops"]), sc.name
            if str(g["status"]) == "ok":
                final = np.concatenate([g["astar_path"], np.stack([g["rs_x"], g["rs_y"], g["rs_yaw"]], 1)[1:]], 0)
                assert np.abs(res.path(k) - final).max() <= POSE_TOL and np.array_equal(res.path(k), final)
            if str(g["status"]) == "open_exhausted":
                assert int(res.summaries["status"][k]) == 1          # reference: AttributeError (Case20)
    # statuses of the three cases the reference cannot finish: capacity, same prefix as the oracle
    for k in (6, 7, 18):
        assert int(res.summaries["status"][k]) == 5 and int(res.summaries["n_pops"][k]) == 20000


def test_visited_grid_cells_bit_exact(device_planner, cfg):
    """The Dijkstra h table (distance of the first closedlist entry per grid id, compute_h.py:80)
    after the whole search: same visited cells, same distances, incl. Case4's history-dependent
    holes."""
    dp = device_planner
    cases = (1, 4, 16, 5)
    scs = [scn.benchmark_case(c) for c in cases]
    dp.load(scs)
    res = dp.plan(cap_path=512, cap_pops=4096)
    for k, sc in enumerate(scs):
        m = O.OracleMap(sc)
        r = O.plan(m, cfg, want_hval=True)
        hv = dp.hvalues(k)
        ref = r["hval"][:len(hv)]
        assert np.array_equal(hv, ref), sc.name
        assert int((hv >= 0).sum()) > 1000
        assert int(res.summaries["h_closed"][k]) == r["h_closed"]


def test_perturbed_batch_vs_oracle(device_planner, cfg):
    """BASELINE config 2/3 recipe at a size the oracle finishes in seconds."""
    dp = device_planner
    # recipe with the collision-free start/goal rule, checked on the GPU and cross-checked on the oracle
    cands = scn.perturbed_candidates(scn.benchmark_case(1), 200, seed=1)
    dp.load(cands)
    a, b = dp.start_goal_collisions()
    from automatedvaletparking_b200.batch import _pi_2_pi
    for k in range(0, 200, 7):
        m = O.OracleMap(cands[k])
        assert a[k] == m.check(cfg, cands[k].x0, cands[k].y0, _pi_2_pi(cands[k].theta0))
        assert b[k] == m.check(cfg, cands[k].xf, cands[k].yf, _pi_2_pi(cands[k].thetaf))
    scs = scn.keep_collision_free(cands, a, b, 40)
    # plus unfiltered perturbations of other cases (colliding starts give status 2 / 1: status parity)
    for c in (4, 9, 13, 16, 18, 20):
        scs += scn.perturbed_set(scn.benchmark_case(c), 4, seed=100 + c)
    dp.load(scs)
    res = dp.plan(cap_path=512, cap_pops=20000)
    n_ok = 0
    for k, sc in enumerate(scs):
        r = _compare_plan(dp, res, k, sc, cfg)
        n_ok += r["status"] == 0
    assert n_ok >= 30


def test_synthetic_stress_maps_vs_oracle(device_planner, cfg):
    """BASELINE config 4 recipe (200x200 raster, 256 polygons) incl. failing scenarios (status parity)."""
    dp = device_planner
    scs = scn.synthetic_set(2, 6, seed=4)
    dp.load(scs)
    res = dp.plan(cap_path=512, cap_pops=20000)
    for k, sc in enumerate(scs):
        _compare_plan(dp, res, k, sc, cfg)


def test_full_size_batch_properties(device_planner, cfg):
    """BASELINE config 2 at full size (1024 scenarios): size-independent properties."""
    dp = device_planner
    base = scn.benchmark_case(1)
    cands = scn.perturbed_candidates(base, 8 * 1024, seed=1)
    dp.load(cands)
    a, b = dp.start_goal_collisions()
    scs = scn.keep_collision_free(cands, a, b, 1024)
    dp.load(scs)
    res = dp.plan(cap_path=512, cap_pops=0)
    s = res.summaries
    assert (s["n_obs"] > 500).all() and (s["nx"] >= 240).all()
    ok = s["status"] == 0
    assert ok.mean() > 0.5
    # expansions: global_index = 10 * (pops - 1) on success (the last pop is the goal shot)
    assert (s["global_index"][ok] == 10 * (s["n_pops"][ok] - 1)).all()
    assert (s["n_final"][ok] == s["n_astar"][ok] + s["n_rs"][ok] - 1).all()
    assert ((s["n_astar"][ok] - 1) % 3 == 0).all()
    poses = np.array([sc.pose for sc in scs])
    idx = np.where(ok)[0]
    for k in idx:
        p = res.path(k)
        assert p[0, 0] == poses[k, 0] and p[0, 1] == poses[k, 1]                       # starts at the start pose
        assert abs(p[-1, 0] - poses[k, 3]) < 1e-6 and abs(p[-1, 1] - poses[k, 4]) < 1e-6   # rs shot ends at the goal
    # idempotence: planning the same batch again gives identical records
    res2 = dp.plan(cap_path=512, cap_pops=0)
    assert np.array_equal(res2.summaries, s)
    # a sample of the batch against the oracle, incl. collision-freeness of the returned path
    for k in list(idx[:12]) + list(np.where(~ok)[0][:4]):
        m = O.OracleMap(scs[k])
        r = O.plan(m, cfg)
        assert int(s["status"][k]) == r["status"] and int(s["n_pops"][k]) == r["n_pops"]
        if r["status"] == 0:
            assert np.array_equal(res.path(k), r["final_path"])
            assert not any(m.check(cfg, *q) for q in res.path(k)[::3])


def test_edge_inputs(device_planner, cfg):
    """single scenario, scenario without obstacles, start == colliding pose, tiny capacity."""
    from automatedvaletparking_b200.batch import DevicePlanner
    dp = device_planner
    empty = scn.Scenario(0.0, 0.0, 0.1, 6.0, 3.0, 0.4, [], None, "empty")
    dp.load([empty])
    res = dp.plan(cap_path=256, cap_pops=64)
    _compare_plan(dp, res, 0, empty, cfg)
    assert int(res.summaries["n_obs"][0]) == 0
    with DevicePlanner(max_pops=7) as small:
        small.load([scn.benchmark_case(1)])
        r = small.plan(cap_path=256, cap_pops=16)
        ref = O.plan(O.OracleMap(scn.benchmark_case(1)), small.cfg)
        assert int(r.summaries["status"][0]) == ref["status"] == 5 and np.array_equal(r.pop_indices(0), ref["pops"])


def _oracle_many(scs, cfg, threads=None):
    """the oracle on a thread pool (ctypes releases the GIL); one summary dict per scenario"""
    from concurrent.futures import ThreadPoolExecutor
    O.lib()

    def one(sc):
        r = O.plan(O.OracleMap(sc), cfg, cap_pops=1, cap_path=512)
        return {k: r[k] for k in ("status", "n_pops", "global_index", "n_closed", "n_open", "n_astar", "n_rs", "n_final", "n_hq", "h_closed", "n_hcalls",
                                  "final_path", "rs_L", "rs_ctypes")}

    with ThreadPoolExecutor(max_workers=threads or (os.cpu_count() or 4)) as ex:
        return list(ex.map(one, scs))


def _compare_many(res, scs, refs):
    s = res.summaries
    for k, (sc, r) in enumerate(zip(scs, refs)):
        for key in ("status", "n_pops", "global_index", "n_closed", "n_open", "n_astar", "n_rs", "n_final", "n_hq", "h_closed", "n_hcalls"):
            assert int(s[key][k]) == r[key], (sc.name, key, int(s[key][k]), r[key])
        if r["status"] in (0, 2):
            assert np.array_equal(res.path(k), r["final_path"]), sc.name
            assert s["rs_ctypes"][k].decode() == r["rs_ctypes"] and s["rs_L"][k] == r["rs_L"], sc.name


def test_config2_full_batch_every_scenario_vs_oracle(device_planner, cfg):
    """BASELINE configs[1] at FULL size: all 1024 scenarios of the bench workload, every summary field and every
    returned path bit-identical to the oracle (the 51 searches that hit the 20 000-pop cap included)."""
    import bench
    dp = device_planner
    scs = bench.make_scenarios(0, 1024, dp)
    dp.load(scs)
    res = dp.plan(cap_path=512, cap_pops=0)
    refs = _oracle_many(scs, cfg)
    _compare_many(res, scs, refs)
    assert sum(r["status"] == 0 for r in refs) > 900 and sum(r["status"] == 5 for r in refs) > 10


def test_config3_all_cases_perturbed_vs_oracle(device_planner, cfg):
    """BASELINE configs[2] recipe: every BenchmarkCase x 16 perturbed start/goal poses (seed 100 + case), unfiltered
    (colliding starts are status parity), all against the oracle."""
    dp = device_planner
    scs = []
    for c in range(1, 21):
        scs += scn.perturbed_set(scn.benchmark_case(c), 16, seed=100 + c)
    dp.load(scs)
    res = dp.plan(cap_path=512, cap_pops=0)
    _compare_many(res, scs, _oracle_many(scs, cfg))


def test_config4_synthetic_batch_vs_oracle(device_planner, cfg):
    """BASELINE configs[3] recipe: 16 synthetic 200x200 maps (256 polygons each) x 32 start/goal pairs."""
    dp = device_planner
    scs = scn.synthetic_set(16, 32, seed=4)
    dp.load(scs)
    res = dp.plan(cap_path=512, cap_pops=0)
    assert (res.summaries["nx"] == 200).all() and (res.summaries["n_obs"] > 1000).all()
    _compare_many(res, scs, _oracle_many(scs, cfg))


def test_popped_nodes_f_g_h_bit_exact(device_planner, cfg):
    """calc_node_cost / calc_node_heuristic (hybrid_a_star.py:243-298) checked directly: f, g, h of every popped node as they
    are at open_list.get(), against the oracle and against the reference's own traces (tests/golden/cases pop_fgh)."""
    dp = device_planner
    cases = (1, 4, 5, 13, 16, 17)
    scs = [scn.benchmark_case(c) for c in cases] + scn.perturbed_set(scn.benchmark_case(9), 3, seed=109)
    dp.load(scs)
    dp.trace_fgh(True)
    try:
        res = dp.plan(cap_path=512, cap_pops=4096)
        fgh = dp.pop_fgh(4096)
    finally:
        dp.trace_fgh(False)
    for k, sc in enumerate(scs):
        r = O.plan(O.OracleMap(sc), cfg)
        n = min(int(res.summaries["n_pops"][k]), 4096)
        assert n == min(r["n_pops"], 4096) and np.array_equal(fgh[k, :n], r["pop_fgh"][:n]), sc.name
        g = os.path.join(GOLDEN, "cases", f"{sc.name}.npz")
        if os.path.exists(g):
            ref = np.load(g)["pop_fgh"]
            m = min(len(ref), n)
            assert np.array_equal(fgh[k, :m], ref[:m]), sc.name


def test_config3_full_size_vs_oracle(device_planner, cfg):
    """BASELINE configs[2] at FULL size: 20 cases x 256 collision-free perturbations = 5120 scenarios (the bench recipe), every
    summary field and every returned path against the oracle."""
    import bench
    dp = device_planner
    scs = bench.make_c3(dp)
    assert len(scs) == 5120
    dp.load(scs)
    res = dp.plan(cap_path=512, cap_pops=0)
    _compare_many(res, scs, _oracle_many(scs, cfg))
    assert int((res.summaries["status"] == 5).sum()) >= 3 * 200          # Cases 7, 8, 19 never finish


def test_config4_full_size_vs_oracle(device_planner, cfg):
    """BASELINE configs[3] at FULL size: 64 synthetic maps x 64 start/goal pairs = 4096 scenarios (status parity: at this obstacle
    density every search ends at its first pop)."""
    import bench
    dp = device_planner
    scs = bench.make_c4(dp)
    assert len(scs) == 4096
    dp.load(scs)
    res = dp.plan(cap_path=512, cap_pops=0)
    _compare_many(res, scs, _oracle_many(scs, cfg))


def test_root_expansion_boundary_rule_matches_oracle(device_planner, cfg):
    """closed_list is re-read for every successor (hybrid_a_star.py:155-163): in the ROOT expansion an earlier sibling that collides
    switches the boundary test on for the later ones.  Tight boundary overrides around the start pose make that visible."""
    dp = device_planner
    base = scn.benchmark_case(1)
    rng = np.random.default_rng(11)
    scs = []
    for i in range(96):
        x0, y0 = base.x0 + rng.uniform(-1, 1), base.y0 + rng.uniform(-1, 1)
        pad = rng.uniform(0.2, 2.5, size=4)
        b = (float(np.floor(min(x0, base.xf) - 12)), float(x0 + pad[1]), float(np.floor(min(y0, base.yf) - 12)), float(y0 + pad[3]))
        scs.append(scn.Scenario(float(x0), float(y0), float(rng.uniform(-np.pi, np.pi)), base.xf, base.yf, base.thetaf, base.obs, b, f"tight{i}"))
    dp.load(scs)
    res = dp.plan(cap_path=512, cap_pops=20000)
    n_diff = 0
    for k, sc in enumerate(scs):
        r = _compare_plan(dp, res, k, sc, cfg)
        n_diff += r["n_closed"] > 0
    assert n_diff > 10


@pytest.mark.parametrize("env", [
    {"AVP_QUANTUM": "7", "AVP_FORCE_YIELD": "1"},                                   # every search is let go of every 7 pops and resumed (often on another SM)
    {"AVP_QUANTUM": "5", "AVP_FORCE_YIELD": "1", "AVP_SLOTS": "3"},                 # ... with a slot pool that runs dry: fresh scenarios go to the back of the queue
    {"AVP_QUANTUM": "16", "AVP_FORCE_YIELD": "1", "AVP_PLAN_BLOCK": "256"},         # two 256-thread CTAs per SM
    {"AVP_QUANTUM": "64", "AVP_SPREAD": "0"},
    {"AVP_TWO_PHASE": "1", "AVP_NARROW_BUDGET": "5"},                               # the narrow first launch (64-thread CTAs, 5 pops each), then the wide one
    {"AVP_TWO_PHASE": "1", "AVP_NARROW_BUDGET": "40", "AVP_NARROW_BLOCK": "128"},   # ... with 128-thread narrow CTAs (word candidates in shared memory)
    {"AVP_TWO_PHASE": "1", "AVP_NARROW_BUDGET": "9", "AVP_SLOTS": "4", "AVP_QUANTUM": "33", "AVP_FORCE_YIELD": "1"},
    {"AVP_TWO_PHASE": "0"},
])
def test_results_do_not_depend_on_the_scheduling(device_planner, cfg, monkeypatch, env):
    """Suspend / resume (run queue, slot pool, save areas of the heap heads), CTA width and SM-pair placement are scheduling only:
    pops, traces and paths of the benchmark cases and of a perturbed batch stay bit-identical to the oracle."""
    dp = device_planner
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    scs = [scn.benchmark_case(i) for i in (1, 2, 4, 5, 9, 13, 16, 17, 20)] + scn.perturbed_set(scn.benchmark_case(3), 6, seed=103)
    dp.load(scs)
    res = dp.plan(cap_path=512, cap_pops=8192)
    _, _, n_suspend, block = dp.last_search_passes()
    for k, sc in enumerate(scs):
        _compare_plan(dp, res, k, sc, cfg)
    if "AVP_FORCE_YIELD" in env or env.get("AVP_TWO_PHASE") == "1":
        assert n_suspend > (100 if "AVP_FORCE_YIELD" in env else 5)
    assert block == int(env.get("AVP_PLAN_BLOCK", 512))
