// avp_search_pipe.cuh -- the pipelined search kernel for the long tail (pass 2).
//
// PathPlanner.a_star_plan (path_planner.py:58-110) has two kinds of work per popped node:
//
//   PURE    a function of the node's pose only: the 10 successor poses (hybrid_a_star.py:134-151), their
//           sub-step collision checks (:185-204), their rs lengths (:286-292) and the goal shot of the node
//           itself (try_rs_curve, :318-349).
//   COMMIT  order dependent: closed/open lookups (:154-172), node creation, the Dijkstra term of the
//           heuristic (history dependent, compute_h.py:198-214), heap pushes / in-place updates (:206-230)
//           and the next open_list.get() (path_planner.py:70).
//
// k_search evaluates both one after the other for every pop.  Here warp 0 (the COMMIT warp) runs the
// reference's sequential loop while warps 1.. (the EVALUATORS) compute the PURE part of the node that
// will be popped next, one step ahead.  The next node is predicted after the lookups of the current
// commit: it is the first successor with the smallest f if that f is below the f of the heap root, else
// the heap root -- exactly what the pushes of this commit produce, unless a successor's Dijkstra value
// is not in the table yet (then the prediction may miss).  A miss costs one un-overlapped evaluation;
// results never depend on the prediction (a PURE result is only used for the node it was computed for).
//
//   barrier A | C: accept the result of the popped node; lookups (lanes 0..9, probes read ahead by    | barrier B
//             |    the evaluators); predict the next pop; publish it as the evaluators' target and       |
//             |    initialise their queue                                                             |
//   barrier B | C: node records, sequential commit in slot order (Dijkstra resumes inside), heappop,  | barrier A
//             |    then it takes E1 items if any are left                                             |
//             | E: ONE dependency-ordered queue of warp items (see E1 / E2 below)                     |
//
// The evaluation is a single queue of warp items without a barrier inside: items whose inputs come from
// other items (rs words <- successor poses, sub-step checks <- sub-step poses, course point checks <-
// course plan, word selection <- all rs items, table probes <- the commit warp's inserts) sit behind their
// producers in the queue and wait on a shared-memory flag / counter (release / acquire at CTA scope).  A
// producer never waits, so the queue cannot dead-lock.  A and B are __syncthreads(), the only barriers
// of a pop.
// The word of the goal shot is the word selected when the node was scored (NodeShot), so the shot costs
// no rs solve; its course is planned by the first queue item and its points are collision-checked by
// whichever warps run out of rs / sub-step items first.
// The exact-pose table probes and h-table reads of the NEXT commit's lookups are issued by the evaluators
// (item 3) as soon as the running commit has finished its table inserts: the DRAM round trips of the
// lookups leave the serial section between A and B (7-8 k -> 4.5 k cycles per pop).
#pragma once
#include "avp_kernels.cuh"

enum { CTL_FINISH = 2 };

// The rs warp items of the E1 queue: up to three word instances per item (lanes = instance slot x successor), formed so
// that a warp runs one word formula where possible (4 instances per family: 3 + 1 left over).  Ordered by measured cost,
// longest first (profiles/: 21 k ... 3 k cycles per item); a left-over item with three different formulas costs the sum of
// the three (41 k cycles for {9,29,37}), so those instances are items of their own.
#define RS_NITEM 19
__device__ __constant__ int8_t rs_item_inst[RS_NITEM][3] = {
  {45, 13, 17}, {0, 1, -1}, {6, 7, 8}, {5, 33, 41}, {20, 21, -1}, {22, 23, -1}, {10, 11, 12}, {14, 15, 16}, {26, 27, 28},
  {34, 35, 36}, {24, 25, -1}, {9, -1, -1}, {29, -1, -1}, {37, -1, -1}, {2, 3, 4}, {38, 39, 40}, {30, 31, 32}, {42, 43, 44}, {18, 19, -1}};

struct PureRes {
  double cpose[AVP_NCHILD_MAX][3];
  int32_t found[AVP_NCHILD_MAX], hv[AVP_NCHILD_MAX];     // exact-pose table probe and h-table value, read ahead by the evaluators (see E2)
  double rsL[AVP_NCHILD_MAX];
  NodeShot shot[AVP_NCHILD_MAX];
  int32_t coll[AVP_NCHILD_MAX], rsok[AVP_NCHILD_MAX], inrad[AVP_NCHILD_MAX], hid[AVP_NCHILD_MAX];
  int32_t node;                              // the node this result belongs to (-1: none)
  int32_t in_radius, shot_ok;                // the shot's collision / degeneracy flags stay in shared memory (s_shot_coll, s_shot_bad)
};
struct EvalTarget { double x, y, theta; NodeShot shot; int32_t node, in_radius, is_root, valid; };

// clock read that the compiler may not move across barriers or memory operations (profiling counters)
__device__ __forceinline__ long long clock_ordered() { long long t; asm volatile("mov.u64 %0, %%clock64;" : "=l"(t)::"memory"); return t; }


// CTA-scope release / acquire on shared-memory words: the hand-off between producers and consumers of the evaluators'
// queue (the payload is written with plain stores before the release and read with plain loads after the acquire).
__device__ __forceinline__ void st_release_cta(int *p, int v) { asm volatile("st.release.cta.shared::cta.s32 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(p)), "r"(v) : "memory"); }
__device__ __forceinline__ int ld_acquire_cta(const int *p) { int v; asm volatile("ld.acquire.cta.shared::cta.s32 %0, [%1];" : "=r"(v) : "r"((unsigned)__cvta_generic_to_shared(p)) : "memory"); return v; }
// wait until *p >= want.  A producer never waits (see the header), so this returns after a bounded time; the guard
// (2^27 cycles, three orders of magnitude above any real wait) turns a protocol bug into an error status instead of a hang.
__device__ __forceinline__ bool wait_ge_cta(const int *p, int want) {
  if (ld_acquire_cta(p) >= want) return true;
  const long long t0 = clock64();
  for (int k = 1;; ++k) {
    if (ld_acquire_cta(p) >= want) return true;
    if ((k & 255) == 0 && clock64() - t0 > (1ll << 27)) return false;
#ifdef AVP_SPIN_SLEEP        // A/B build: the polls are 8 % of the kernel's issued instructions (profiles/hot_footprint_r01d.txt)
    __nanosleep(AVP_SPIN_SLEEP);
#endif
  }
}
__device__ __forceinline__ void add_release_cta(int *p, int v) { asm volatile("red.release.cta.shared::cta.add.s32 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(p)), "r"(v) : "memory"); }

template <int BLOCK>
__device__ __forceinline__ void search_pipe_body(const KParams &P) {
  static_assert(BLOCK >= 128 && BLOCK % 32 == 0, "one commit warp + at least three evaluator warps");
  constexpr int SMO = avp_sm_open(BLOCK);
  extern __shared__ __align__(16) unsigned char s_dyn[];
  double *s_of = reinterpret_cast<double *>(s_dyn);
  int32_t *s_oi = reinterpret_cast<int32_t *>(s_dyn + sizeof(double) * SMO);
  __shared__ unsigned long long s_heap[AVP_SM_HEAP];
  __shared__ RsCand s_cand[AVP_NCHILD_MAX][RS_NINST];
  __shared__ unsigned long long s_valid[AVP_NCHILD_MAX];
  __shared__ RsQuery s_Q[AVP_NCHILD_MAX];
  __shared__ double s_sub[AVP_NCHILD_MAX][4][4];
  __shared__ RsGroupBest s_grp[AVP_NCHILD_MAX][RS_NGROUP];
  __shared__ double s_org[AVP_MAX_RS_SEG + 1][3];
  __shared__ PureRes s_res[2];
  __shared__ EvalTarget s_tgt;
  __shared__ RsBest s_best;
  __shared__ double s_tcs[2];
  __shared__ double s_g[AVP_NCHILD_MAX], s_oldf[AVP_NCHILD_MAX], s_h1[AVP_NCHILD_MAX];
  __shared__ int s_found[AVP_NCHILD_MAX], s_need[AVP_NCHILD_MAX], s_skip[AVP_NCHILD_MAX], s_hv[AVP_NCHILD_MAX];
  __shared__ int s_trace_on;
  __shared__ int s_chit[AVP_NCHILD_MAX];
  __shared__ int s_scen, s_ctlA, s_ctlB, s_cur, s_do_commit, s_rb, s_nplan, s_npts, s_shot_coll, s_shot_bad, s_work;
  __shared__ int s_G, s_nclosed, s_npops, s_status, s_nhq, s_nhcalls, s_on, s_in_radius, s_best_ok;
  __shared__ int s_work2, s_work3, s_rs_done, s_course_rdy, s_q_rdy, s_sub_rdy, s_ins_done, s_cstride;   // the evaluators' queue: tail counter, finished rs items, course / rs queries / sub-step poses published, table inserts of the running commit done, stride of the course point order
  __shared__ VehGeom s_vg[BLOCK / 32];           // per warp: the vehicle rectangle of the pose being checked (check_distance_warp_sm)
  __shared__ DijCtx s_D;
  __shared__ long long s_ic[48];              // cycles per queue item (0..39: E1 items, 40/41: selections / course checks, 42/43: their counts), development aid
  __shared__ long long s_wp[BLOCK / 32][8];   // per warp: cycles lane 0 spent working in each evaluator phase (barrier waits excluded)

  const avp_config &cfg = P.cfg;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int slot = (int)blockIdx.x;
  const int nchild = 2 * cfg.steering_angle_num;
  const double maxc = 1 / cfg.min_radius_turn;
  Node *nodes = P.nodes + (size_t)slot * P.node_cap;
  NodeShot *nshot = P.nshot + (size_t)slot * P.node_cap;
  int32_t *htab = P.htab + (size_t)slot * P.htab_stride;
  OEnt *oge = P.oheap + (size_t)slot * P.node_cap;
  const int hmask = P.htab_size - 1;
  double *CX = P.course + (size_t)slot * 3 * AVP_COURSE_CAP, *CY = CX + AVP_COURSE_CAP, *CYAW = CY + AVP_COURSE_CAP;
  int32_t *CDIR = P.course_dir + (size_t)slot * AVP_COURSE_CAP;

  // P.spread: the grid has one CTA per SM of the device although there is less work than SMs, and the CTAs on odd SM ids
  // leave the work to the even ones: each scenario then has an SM pair (TPC) to itself.  The kernel is instruction-fetch
  // bound and the pair shares fetch resources: a 20 000-pop scenario takes 1.43 G cycles with an idle neighbour SM, 1.9 G
  // beside another such scenario (profiles/).  An odd CTA only watches the work counter: it leaves when every item is taken,
  // and takes items itself if the counter has not moved for ~0.5 s (no even CTA resident, e.g. on a shared device).
  if (P.spread) {
    if (tid == 0) {
      unsigned sm_; asm("mov.u32 %0, %%smid;" : "=r"(sm_));
      int go = 1;
      if (sm_ & 1u) {
        go = 0;
        long long t_last = clock64(); int last = -1;
        for (;;) {
          const int c = *(volatile int *)P.work_counter;
          if (c >= P.n_work) break;
          if (c != last) { last = c; t_last = clock64(); }
          else if (clock64() - t_last > 1000000000ll) { go = 1; break; }
          __nanosleep(20000);
        }
      }
      s_scen = go;
    }
    __syncthreads();
    if (!s_scen) return;
    __syncthreads();
  }
  for (;;) {
    if (tid == 0) s_scen = atomicAdd(P.work_counter, 1);
    __syncthreads();
    const int scen_i = s_scen;
    if (scen_i >= P.n_work) break;
    const int sc = P.work_list ? P.work_list[scen_i] : scen_i;
    const ScenDev &S = P.scen[sc];
    const double2 *cells = P.cells + S.cell_off;
    const int32_t *col_start = P.col_start + S.col_off;
    int32_t *hval = P.hval + S.id_off, *ost = P.ost + S.id_off;
    const double goal[3] = {S.pose[3], S.pose[4], pi_2_pi(S.pose[5])};
    int32_t *pops = P.pops ? P.pops + (size_t)sc * P.cap_pops : nullptr;
    int32_t *hql = P.hq_log ? P.hq_log + (size_t)sc * AVP_HQ_CAP * 3 : nullptr;
    int *dbg = P.dbg ? P.dbg + (size_t)sc * 8 : nullptr;
    const long long t_start = clock64();
    // cycle accumulators: thread 0 (commit warp) and thread 32 (evaluators) each keep their own
    long long pc[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, tp = t_start;
#ifdef AVP_NO_PROFILE        // A/B build without the cycle counters in the hot loop (the profile outputs stay zero)
#define PIPE_TICK(who, k) do { } while (0)
#else
#define PIPE_TICK(who, k) do { if (tid == (who)) { const long long t_ = clock_ordered(); pc[k] += t_ - tp; tp = t_; } } while (0)
#endif
    long long wt = 0;
#ifdef AVP_NO_PROFILE
#define WP_START() do { } while (0)
#define WP_ACC(k) do { } while (0)
#else
#define WP_START() do { if (lane == 0) wt = clock_ordered(); } while (0)
#define WP_ACC(k) do { if (lane == 0) { const long long t_ = clock_ordered(); s_wp[warp][k] += t_ - wt; wt = t_; } } while (0)
#endif
    if (lane == 0) for (int k = 0; k < 8; ++k) s_wp[warp][k] = 0;
    if (tid < 48) s_ic[tid] = 0;
    // timeline of ONE pop (the AVP_TRACE_POP-th of the scenario): absolute clocks of every warp at the phase boundaries
    long long *tsw = (P.wprof && warp < 16) ? P.wprof + ((size_t)sc * 16 + warp) * 24 + 8 : nullptr;
#ifdef AVP_NO_PROFILE
#define TS(k) do { } while (0)
#else
#define TS(k) do { __syncwarp(); if (lane == 0 && tsw && ((k) < 2 ? (s_npops == P.trace_pop) : s_trace_on)) tsw[k] = clock_ordered(); } while (0)
#endif

    // open_list.get() (path_planner.py:70) with the loop's exit tests; lane 0 of the commit warp
    auto do_pop = [&]() {
      if (dbg) { dbg[0] = 2; dbg[1] = s_npops; dbg[2] = s_D.closed_len; dbg[3] = s_on; }
      if (P.watchdog_cycles > 0 && clock64() - t_start > P.watchdog_cycles && s_status == 0) s_status = AVP_CAPACITY;
      if (s_status != 0 || s_on == 0) s_ctlA = CTL_EXIT;
      else if (s_npops >= cfg.max_pops) { s_status = AVP_CAPACITY; s_ctlA = CTL_EXIT; }
      else if (s_npops >= P.pop_budget) { s_status = AVP_PENDING; s_ctlA = CTL_EXIT; }
      else {
        const int ret = s_oi[0];
        s_cur = ret;
        if (pops && s_npops < P.cap_pops) pops[s_npops] = ret;
        s_npops++;
        int n_ = s_on; oh_pop_fix<SMO>(s_of, s_oi, oge, nodes, n_); s_on = n_;
        s_ctlA = CTL_RUN;
      }
    };

    // ---- per-scenario initialisation (all threads)
    for (int i = tid; i < S.n_ids; i += BLOCK) { hval[i] = -1; ost[i] = -1; }
    for (int i = tid; i < P.htab_size; i += BLOCK) htab[i] = -1;
    if (tid == 0) {
      s_D.S = &S; s_D.cost = P.cost + S.cost_off; s_D.hval = hval; s_D.ost = ost;
      s_D.gx = P.gx + S.id_off; s_D.gy = P.gy + S.id_off;
      s_D.sheap = s_heap; s_D.gheap = P.dheap + (size_t)slot * P.dheap_cap; s_D.gcap = P.dheap_cap;
      s_D.hn = 0; s_D.closed_len = 0; s_D.status = 0;
      s_on = 0;
      s_G = 0; s_nclosed = 0; s_npops = 0; s_nhq = 0; s_nhcalls = 0;
      s_status = S.raster_error ? AVP_RASTER_AMBIGUOUS : 0;
      s_cur = -1; s_in_radius = 0; s_best_ok = 0; s_shot_coll = 0; s_npts = 0; s_best.ok = 0;
      s_res[0].node = -1; s_res[1].node = -1; s_rb = 0; s_do_commit = 0; s_ctlA = CTL_RUN; s_ctlB = CTL_RUN;
      s_best.ct = 0; s_best.n = 0; s_best.inst = -1; s_best.degenerate = 0; s_best.L = 0.0;
      if (dbg) { dbg[0] = 1; dbg[1] = 0; }
    }
    __syncthreads();

    // ---- hybrid_a_star.__init__: eager Dijkstra to the start cell (hybrid_a_star.py:89-91), root node (:102-112);
    //      warp 1 meanwhile solves the root's rs word (the root's theta is a Python float: phi_np = 0)
    if (warp == 0 && s_status == 0) {
      long long term;
      const int d = dij_compute_path(s_D, s_heap, S.pose[0], S.pose[1], &term);
      if (lane == 0) {
        if (hql && s_nhq < AVP_HQ_CAP) { hql[3 * s_nhq] = (int)term; hql[3 * s_nhq + 1] = d; hql[3 * s_nhq + 2] = s_D.closed_len; }
        s_nhq++;
        if (d < 0) s_status = s_D.status ? s_D.status : AVP_H_UNREACHABLE;     // the reference never returns from this compute_path: no root node
        else {
        Node r; r.x = S.pose[0]; r.y = S.pose[1]; r.theta = pi_2_pi(S.pose[2]); r.f = 0; r.g = 0; r.h = 0; r.parent = -1;
        r.forward = 1; r.steer_idx = 0; r.in_open = 1; r.in_closed = 0; r.hpos = 0;
        r.in_radius = sqrt(d_pow2(r.x - goal[0]) + d_pow2(r.y - goal[1])) < cfg.flag_radius;
        nodes[0] = r;
        htab_insert(htab, hmask, nodes, 0);
        { int n_ = s_on; oh_push<SMO>(s_of, s_oi, oge, nodes, n_, 0.0, 0); s_on = n_; }
        }
      }
    } else if (warp == 1) {
      const double q0[3] = {S.pose[0], S.pose[1], pi_2_pi(S.pose[2])};
      RsBest b; rs_length_warp(q0, goal, maxc, 1, 0, s_cand[0], b);
      if (lane == 0) {
        NodeShot w; w.t = 0.0; w.u = 0.0; w.v = 0.0; w.L = 0.0; w.inst = -1; w.ok = 0;
        if (b.ok && !b.degenerate) { w.t = s_cand[0][b.inst].t; w.u = s_cand[0][b.inst].u; w.v = s_cand[0][b.inst].v; w.L = b.L; w.inst = b.inst; w.ok = 1; }
        nshot[0] = w;
      }
    }
    __syncthreads();
    if (tid == 0) do_pop();                      // the first get() returns the root
    PIPE_TICK(0, 0);                             // init + eager Dijkstra
    if (tid == 32) tp = clock_ordered();

    bool reached = false;
    for (;;) {
      __syncthreads();                                 // ---- barrier A: the evaluators' result is complete, the next node is popped
      PIPE_TICK(0, 4);                           // commit warp waiting for the evaluators
      TS(0);
      // s_ctlA is written by do_pop (between B and A) and read here; s_ctlB is written between A and B and read
      // after B: no control word is written while another warp may still be reading it
      if (s_ctlA == CTL_EXIT) break;
      if (warp == 0) {
        WP_START();
        const int wb = s_rb ^ 1;
        const int cur = s_cur;
        if (lane == 0) { s_ctlB = CTL_RUN; s_trace_on = (s_npops == P.trace_pop); }
        __syncwarp();
        if (s_res[wb].node == cur) {             // the result of the popped node is there
          PureRes &R = s_res[wb];
          const int shot_bad = R.in_radius ? s_shot_bad : 0, shot_coll = R.in_radius ? s_shot_coll : 0;   // of the evaluation just finished
          if (lane < nchild) R.coll[lane] = s_chit[lane];          // sub-step collision flags of the evaluation just finished
          if (lane == 0) { s_rb = wb; s_in_radius = R.in_radius; s_best_ok = R.shot_ok; pc[5]++; }
          __syncwarp();
          if (shot_bad) { if (lane == 0) { s_status = (shot_bad == 1) ? AVP_RS_DEGENERATE : AVP_CAPACITY; s_ctlB = CTL_EXIT; } }
          else if (R.in_radius && !shot_coll) { if (lane == 0) s_ctlB = CTL_FINISH; }     // path_planner.py:86-88
          else {
            // ---- lookups, g values, h-table prefetch (hybrid_a_star.py:154-172, :206-222): lanes 0..nchild-1.
            //      Node records and table inserts follow after barrier B (the evaluators do not need them).
            int found = -1, skip = 1, coll = 0;
            if (lane < nchild) {
              const int i = lane;
              const int id = R.hid[i];
              // probe and h value were read ahead by the evaluators (E2) after the table inserts of the commit that ran beside
              // them: no insert and no Dijkstra resume has happened since, except that an h value missing then may be there now
              int hvp = R.hv[i];
              if (hvp < 0 && id >= 0) hvp = hval[id];
              const Node cn = nodes[cur];
              const double x_ = R.cpose[i][0], y_ = R.cpose[i][1], th = R.cpose[i][2];
#ifdef AVP_NO_LOOKAHEAD
              found = htab_find(htab, hmask, nodes, x_, y_, th);
#else
              found = R.found[i];
#endif
              const bool in_closed = found >= 0 && nodes[found].in_closed;
              const bool oob = (s_nclosed > 0) && (x_ > S.b[1] || x_ < S.b[0] || y_ > S.b[3] || y_ < S.b[2]);
              skip = (in_closed || oob) ? 1 : 0;
              coll = R.coll[i];
              const int need = (!skip) && ((found < 0 && !coll) || (found >= 0));
              const bool fwd = i < nchild / 2.0;
              double g = 0.0;
              if (!skip) {
                if (found < 0) g = coll ? 0.0 : node_cost(cfg, fwd, th, cn.theta, cn.forward != 0);           // :206-209
                else { const Node &n = nodes[found]; g = node_cost(cfg, n.forward != 0, n.theta, cn.theta, cn.forward != 0); s_oldf[i] = n.f; }   // :219-222
              }
              if (need && !R.rsok[i]) s_status = AVP_RS_DEGENERATE;
              s_found[i] = found; s_skip[i] = skip; s_need[i] = need; s_g[i] = g;
              s_hv[i] = need ? hvp : -1;                                  // calc_node_heuristic (:261-283)
              s_h1[i] = s_hv[i] / 100.0;                                  // h_value_1 / 100 (:295)
            }
            __syncwarp();
            // ---- predict the next open_list.get(): the pushes of this commit put a successor at the root iff its f is
            //      below the root's; among successors the first one with the smallest f wins (heapq._siftdown is strict)
            double fc = INFINITY;
            if (lane < nchild && !skip && found < 0 && !coll && s_hv[lane] >= 0) {
              const double h2 = R.rsL[lane], h1 = s_h1[lane];
              fc = s_g[lane] + ((h2 > h1) ? h2 : h1);
            }
            int bi = (fc < INFINITY) ? lane : 64;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
              const double of_ = shfl_d(fc, lane ^ o); const int oi_ = __shfl_xor_sync(AVP_FULL_MASK, bi, o);
              if (of_ < fc || (of_ == fc && oi_ < bi)) { fc = of_; bi = oi_; }
            }
            if (lane == 0) {
              EvalTarget T; T.valid = 0; T.node = -1; T.is_root = 0; T.in_radius = 0; T.x = 0.0; T.y = 0.0; T.theta = 0.0;
              T.shot.t = 0.0; T.shot.u = 0.0; T.shot.v = 0.0; T.shot.L = 0.0; T.shot.inst = -1; T.shot.ok = 0;
              const int on = s_on;
              if (bi < nchild && (on == 0 || fc < s_of[0])) {
                T.valid = 1; T.node = s_G + bi + 1; T.x = R.cpose[bi][0]; T.y = R.cpose[bi][1]; T.theta = R.cpose[bi][2];
                T.in_radius = R.inrad[bi]; T.shot = R.shot[bi];
              } else if (on > 0) {
                const int id = s_oi[0]; const Node &n = nodes[id];
                T.valid = 1; T.node = id; T.x = n.x; T.y = n.y; T.theta = n.theta; T.in_radius = n.in_radius; T.is_root = (id == 0); T.shot = nshot[id];
              }
              s_tgt = T; s_do_commit = 1;
            }
          }
        } else if (lane == 0) {                  // prediction missed (or the very first pop): evaluate the popped node itself
          const Node &n = nodes[cur];
          EvalTarget T; T.valid = 1; T.node = cur; T.x = n.x; T.y = n.y; T.theta = n.theta; T.in_radius = n.in_radius; T.is_root = (cur == 0); T.shot = nshot[cur];
          s_tgt = T; s_do_commit = 0; pc[6]++;
        }
        // ---- the evaluators' queue for the target just published: counters, flags, header of the result buffer
        //      (the flags of the evaluation just finished were read above)
        __syncwarp();
        if (s_ctlB == CTL_RUN) {
          if (lane < nchild) { s_valid[lane] = 0ull; s_chit[lane] = 0; }
          if (lane == 0) {
            s_nplan = 0; s_npts = 0; s_cstride = 1; s_work = 0; s_work2 = 0; s_work3 = 0; s_rs_done = 0; s_course_rdy = 0; s_q_rdy = 0; s_sub_rdy = 0;
            s_ins_done = 0; s_shot_coll = 0; s_shot_bad = 0;
            PureRes &Wn = s_res[s_rb ^ 1];
            Wn.node = s_tgt.valid ? s_tgt.node : -1; Wn.in_radius = s_tgt.in_radius; Wn.shot_ok = (s_tgt.in_radius && s_tgt.shot.ok) ? 1 : 0;
          }
        }
      }
      if (warp == 0) WP_ACC(0);
      PIPE_TICK(0, 1);                           // accept + lookups + prediction
      TS(1);
      __syncthreads();                                 // ---- barrier B: target published
      TS(2);
      PIPE_TICK(32, 15);                         // evaluators waiting for the target
      if (s_ctlB != CTL_RUN) { reached = (s_ctlB == CTL_FINISH); break; }

      EvalTarget T;
      PureRes &W = s_res[s_rb ^ 1];                // s_rb only changes between A and B
      if (warp == 0) {
        // =========================== COMMIT warp ===========================
        WP_START();
        if (s_do_commit && s_status == 0) {
          const PureRes &R = s_res[s_rb];
          const int cur = s_cur;
          // node records of the new successors + exact-pose table inserts (hybrid_a_star.py:175-183): lanes 0..nchild-1
          if (lane < nchild && !s_skip[lane] && s_found[lane] < 0) {
            const int i = lane, child = s_G + i + 1;
            if (child >= P.node_cap) s_status = AVP_CAPACITY;
            else {
              const int coll = R.coll[i];
              Node n; n.x = R.cpose[i][0]; n.y = R.cpose[i][1]; n.theta = R.cpose[i][2]; n.parent = cur;
              n.g = s_g[i]; n.f = 0; n.h = 0;
              n.forward = (i < nchild / 2.0) ? 1 : 0; n.steer_idx = (uint8_t)(i % cfg.steering_angle_num); n.in_open = 0;
              n.in_closed = coll ? 1 : 0; n.hpos = -1;
              n.in_radius = coll ? 0 : R.inrad[i];
              nodes[child] = n;
              if (!coll) nshot[child] = R.shot[i];
              __threadfence_block();
              htab_insert(htab, hmask, nodes, child);
            }
          }
          __syncwarp();
        }
        if (lane == 0) { __threadfence_block(); st_release_cta(&s_ins_done, 1); }      // the evaluators may probe the table now
        if (s_do_commit && s_status == 0) {
          const PureRes &R = s_res[s_rb];
          const int cur = s_cur;
          // sequential commit in slot order (hybrid_a_star.py:154-239).  Lane 0 runs ahead over the successors
          // whose h value is already in the table; a miss resumes the Dijkstra search, which is warp-collective.
          int i = 0, n_miss = 0;
          int on = s_on;
          for (;;) {
            int stop = nchild;
            if (lane == 0) {
              for (; i < nchild; ++i) {
                if (s_skip[i]) continue;
                if (s_found[i] < 0 && R.coll[i]) { s_nclosed++; continue; }
                int hv = s_hv[i];
                double h1 = s_h1[i];
                if (n_miss > 0) {                    // a Dijkstra resume since the prefetch: re-read the table
                  const int id = R.hid[i];
                  hv = (id >= 0) ? hval[id] : -1;
                  h1 = hv / 100.0;
                }
                if (hv < 0) break;                   // miss: needs the warp
                s_nhcalls++;
                const double h2 = R.rsL[i];
                const double h = (h2 > h1) ? h2 : h1;                       // max(h_value_1, h_value_2) (:294-296)
                const int found = s_found[i];
                if (found < 0) {                                            // :206-216
                  const int child = s_G + i + 1;
                  Node &n = nodes[child];
                  const double f = s_g[i] + h;
                  n.h = h; n.f = f; n.in_open = 1;
#ifdef AVP_NO_PROFILE
                  oh_push<SMO>(s_of, s_oi, oge, nodes, on, f, child);
#else
                  { const long long t_ = clock64(); oh_push<SMO>(s_of, s_oi, oge, nodes, on, f, child); pc[8] += clock64() - t_; pc[9]++; }
#endif
                } else {                                                    // :219-230 (in place, no re-heapify)
                  const double new_f = h + s_g[i];
                  if (new_f < s_oldf[i]) {
                    Node &n = nodes[found];
                    n.f = new_f; n.g = s_g[i]; n.h = h; n.parent = cur; n.forward = (i < nchild / 2.0) ? 1 : 0; n.steer_idx = (uint8_t)(i % cfg.steering_angle_num);
                    oh_set_key<SMO>(s_of, oge, n.hpos, new_f);
                  }
                }
              }
              stop = i;
            }
            stop = __shfl_sync(AVP_FULL_MASK, stop, 0);
            if (stop >= nchild) break;
            // heuristic miss for successor `stop`: Dijkstra.compute_path resumes (compute_h.py:198-214)
            long long term;
            const long long td_ = clock64();
            const int d = dij_compute_path(s_D, s_heap, R.cpose[stop][0], R.cpose[stop][1], &term);
            if (lane == 0) { pc[10] += clock64() - td_; pc[11]++; }
            ++n_miss;
            if (lane == 0) {
              if (hql && s_nhq < AVP_HQ_CAP) { hql[3 * s_nhq] = (int)term; hql[3 * s_nhq + 1] = d; hql[3 * s_nhq + 2] = s_D.closed_len; }
              s_nhq++;
              if (d < 0) s_status = s_D.status ? s_D.status : AVP_H_UNREACHABLE;
            }
            __syncwarp();
            if (__shfl_sync(AVP_FULL_MASK, s_status, 0)) break;
          }
          if (lane == 0) s_on = on;
          if (lane == 0 && !s_status) { nodes[cur].in_closed = 1; nodes[cur].in_open = 0; s_nclosed++; s_G += nchild; }   // :235-239
        }
        __syncwarp();
        WP_ACC(1);
        TS(3);
        PIPE_TICK(0, 2);                         // node records + sequential commit
        if (lane == 0 && (s_do_commit || s_status != 0)) do_pop();
        WP_ACC(2);
        TS(4);
        PIPE_TICK(0, 3);                         // heappop
        // ---- the commit warp joins the evaluators' queue (initialised before barrier B)
        __syncwarp();
        T = s_tgt;
        if (!T.valid) continue;
        WP_START();
      } else {
        // =========================== EVALUATORS ===========================
        T = s_tgt;
        if (!T.valid) continue;
        WP_START();
        TS(3); TS(4);
      }
      // =========================== the evaluation queue (evaluators; the commit warp when it is done) ===========================
      // ---- E1: the producers and the items that only need the successor poses:
      //   0                      the shot's course from the node's stored word: plan (generate_local_course, rs_curve.py:537-594)
      //                          and points in the local frame (:597-624); publishes s_npts / s_cstride, then s_course_rdy
      //   1                      successor poses with their normalised rs queries (hybrid_a_star.py:145-151); publishes s_q_rdy
      //   2                      sub-step poses (:185-194); publishes s_sub_rdy
      //   3                      per successor (lanes): in_radius, cell id, and the read-ahead for the commit warp's lookups
      //                          (exact-pose table probe, h-table value), after s_q_rdy and s_ins_done
      //   4 .. 3+RS_NITEM        rs word instances (rs_item_inst: up to three instances x all successors per warp item), after
      //                          s_q_rdy; each finished item adds one to s_rs_done
      //   4+RS_NITEM ..          the successors' sub-step collision checks (:185-204), one successor each, after s_sub_rdy
      {
        const int phi_np = !T.is_root;             // the root's theta is a Python float (see oracle generate_path)
        const int nsubs = cfg.n_substeps;
        const int n_items = 4 + RS_NITEM + nchild;
        VehGeom *vg = &s_vg[warp];
        for (;;) {
          int it = 0;
          if (lane == 0) it = atomicAdd(&s_work, 1);
          it = __shfl_sync(AVP_FULL_MASK, it, 0);
          if (it >= n_items) break;
          const long long ti_ = clock64();
          if (it == 0) {
            int npts = 0;
            if (T.in_radius) {
              if (lane == 0) {
                RsBest b; b.ok = 0; b.degenerate = 0; b.n = 0; b.ct = 0; b.L = 0.0; b.inst = -1;
                if (!T.shot.ok) s_shot_bad = 1;
                else {
                  unsigned mask;
                  b.ok = 1; b.inst = T.shot.inst; b.L = T.shot.L;
                  b.n = rs_arrange(T.shot.inst, T.shot.t, T.shot.u, T.shot.v, 1, phi_np, b.len, b.ct, mask);
                  if ((int)(b.L / (0.5 * maxc)) + b.n + 3 > AVP_COURSE_CAP) s_shot_bad = 2;
                }
                s_best = b;
                s_tcs[0] = d_cos(-T.theta); s_tcs[1] = d_sin(-T.theta);
              }
              __syncwarp();
              if (!s_shot_bad) {
                const int nseg = s_best.n;
                const char *mode = rs_ct_names[s_best.ct];
                if (lane < nseg) {
                  double oyaw = 0.0;                                        // heading at the start of segment `lane`
                  for (int i = 0; i < lane; ++i) { if (mode[i] == 'L') oyaw = oyaw + s_best.len[i]; else if (mode[i] == 'R') oyaw = oyaw - s_best.len[i]; }
                  double ix, iy, yaw_next = oyaw; int dir;
                  rs_interpolate(s_best.len[lane], mode[lane], maxc, 0.0, 0.0, oyaw, ix, iy, yaw_next, dir);
                  s_org[lane + 1][0] = ix; s_org[lane + 1][1] = iy; s_org[lane + 1][2] = yaw_next;    // increments for now
                }
                __syncwarp();
                if (lane == 0) {
                  const double step = 0.5 * maxc;
                  s_org[0][0] = 0.0; s_org[0][1] = 0.0; s_org[0][2] = 0.0;
                  for (int i = 0; i < nseg; ++i) { s_org[i + 1][0] = s_org[i][0] + s_org[i + 1][0]; s_org[i + 1][1] = s_org[i][1] + s_org[i + 1][1]; }
                  int ind = 1; double d, pd, ll = 0.0;
                  CYAW[0] = 0.0; CDIR[0] = -1;                    // point 0 is never written by interpolate
                  for (int i = 0; i < nseg; ++i) {
                    const double l = s_best.len[i];
                    d = (l > 0.0) ? step : -step;
                    ind -= 1;
                    if (i >= 1 && (s_best.len[i - 1] * s_best.len[i]) > 0) pd = -d - ll; else pd = d - ll;
                    while (fabs(pd) <= fabs(l) && ind + 2 < AVP_COURSE_CAP) { ind += 1; CYAW[ind] = pd; CDIR[ind] = i; pd += d; }
                    if (ind + 2 >= AVP_COURSE_CAP) { s_shot_bad = 2; break; }
                    ll = l - pd - d;
                    ind += 1; CYAW[ind] = l; CDIR[ind] = i;
                  }
                  s_nplan = ind + 1;
                }
                __syncwarp();
                if (!s_shot_bad) {
                  const int nplan = s_nplan;
                  for (int j = lane; j < nplan; j += 32) {
                    if (j == 0) { CX[0] = 0.0; CY[0] = 0.0; CYAW[0] = 0.0; CDIR[0] = (s_best.len[0] > 0.0) ? 1 : -1; continue; }
                    const int seg = CDIR[j]; const double l = CYAW[j];
                    double px, py, pyaw = 0.0; int dir;
                    rs_interpolate(l, mode[seg], maxc, s_org[seg][0], s_org[seg][1], s_org[seg][2], px, py, pyaw, dir);
                    if (mode[seg] == 'S') pyaw = s_org[seg][2];
                    CX[j] = px; CY[j] = py; CYAW[j] = pyaw; CDIR[j] = dir;
                  }
                  __syncwarp();
                  if (lane == 0) { int n = nplan; while (n > 0 && CX[n - 1] == 0.0) --n; npts = n; }     // trailing points with local x == 0.0 dropped (rs_curve.py:588-592)
                }
              }
            }
            __syncwarp();
            if (lane == 0) {
              // the points are checked in the order (k * stride) mod npts, stride ~ 0.38 npts and coprime to npts: the node's own
              // neighbourhood is free, so a colliding course is found after fewer checks than in path order (any hit decides)
              int st = 1;
#ifndef AVP_NO_CSTRIDE
              if (npts > 4) { st = (npts * 49 + 64) >> 7; for (;;) { int a = npts, b = st; while (b) { const int t_ = a % b; a = b; b = t_; } if (a == 1) break; ++st; } }
#endif
              s_npts = npts; s_cstride = st; __threadfence_block(); st_release_cta(&s_course_rdy, 1);
            }
          } else if (it == 1) {
            if (lane < nchild) {
              const int c = lane;
              double q0[3];
              const double tn = cfg.tan_steer[c % cfg.steering_angle_num];
              const double speed = (c < nchild / 2.0) ? cfg.max_v : -cfg.max_v;
              const double td = speed * cfg.dt;
              q0[2] = pi_2_pi(T.theta + (cfg.max_v * tn) / cfg.lw * cfg.dt);
              const double cs = d_cos(q0[2]), sn = d_sin(q0[2]);
              q0[0] = T.x + td * cs; q0[1] = T.y + td * sn;
              W.cpose[c][0] = q0[0]; W.cpose[c][1] = q0[1]; W.cpose[c][2] = q0[2];
              rs_query_cs(q0, cs, sn, goal, maxc, s_Q[c]);
            }
            __syncwarp();
            if (lane == 0) { __threadfence_block(); st_release_cta(&s_q_rdy, 1); }
          } else if (it == 2) {
            const int nsub = nsubs <= 4 ? nsubs : 4;
            for (int item = lane; item < nchild * nsub; item += 32) {
              const int i = item / nsub, k = item % nsub;
              const double tn = cfg.tan_steer[i % cfg.steering_angle_num];
              const double speed = (i < nchild / 2.0) ? cfg.max_v : -cfg.max_v;
              const double td_i = speed * cfg.ddt * (k + 1);
              const double th_i = pi_2_pi(T.theta + (cfg.max_v * tn) / cfg.lw * cfg.ddt * (k + 1));
              const double cs = d_cos(th_i), sn = d_sin(th_i);
              s_sub[i][k][0] = T.x + td_i * cs; s_sub[i][k][1] = T.y + td_i * sn; s_sub[i][k][2] = cs; s_sub[i][k][3] = sn;
            }
            __syncwarp();
            if (lane == 0) { __threadfence_block(); st_release_cta(&s_sub_rdy, 1); }
          } else if (it == 3) {
            if (!wait_ge_cta(&s_q_rdy, 1)) { if (lane == 0) s_status = AVP_CAPACITY; break; }
            if (lane < nchild) {
              const int i = lane;
              const double x_ = W.cpose[i][0], y_ = W.cpose[i][1];
              W.inrad[i] = sqrt(d_pow2(x_ - goal[0]) + d_pow2(y_ - goal[1])) < cfg.flag_radius;       // hybrid_a_star.py:308-310
              const long long id = map_index(S, x_, y_);
              const int hid = (id >= 0 && id < S.n_ids) ? (int)id : -1;
              W.hid[i] = hid;
            }
            // read ahead for the commit warp's lookups (hybrid_a_star.py:154-172, :272): the probe of the exact-pose table and the
            // h-table value.  The table inserts of the commit running beside this evaluation are finished (s_ins_done) and nothing
            // else is inserted before this result is used, so the probe is final; an h value can only change from missing to set.
            int fnd = -1, hv = -1;
#ifndef AVP_NO_LOOKAHEAD
            if (!wait_ge_cta(&s_ins_done, 1)) { if (lane == 0) s_status = AVP_CAPACITY; break; }
            if (lane < nchild) {
              const int hid = W.hid[lane];
              if (hid >= 0) hv = *(volatile const int32_t *)&hval[hid];
              fnd = htab_find_cg(htab, hmask, nodes, W.cpose[lane][0], W.cpose[lane][1], W.cpose[lane][2]);
            }
#endif
            if (lane < nchild) { W.found[lane] = fnd; W.hv[lane] = hv; }
          } else if (it <= 3 + RS_NITEM) {
            if (!wait_ge_cta(&s_q_rdy, 1)) { if (lane == 0) s_status = AVP_CAPACITY; break; }
            const int k = lane / nchild, row = lane - k * nchild;
            const int inst = (k < 3) ? rs_item_inst[it - 4][k] : -1;
            if (inst >= 0) {
              double t, u, v;
              if (rs_eval_instance(inst, s_Q[row], t, u, v)) {
                RsCand c; c.t = t; c.u = u; c.v = v; c.L = 0.0;
                c.L = rs_cand_L(inst, c, 1, 1);
                s_cand[row][inst] = c; atomicOr(&s_valid[row], 1ull << inst);
              }
            }
            __syncwarp();
            if (lane == 0) { __threadfence_block(); add_release_cta(&s_rs_done, 1); }
          } else {
            if (!wait_ge_cta(&s_sub_rdy, 1)) { if (lane == 0) s_status = AVP_CAPACITY; break; }
            const int i = it - 4 - RS_NITEM;
            int coll = 0;
            for (int k = 0; k < nsubs; ++k) {
              bool hit;
              if (k < 4) hit = check_pose_cs_warp_sm(cfg, S, cells, col_start, s_sub[i][k][0], s_sub[i][k][1], s_sub[i][k][2], s_sub[i][k][3], vg);
              else {
                const double tn = cfg.tan_steer[i % cfg.steering_angle_num];
                const double speed = (i < nchild / 2.0) ? cfg.max_v : -cfg.max_v;
                const double td_i = speed * cfg.ddt * (k + 1);
                const double th_i = pi_2_pi(T.theta + (cfg.max_v * tn) / cfg.lw * cfg.ddt * (k + 1));
                const double cs = d_cos(th_i), sn = d_sin(th_i);
                hit = check_pose_cs_warp_sm(cfg, S, cells, col_start, T.x + td_i * cs, T.y + td_i * sn, cs, sn, vg);
              }
              if (hit) { coll = 1; break; }
            }
            if (lane == 0) s_chit[i] = coll;
          }
#ifndef AVP_NO_PROFILE
          if (lane == 0 && it < 40) atomicAdd(reinterpret_cast<unsigned long long *>(&s_ic[it]), (unsigned long long)(clock64() - ti_));
#endif
        }
        if (warp != 0) { WP_ACC(1); TS(5); WP_START(); PIPE_TICK(32, 12); }
        // ---- E2: the items that consume other items' results:
        //   course point checks (hybrid_a_star.py:334-347), after s_course_rdy, in strided order, skipped once one of them hit
        //   word selection, two successors per warp item (lanes 0..10 / 16..26 = ctype groups): set_path de-duplication + minimum
        //   per group, then calc_optimal_path (combine the groups) and the word kept for the successor's own shot; after every
        //   rs item is finished (s_rs_done).  A warp takes a selection whenever the rs items are finished, else a course point.
        // the commit warp only helps while E1 items are left: it arrives late, and a selection or a course point taken then
        // would make it the last warp at barrier A
        if (warp == 0) { WP_ACC(3); continue; }
        if (!wait_ge_cta(&s_course_rdy, 1)) { if (lane == 0) s_status = AVP_CAPACITY; continue; }
        const int npts = s_npts, cstride = s_cstride;
        const int n_sel = (nchild + 1) / 2;
        bool sel_left = true, chk_left = npts > 0;
        for (;;) {
          const long long ti_ = clock64();
          int kind = -1, it = 0;                      // 0: selection, 1: course point
          // the choice must be the same on every lane (the branches contain shuffles): lane 0 reads the counter
          int rsd = 0;
          if (lane == 0) rsd = ld_acquire_cta(&s_rs_done);
          rsd = __shfl_sync(AVP_FULL_MASK, rsd, 0);
          if (sel_left && (!chk_left || rsd >= RS_NITEM)) {
            if (!wait_ge_cta(&s_rs_done, RS_NITEM)) { if (lane == 0) s_status = AVP_CAPACITY; break; }
            if (lane == 0) it = atomicAdd(&s_work3, 1);
            it = __shfl_sync(AVP_FULL_MASK, it, 0);
            if (it >= n_sel) { sel_left = false; continue; }
            kind = 0;
          } else if (chk_left) {
            if (lane == 0) it = atomicAdd(&s_work2, 1);
            it = __shfl_sync(AVP_FULL_MASK, it, 0);
            if (it >= npts) { chk_left = false; continue; }
            kind = 1;
          } else break;
          if (kind == 1) {
            const int j = (int)(((long long)it * cstride) % npts);
            const int stop = __shfl_sync(AVP_FULL_MASK, *(volatile int *)&s_shot_coll, 0);   // warp-uniform early exit
            if (stop) { chk_left = false; continue; }
            const double ix = CX[j], iy = CY[j], cm = s_tcs[0], sm = s_tcs[1];
            const double gx_ = cm * ix + sm * iy + T.x, gy_ = -sm * ix + cm * iy + T.y;      // rs_curve.py:124-130
            const double gyaw = pi_2_pi(CYAW[j] + T.theta);
            const double gth = pi_2_pi(gyaw);
            if (check_pose_cs_warp_sm(cfg, S, cells, col_start, gx_, gy_, d_cos(gth), d_sin(gth), vg)) { if (lane == 0) s_shot_coll = 1; }
          } else {
            const int half = lane >> 4, gl = lane & 15, i = 2 * it + half;
            if (i < nchild && gl < RS_NGROUP) rs_select_group(s_cand[i], s_valid[i], gl, 1, 1, maxc, s_grp[i][gl]);
            __syncwarp();
            if (i < nchild && gl == 0) {
              RsBest b; rs_combine_groups(s_grp[i], s_cand[i], 1, 1, b);
              const int ok = (b.ok && !b.degenerate) ? 1 : 0;
              W.rsok[i] = ok;
              W.rsL[i] = b.ok ? b.L / maxc : 0.0;
              NodeShot w; w.t = 0.0; w.u = 0.0; w.v = 0.0; w.L = 0.0; w.inst = -1; w.ok = 0;
              if (b.ok) { w.t = s_cand[i][b.inst].t; w.u = s_cand[i][b.inst].u; w.v = s_cand[i][b.inst].v; w.L = b.L; w.inst = b.inst; w.ok = ok; }
              W.shot[i] = w;
            }
          }
#ifndef AVP_NO_PROFILE
          if (lane == 0) { const int k_ = 40 + kind; atomicAdd(reinterpret_cast<unsigned long long *>(&s_ic[k_]), (unsigned long long)(clock64() - ti_)); atomicAdd(reinterpret_cast<unsigned long long *>(&s_ic[k_ + 2]), 1ull); }
#endif
        }
        WP_ACC(2); TS(7); PIPE_TICK(32, 13);
      }
    }
    __syncthreads();

    // ---- finish: summary + finish_path (hybrid_a_star.py:351-389) + rs tail (path_planner.py:100-108)
    if (lane == 0 && P.wprof && warp < 16) { long long *o = P.wprof + ((size_t)sc * 16 + warp) * 24; for (int k = 0; k < 8; ++k) o[k] = s_wp[warp][k]; }
    if (tid < 48 && P.wprof) P.wprof[((size_t)sc * 16 + (tid >> 3)) * 24 + 16 + (tid & 7)] = s_ic[tid];
    if (tid == 32 && P.prof) { long long *o = P.prof + (size_t)sc * 16; o[7] = pc[7]; o[12] = pc[12]; o[13] = pc[13]; { unsigned sm_; asm("mov.u32 %0, %%smid;" : "=r"(sm_)); o[14] = ((long long)sm_ << 40) | (t_start & 0xffffffffffll); } o[15] = pc[15]; }     // o[14]: SM id and start clock (which scenarios shared a TPC, development aid)
    if (tid == 0) {
      avp_plan_summary &R = P.sums[sc];
      int status = s_status;
      if (!status && !reached) status = (s_in_radius && s_best_ok) ? AVP_OPEN_EXHAUSTED_RS : AVP_OPEN_EXHAUSTED;
      R.status = status; R.n_pops = s_npops; R.global_index = s_G; R.n_closed = s_nclosed; R.n_open = s_on;
      R.last_index = s_cur; R.n_hq = s_nhq; R.h_closed = s_D.closed_len; R.nx = S.nx; R.ny = S.ny; R.n_obs = S.n_obs;
      R.n_hcalls = s_nhcalls; R.pitch[0] = S.dx; R.pitch[1] = S.dy;
      for (int i = 0; i < 4; ++i) R.boundary[i] = S.b[i];
      R.origin[0] = S.b[0]; R.origin[1] = S.b[2];
      R.n_astar = 0; R.n_rs = 0; R.n_final = 0; R.rs_nseg = 0; R.rs_L = 0.0;
      for (int i = 0; i < 5; ++i) R.rs_lengths[i] = 0.0;
      for (int i = 0; i < 8; ++i) R.rs_ctypes[i] = 0;
      if (status == AVP_OK || status == AVP_OPEN_EXHAUSTED_RS) {
        // the last popped node's rs path (hybrid_a_star.py:326-332), recomputed from its stored word
        const NodeShot w = nshot[s_cur];
        const Node &ln = nodes[s_cur];
        RsBest b; b.ok = 1; b.degenerate = 0; b.inst = w.inst; b.L = w.L;
        unsigned mask;
        b.n = rs_arrange(w.inst, w.t, w.u, w.v, 1, s_cur != 0, b.len, b.ct, mask);
        const double q0[3] = {ln.x, ln.y, ln.theta};
        int npts = rs_course(b, maxc, 0.5, q0, AVP_COURSE_CAP, CX, CY, CYAW, CDIR);
        if (npts < 0) npts = 0;
        double *fp = P.paths + (size_t)sc * P.cap_path * 3;
        int np_ = 0;
        int depth = 0; for (int k = s_cur; k != 0; k = nodes[k].parent) ++depth;
        auto push = [&](double px, double py, double pt) { if (np_ < P.cap_path) { fp[3 * np_] = px; fp[3 * np_ + 1] = py; fp[3 * np_ + 2] = pt; } ++np_; };
        push(nodes[0].x, nodes[0].y, nodes[0].theta);
        for (int lvl = 1; lvl <= depth; ++lvl) {
          int ch = s_cur; for (int k = 0; k < depth - lvl; ++k) ch = nodes[ch].parent;
          const Node &c = nodes[ch]; const Node &par = nodes[c.parent];
          for (int j = 0; j < cfg.n_substeps; ++j) {
            const double speed = c.forward ? cfg.max_v : -cfg.max_v;
            const double td_j = speed * cfg.ddt * (j + 1);
            const double th_j = pi_2_pi(par.theta + (cfg.max_v * cfg.tan_steer[c.steer_idx]) / cfg.lw * cfg.ddt * (j + 1));
            push(par.x + td_j * d_cos(th_j), par.y + td_j * d_sin(th_j), th_j);
          }
        }
        R.n_astar = np_;
        for (int i = 1; i < npts; ++i) push(CX[i], CY[i], CYAW[i]);
        R.n_final = np_; R.n_rs = npts; R.rs_nseg = b.n; R.rs_L = b.L / maxc;
        for (int i = 0; i < b.n; ++i) R.rs_lengths[i] = b.len[i] / maxc;
        for (int i = 0; i < 8; ++i) R.rs_ctypes[i] = rs_ct_names[b.ct][i];
      }
      if (dbg) dbg[0] = 9;
      if (P.prof) { long long *o = P.prof + (size_t)sc * 16; for (int k = 0; k < 7; ++k) o[k] = pc[k]; for (int k = 8; k < 12; ++k) o[k] = pc[k]; }
    }
    __syncthreads();
  }
#undef PIPE_TICK
#undef WP_START
#undef WP_ACC
#undef TS
}

template <int BLOCK>
__global__ void __launch_bounds__(BLOCK, 1) k_search_pipe(KParams P) { search_pipe_body<BLOCK>(P); }

