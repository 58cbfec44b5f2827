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
//   barrier A | C: accept the result of the popped node; lookups + node records (lanes 0..9);    | barrier B
//             |    predict the next pop; publish it as the evaluators' target                    |
//   barrier B | C: sequential commit in slot order (Dijkstra resumes inside), then heappop       | barrier A
//             | E: successor poses -> rs words || course points -> selection, collision checks   |
//             |    of the successors' sub-steps and of the course points (dynamic queue) -> combine
//
// The evaluators synchronise among themselves with named barrier 1; A and B are __syncthreads().
// The word of the goal shot is the word selected when the node was scored (NodeShot), so the shot costs
// no rs solve; its course is planned while the successor poses are computed and its points are
// collision-checked in the same queue as the successors' sub-steps.
#pragma once
#include <cooperative_groups.h>
#include "avp_kernels.cuh"
namespace cg = cooperative_groups;

enum { CTL_FINISH = 2 };

// The rs warp items of the E1 queue: up to three word instances per item (lanes = instance slot x successor), formed so
// that a warp runs one word formula where possible (4 instances per family: 3 + 1 left over, the left-overs grouped by
// shared code), and ordered longest first: the tau/omega families (LRLRn 18-21, LRLRp 22-25) as half-size items in front.
#define RS_NITEM 17
__device__ __constant__ int8_t rs_item_inst[RS_NITEM][3] = {
  {18, 19, -1}, {20, 21, -1}, {22, 23, -1}, {24, 25, -1},
  {6, 7, 8}, {26, 27, 28}, {34, 35, 36}, {42, 43, 44}, {9, 29, 37},
  {10, 11, 12}, {14, 15, 16}, {45, 13, 17}, {2, 3, 4}, {30, 31, 32}, {38, 39, 40}, {5, 33, 41}, {0, 1, -1}};

struct PureRes {
  double cpose[AVP_NCHILD_MAX][3];
  double rsL[AVP_NCHILD_MAX];
  NodeShot shot[AVP_NCHILD_MAX];
  int32_t coll[AVP_NCHILD_MAX], rsok[AVP_NCHILD_MAX], inrad[AVP_NCHILD_MAX], hid[AVP_NCHILD_MAX];
  int32_t node;                              // the node this result belongs to (-1: none)
  int32_t in_radius, shot_ok;                // the shot's collision / degeneracy flags stay in shared memory (s_shot_coll, s_shot_bad)
};
struct EvalTarget { double x, y, theta; NodeShot shot; int32_t node, in_radius, is_root, valid; };

// clock read that the compiler may not move across barriers or memory operations (profiling counters)
__device__ __forceinline__ long long clock_ordered() { long long t; asm volatile("mov.u64 %0, %%clock64;" : "=l"(t)::"memory"); return t; }

template <int NTHREADS>
__device__ __forceinline__ void eval_barrier() { asm volatile("bar.sync 1, %0;" ::"n"(NTHREADS) : "memory"); }

// CLUSTER: the kernel is launched as thread-block clusters of two CTAs (two SMs) per scenario.  CTA 0 is the CTA
// described above; CTA 1 (the HELPER) takes the collision work of every evaluation -- the successors' sub-step
// checks and the shot's course (plan, points, checks) -- and writes the flags into CTA 0's shared memory
// (distributed shared memory); barriers A and B become cluster barriers.  Used when the scenarios of pass 2 fit
// n_sm / 2 clusters: the evaluation is throughput bound on one SM (see profiles/), two SMs halve it.
template <int BLOCK, bool CLUSTER>
__device__ __forceinline__ void search_pipe_body(const KParams &P) {
  static_assert(BLOCK >= 128 && BLOCK % 32 == 0, "one commit warp + at least three evaluator warps");
  constexpr int NWARPS = BLOCK / 32, NE = NWARPS - 1, ET = NE * 32;
  constexpr int SMO = avp_sm_open(BLOCK);
  constexpr int SELW = NE >= 8 ? 4 : (NE >= 4 ? 2 : 1);   // evaluator warps that run the word selection before they join the collision queue
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
  __shared__ DijCtx s_D;
  __shared__ long long s_wp[BLOCK / 32][8];   // per warp: cycles lane 0 spent working in each evaluator phase (barrier waits excluded)

  const avp_config &cfg = P.cfg;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int etid = tid - 32, ewarp = warp - 1;
  const int slot = CLUSTER ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int rank = CLUSTER ? (int)(blockIdx.x & 1) : 0;          // cluster dims (2,1,1): rank in cluster = blockIdx.x & 1
  cg::cluster_group cluster = cg::this_cluster();
#define SYNC_AB() do { if (CLUSTER) cluster.sync(); else __syncthreads(); } while (0)
  const int nchild = 2 * cfg.steering_angle_num;
  const double maxc = 1 / cfg.min_radius_turn;
  Node *nodes = P.nodes + (size_t)slot * P.node_cap;
  NodeShot *nshot = P.nshot + (size_t)slot * P.node_cap;
  int32_t *htab = P.htab + (size_t)slot * P.htab_stride;
  OEnt *oge = P.oheap + (size_t)slot * P.node_cap;
  const int hmask = P.htab_size - 1;
  double *CX = P.course + (size_t)slot * 3 * AVP_COURSE_CAP, *CY = CX + AVP_COURSE_CAP, *CYAW = CY + AVP_COURSE_CAP;
  int32_t *CDIR = P.course_dir + (size_t)slot * AVP_COURSE_CAP;

  for (;;) {
    if (rank == 0 && tid == 0) s_scen = atomicAdd(P.work_counter, 1);
    SYNC_AB();
    const int scen_i = (CLUSTER && rank == 1) ? *cluster.map_shared_rank(&s_scen, 0) : s_scen;
    if (scen_i >= P.n_work) break;
    const int sc = P.work_list ? P.work_list[scen_i] : scen_i;
    const ScenDev &S = P.scen[sc];
    const double2 *cells = P.cells + S.cell_off;
    const int32_t *col_start = P.col_start + S.col_off;
    if (CLUSTER && rank == 1) {
      // =========================== HELPER CTA (cluster rank 1) ===========================
      // Every step: barrier A -> read CTA 0's exit word -> barrier B -> read the target -> sub-step poses ->
      // one queue: [0] the shot's course plan + points, [1..nchild] one successor's sub-step checks each,
      // [nchild+1 ..] the course points' checks (they wait for item 0) -> flags to CTA 0's shared memory.
      const int *r_ctlA = cluster.map_shared_rank(&s_ctlA, 0), *r_ctlB = cluster.map_shared_rank(&s_ctlB, 0);
      const EvalTarget *r_tgt = cluster.map_shared_rank(&s_tgt, 0);
      int *r_chit = cluster.map_shared_rank(&s_chit[0], 0), *r_scoll = cluster.map_shared_rank(&s_shot_coll, 0), *r_sbad = cluster.map_shared_rank(&s_shot_bad, 0);
      for (;;) {
        cluster.sync();                            // A
        if (*r_ctlA == CTL_EXIT) break;
        cluster.sync();                            // B
        if (*r_ctlB != CTL_RUN) break;
        if (tid < (int)(sizeof(EvalTarget) / 8)) reinterpret_cast<long long *>(&s_tgt)[tid] = reinterpret_cast<const long long *>(r_tgt)[tid];
        if (tid == 32) { s_shot_coll = 0; s_shot_bad = 0; s_nplan = 0; s_work = 0; s_npts = -1; }
        if (tid >= 64 && tid < 64 + nchild) s_chit[tid - 64] = 0;
        __syncthreads();
        const EvalTarget T = s_tgt;
        if (T.valid) {
          const int phi_np = !T.is_root;
          const int nsubs = cfg.n_substeps;
          {
            const int nsub = nsubs <= 4 ? nsubs : 4;
            for (int item = warp + NWARPS * lane; item < nchild * nsub; item += NWARPS * 32) {
              const int i = item / nsub, k = item % nsub;
              const double tn = cfg.tan_steer[i % cfg.steering_angle_num];
              const double speed = (i < nchild / 2.0) ? cfg.max_v : -cfg.max_v;
              const double td_i = speed * cfg.ddt * (k + 1);
              const double th_i = pi_2_pi(T.theta + (cfg.max_v * tn) / cfg.lw * cfg.ddt * (k + 1));
              const double cs = d_cos(th_i), sn = d_sin(th_i);
              s_sub[i][k][0] = T.x + td_i * cs; s_sub[i][k][1] = T.y + td_i * sn; s_sub[i][k][2] = cs; s_sub[i][k][3] = sn;
            }
          }
          __syncthreads();
          for (;;) {
            int it = 0;
            if (lane == 0) it = atomicAdd(&s_work, 1);
            it = __shfl_sync(AVP_FULL_MASK, it, 0);
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
                    double oyaw = 0.0;
                    for (int i = 0; i < lane; ++i) { if (mode[i] == 'L') oyaw = oyaw + s_best.len[i]; else if (mode[i] == 'R') oyaw = oyaw - s_best.len[i]; }
                    double ix, iy, yaw_next = oyaw; int dir;
                    rs_interpolate(s_best.len[lane], mode[lane], maxc, 0.0, 0.0, oyaw, ix, iy, yaw_next, dir);
                    s_org[lane + 1][0] = ix; s_org[lane + 1][1] = iy; s_org[lane + 1][2] = yaw_next;
                  }
                  __syncwarp();
                  if (lane == 0) {
                    const double step = 0.5 * maxc;
                    s_org[0][0] = 0.0; s_org[0][1] = 0.0; s_org[0][2] = 0.0;
                    for (int i = 0; i < nseg; ++i) { s_org[i + 1][0] = s_org[i][0] + s_org[i + 1][0]; s_org[i + 1][1] = s_org[i][1] + s_org[i + 1][1]; }
                    int ind = 1; double d, pd, ll = 0.0;
                    CYAW[0] = 0.0; CDIR[0] = -1;
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
                    if (lane == 0) { int n = nplan; while (n > 0 && CX[n - 1] == 0.0) --n; npts = n; }     // rs_curve.py:588-592
                  }
                }
              }
              __syncwarp();
              if (lane == 0) { __threadfence_block(); *(volatile int *)&s_npts = npts; }      // publishes the course: the check items may start
            } else if (it <= nchild) {
              const int i = it - 1;
              int coll = 0;
              for (int k = 0; k < nsubs; ++k) {
                bool hit;
                if (k < 4) hit = check_pose_cs_warp(cfg, S, cells, col_start, s_sub[i][k][0], s_sub[i][k][1], s_sub[i][k][2], s_sub[i][k][3]);
                else {
                  const double tn = cfg.tan_steer[i % cfg.steering_angle_num];
                  const double speed = (i < nchild / 2.0) ? cfg.max_v : -cfg.max_v;
                  const double td_i = speed * cfg.ddt * (k + 1);
                  const double th_i = pi_2_pi(T.theta + (cfg.max_v * tn) / cfg.lw * cfg.ddt * (k + 1));
                  hit = check_pose_warp(cfg, S, cells, col_start, T.x + td_i * d_cos(th_i), T.y + td_i * d_sin(th_i), th_i);
                }
                if (hit) { coll = 1; break; }
              }
              if (lane == 0) s_chit[i] = coll;
            } else {
              int npts = 0;
              if (lane == 0) { while ((npts = *(volatile int *)&s_npts) < 0) __nanosleep(200); }
              npts = __shfl_sync(AVP_FULL_MASK, npts, 0);
              const int j = it - 1 - nchild;
              if (j >= npts) break;
              const int stop = __shfl_sync(AVP_FULL_MASK, *(volatile int *)&s_shot_coll, 0);
              if (stop) continue;
              const double ix = CX[j], iy = CY[j], cm = s_tcs[0], sm = s_tcs[1];
              const double gx_ = cm * ix + sm * iy + T.x, gy_ = -sm * ix + cm * iy + T.y;      // rs_curve.py:124-130
              const double gyaw = pi_2_pi(CYAW[j] + T.theta);
              if (check_pose_warp(cfg, S, cells, col_start, gx_, gy_, pi_2_pi(gyaw))) { if (lane == 0) s_shot_coll = 1; }
            }
          }
        }
        __syncthreads();
        if (tid < nchild) r_chit[tid] = s_chit[tid];
        if (tid == 32) { *r_scoll = s_shot_coll; *r_sbad = s_shot_bad; }
      }
      continue;                                    // next scenario
    }
    int32_t *hval = P.hval + S.id_off, *ost = P.ost + S.id_off;
    const double goal[3] = {S.pose[3], S.pose[4], pi_2_pi(S.pose[5])};
    int32_t *pops = P.pops ? P.pops + (size_t)sc * P.cap_pops : nullptr;
    int32_t *hql = P.hq_log ? P.hq_log + (size_t)sc * AVP_HQ_CAP * 3 : nullptr;
    int *dbg = P.dbg ? P.dbg + (size_t)sc * 8 : nullptr;
    const long long t_start = clock64();
    // cycle accumulators: thread 0 (commit warp) and thread 32 (evaluators) each keep their own
    long long pc[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, tp = t_start;
#define PIPE_TICK(who, k) do { if (tid == (who)) { const long long t_ = clock_ordered(); pc[k] += t_ - tp; tp = t_; } } while (0)
    long long wt = 0;
#define WP_START() do { if (lane == 0) wt = clock_ordered(); } while (0)
#define WP_ACC(k) do { if (lane == 0) { const long long t_ = clock_ordered(); s_wp[warp][k] += t_ - wt; wt = t_; } } while (0)
    if (lane == 0) for (int k = 0; k < 8; ++k) s_wp[warp][k] = 0;
    // timeline of ONE pop (the AVP_TRACE_POP-th of the scenario): absolute clocks of every warp at the phase boundaries
    long long *tsw = (P.wprof && warp < 16) ? P.wprof + ((size_t)sc * 16 + warp) * 24 + 8 : nullptr;
#define TS(k) do { __syncwarp(); if (lane == 0 && tsw && ((k) < 2 ? (s_npops == P.trace_pop) : s_trace_on)) tsw[k] = clock_ordered(); } while (0)

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
      SYNC_AB();                                 // ---- barrier A: the evaluators' result is complete, the next node is popped
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
              const int hvp = (id >= 0) ? hval[id] : -1;                  // issued first: overlaps the table probes
              const Node cn = nodes[cur];
              const double x_ = R.cpose[i][0], y_ = R.cpose[i][1], th = R.cpose[i][2];
              found = htab_find(htab, hmask, nodes, x_, y_, th);
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
      }
      if (warp == 0) WP_ACC(0);
      PIPE_TICK(0, 1);                           // accept + lookups + prediction
      TS(1);
      SYNC_AB();                                 // ---- barrier B: target published
      TS(2);
      PIPE_TICK(32, 15);                         // evaluators waiting for the target
      if (s_ctlB != CTL_RUN) { reached = (s_ctlB == CTL_FINISH); break; }

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
                  { const long long t_ = clock64(); oh_push<SMO>(s_of, s_oi, oge, nodes, on, f, child); pc[8] += clock64() - t_; pc[9]++; }
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
      } else {
        // =========================== EVALUATORS ===========================
        const EvalTarget T = s_tgt;
        PureRes &W = s_res[s_rb ^ 1];
        if (!T.valid) { if (etid == 0) W.node = -1; continue; }
        const int phi_np = !T.is_root;             // the root's theta is a Python float (see oracle generate_path)
        const int nsubs = cfg.n_substeps;
        WP_START();
        // ---- E0: successor poses with their normalised rs queries, and sub-step poses (hybrid_a_star.py:145-151, :185-194),
        //          spread over the warps (one or two code paths per warp)
        if (etid == 0) {
          s_nplan = 0; s_work = 0;
          if (!CLUSTER) { s_shot_coll = 0; s_shot_bad = 0; }       // cluster mode: the helper CTA owns these flags and s_chit
          for (int i = 0; i < nchild; ++i) { s_valid[i] = 0ull; if (!CLUSTER) s_chit[i] = 0; }
        }
        {
          const int nsub = CLUSTER ? 0 : (nsubs <= 4 ? nsubs : 4);
          for (int item = ewarp + NE * lane; item < nchild + nchild * nsub; item += NE * 32) {
            if (item < nchild) {
              const int c = item;
              double q0[3];
              const double tn = cfg.tan_steer[c % cfg.steering_angle_num];
              const double speed = (c < nchild / 2.0) ? cfg.max_v : -cfg.max_v;
              const double td = speed * cfg.dt;
              q0[2] = pi_2_pi(T.theta + (cfg.max_v * tn) / cfg.lw * cfg.dt);
              q0[0] = T.x + td * d_cos(q0[2]); q0[1] = T.y + td * d_sin(q0[2]);
              W.cpose[c][0] = q0[0]; W.cpose[c][1] = q0[1]; W.cpose[c][2] = q0[2];
              rs_query(q0, goal, maxc, s_Q[c]);
            } else {
              const int nsd = nsub > 0 ? nsub : 1, i = (item - nchild) / nsd, k = (item - nchild) % nsd;
              const double tn = cfg.tan_steer[i % cfg.steering_angle_num];
              const double speed = (i < nchild / 2.0) ? cfg.max_v : -cfg.max_v;
              const double td_i = speed * cfg.ddt * (k + 1);
              const double th_i = pi_2_pi(T.theta + (cfg.max_v * tn) / cfg.lw * cfg.ddt * (k + 1));
              const double cs = d_cos(th_i), sn = d_sin(th_i);
              s_sub[i][k][0] = T.x + td_i * cs; s_sub[i][k][1] = T.y + td_i * sn; s_sub[i][k][2] = cs; s_sub[i][k][3] = sn;
            }
          }
        }
        WP_ACC(0);
        TS(3);
        eval_barrier<ET>();
        TS(4);
        WP_START();
        PIPE_TICK(32, 7);                          // E0
        // ---- E1: one queue of warp items, longest first:
        //   0                      the shot's course from the node's stored word: plan (generate_local_course, rs_curve.py:537-594)
        //                          and points in the local frame (:597-624)
        //   1 .. RS_NITEM          rs word instances (rs_item_inst: up to three instances x all successors per warp item)
        //   RS_NITEM+1 ..          the successors' sub-step collision checks (hybrid_a_star.py:185-204), one successor each
        {
          const int n_items = CLUSTER ? RS_NITEM : 1 + RS_NITEM + nchild;       // cluster mode: the helper CTA has items 0 and RS_NITEM+1..
          for (;;) {
            int it = 0;
            if (lane == 0) it = atomicAdd(&s_work, 1);
            it = __shfl_sync(AVP_FULL_MASK, it, 0);
            if (it >= n_items) break;
            if (CLUSTER) it += 1;
            if (it == 0) {
              if (!T.in_radius) continue;
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
              if (s_shot_bad) continue;
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
              if (s_shot_bad) continue;
              const int nplan = s_nplan;
              for (int j = lane; j < nplan; j += 32) {
                if (j == 0) { CX[0] = 0.0; CY[0] = 0.0; CYAW[0] = 0.0; CDIR[0] = (s_best.len[0] > 0.0) ? 1 : -1; continue; }
                const int seg = CDIR[j]; const double l = CYAW[j];
                double px, py, pyaw = 0.0; int dir;
                rs_interpolate(l, mode[seg], maxc, s_org[seg][0], s_org[seg][1], s_org[seg][2], px, py, pyaw, dir);
                if (mode[seg] == 'S') pyaw = s_org[seg][2];
                CX[j] = px; CY[j] = py; CYAW[j] = pyaw; CDIR[j] = dir;
              }
            } else if (it <= RS_NITEM) {
              const int k = lane / nchild, row = lane - k * nchild;
              const int inst = (k < 3) ? rs_item_inst[it - 1][k] : -1;
              if (inst >= 0) {
                double t, u, v;
                if (rs_eval_instance(inst, s_Q[row], t, u, v)) {
                  RsCand c; c.t = t; c.u = u; c.v = v; c.L = 0.0;
                  c.L = rs_cand_L(inst, c, 1, 1);
                  s_cand[row][inst] = c; atomicOr(&s_valid[row], 1ull << inst);
                }
              }
            } else {
              const int i = it - 1 - RS_NITEM;
              int coll = 0;
              for (int k = 0; k < nsubs; ++k) {
                bool hit;
                if (k < 4) hit = check_pose_cs_warp(cfg, S, cells, col_start, s_sub[i][k][0], s_sub[i][k][1], s_sub[i][k][2], s_sub[i][k][3]);
                else {
                  const double tn = cfg.tan_steer[i % cfg.steering_angle_num];
                  const double speed = (i < nchild / 2.0) ? cfg.max_v : -cfg.max_v;
                  const double td_i = speed * cfg.ddt * (k + 1);
                  const double th_i = pi_2_pi(T.theta + (cfg.max_v * tn) / cfg.lw * cfg.ddt * (k + 1));
                  hit = check_pose_warp(cfg, S, cells, col_start, T.x + td_i * d_cos(th_i), T.y + td_i * d_sin(th_i), th_i);
                }
                if (hit) { coll = 1; break; }
              }
              if (lane == 0) s_chit[i] = coll;
            }
          }
        }
        WP_ACC(1);
        TS(5);
        eval_barrier<ET>();
        TS(6);
        WP_START();
        PIPE_TICK(32, 12);                         // E1
        // ---- E2: a second queue:
        //   0 .. nchild-1   per successor: set_path de-duplication + minimum per ctype group (lanes = groups), then lane 0:
        //                   calc_optimal_path (combine the groups), the word kept for the successor's own shot, in_radius, cell id
        //   nchild ..       collision checks of the shot's course points (hybrid_a_star.py:334-347; trailing points with
        //                   local x == 0.0 dropped, rs_curve.py:588-592)
        if (etid == 0) { W.node = T.node; W.in_radius = T.in_radius; W.shot_ok = (T.in_radius && T.shot.ok) ? 1 : 0; }
        {
          int npts = 0;
          if (!CLUSTER && T.in_radius && !s_shot_bad) {
            if (lane == 0) { int n = s_nplan; while (n > 0 && CX[n - 1] == 0.0) --n; npts = n; }
            npts = __shfl_sync(AVP_FULL_MASK, npts, 0);
          }
          const int n_items = nchild + npts;
          // s_work was left >= the E1 item count by every warp; the E2 counter continues from a common base
          const int base = (CLUSTER ? RS_NITEM : 1 + RS_NITEM + nchild) + NE;
          for (;;) {
            int it = 0;
            if (lane == 0) it = atomicAdd(&s_work, 1) - base;
            it = __shfl_sync(AVP_FULL_MASK, it, 0);
            if (it >= n_items) break;
            if (it < nchild) {
              const int i = it;
              if (lane < RS_NGROUP) rs_select_group(s_cand[i], s_valid[i], lane, 1, 1, maxc, s_grp[i][lane]);
              __syncwarp();
              if (lane == 0) {
                RsBest b; rs_combine_groups(s_grp[i], s_cand[i], 1, 1, b);
                const int ok = (b.ok && !b.degenerate) ? 1 : 0;
                W.rsok[i] = ok;
                W.rsL[i] = b.ok ? b.L / maxc : 0.0;
                NodeShot w; w.t = 0.0; w.u = 0.0; w.v = 0.0; w.L = 0.0; w.inst = -1; w.ok = 0;
                if (b.ok) { w.t = s_cand[i][b.inst].t; w.u = s_cand[i][b.inst].u; w.v = s_cand[i][b.inst].v; w.L = b.L; w.inst = b.inst; w.ok = ok; }
                W.shot[i] = w;
                const double x_ = W.cpose[i][0], y_ = W.cpose[i][1];
                W.inrad[i] = sqrt(d_pow2(x_ - goal[0]) + d_pow2(y_ - goal[1])) < cfg.flag_radius;       // hybrid_a_star.py:308-310
                const long long id = map_index(S, x_, y_);
                W.hid[i] = (id >= 0 && id < S.n_ids) ? (int)id : -1;
              }
            } else {
              const int j = it - nchild;
              const int stop = __shfl_sync(AVP_FULL_MASK, *(volatile int *)&s_shot_coll, 0);   // warp-uniform early exit
              if (stop) continue;
              const double ix = CX[j], iy = CY[j], cm = s_tcs[0], sm = s_tcs[1];
              const double gx_ = cm * ix + sm * iy + T.x, gy_ = -sm * ix + cm * iy + T.y;      // rs_curve.py:124-130
              const double gyaw = pi_2_pi(CYAW[j] + T.theta);
              if (check_pose_warp(cfg, S, cells, col_start, gx_, gy_, pi_2_pi(gyaw))) { if (lane == 0) s_shot_coll = 1; }
            }
          }
        }
        WP_ACC(2);
        TS(7);
        PIPE_TICK(32, 13);                         // E2
      }
    }
    __syncthreads();

    // ---- finish: summary + finish_path (hybrid_a_star.py:351-389) + rs tail (path_planner.py:100-108)
    if (lane == 0 && P.wprof && warp < 16) { long long *o = P.wprof + ((size_t)sc * 16 + warp) * 24; for (int k = 0; k < 8; ++k) o[k] = s_wp[warp][k]; }
    if (tid == 32 && P.prof) { long long *o = P.prof + (size_t)sc * 16; o[7] = pc[7]; o[12] = pc[12]; o[13] = pc[13]; o[14] = pc[14]; o[15] = pc[15]; }
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
#undef SYNC_AB
  if (CLUSTER) cluster.sync();        // a CTA's shared memory stays valid until its peer has read the last exit word
}

template <int BLOCK>
__global__ void __launch_bounds__(BLOCK, 1) k_search_pipe(KParams P) { search_pipe_body<BLOCK, false>(P); }

template <int BLOCK>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(BLOCK, 1) k_search_pipe2(KParams P) { search_pipe_body<BLOCK, true>(P); }
