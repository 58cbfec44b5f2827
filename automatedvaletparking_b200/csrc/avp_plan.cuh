// avp_plan.cuh -- PathPlanner.a_star_plan (path_planner.py:58-110) for a whole batch: ONE persistent launch.
//
//   k_plan_init   run queue, slot pool and counters of the launch (device side, no host data)
//   k_dij_eager   hybrid_a_star.__init__'s eager Dijkstra.compute_path(x0, y0) (hybrid_a_star.py:89-91) for every
//                 scenario: one WARP per scenario (the heapq emulation of avp_kernels.cuh is a one-warp algorithm),
//                 twenty warps per SM, the queue left in the scenario's own heap array where k_plan resumes it
//   k_plan        the searches.  One CTA works on one scenario at a time: warp 0 commits, the other warps evaluate the
//                 node that will be popped next (the pipelined pop, avp_pipe_defs.cuh).  What is new is
//                 the scheduling around it:
//
// * A search is RESUMABLE.  Everything it owns lives in global memory -- per scenario: h table, Dijkstra queue, counters
//   (ScenState); per slot: nodes, open heap, exact-pose table -- and the shared-memory heads of the two heaps are written
//   back when a CTA lets go of it.  No pass barrier, no host round trip, nothing is planned twice.
// * ONE run queue (a ring of scenario ids, longest start-goal distance first).  A CTA takes the head, runs it for a
//   QUANTUM of pops and, if the search is still going and somebody is waiting, appends it to the tail: round robin.
//   Short searches (the median is 19 pops) leave the system in their first quantum; the long ones -- on the bench
//   workload 5 % of the scenarios hold 78 % of the pops -- share the SMs fairly from the first millisecond instead of
//   queueing behind each other.
// * SM pairs.  The two SMs of a TPC share instruction-fetch resources and this kernel is fetch bound (profiles/): a
//   search runs 25-30 % faster beside an idle neighbour.  While more scenarios are live than 1.5 per pair every CTA
//   works; below that the CTAs on odd SM ids stop taking work (P.spread).
#pragma once
#include "avp_pipe_defs.cuh"

// per scenario: what a suspended search needs besides its slot
struct __align__(16) ScenState {
  int32_t phase;                 // 0 new (eager Dijkstra done), 1 suspended, 2 final
  int32_t slot;                  // workspace slot (-1: none yet)
  int32_t status;                // status the eager Dijkstra ended with (0: go on)
  int32_t dij_hn, dij_closed;    // Dijkstra queue length, len(closedlist)
  int32_t nhq, nhcalls, G, nclosed, npops, on, cur, in_radius, best_ok;
  int32_t quanta, pad;
};
struct PlanCtl { int q_head, q_tail, finalised, slot_head, slot_tail, n_suspends, n_requeues, error, fresh_next, pad[3]; };

struct PlanParams {
  KParams K;                     // cfg, scenario arrays, results (work_list = initial order of the run queue)
  ScenState *state;
  PlanCtl *ctl;
  int32_t *queue; int q_mask;    // ring of scenario ids, -1 = empty cell
  int32_t *slot_ring; int slot_mask; int n_slots;
  int quantum;                   // pops per turn while others wait
  int spread_max;                // the CTAs on odd SM ids only work while more scenarios than this are live (P.K.spread)
  int phase;                     // 0: one launch does everything; 1: the narrow first launch (fresh scenarios from work_list, a few pops each, then
                                 //    every unfinished search is left in the run queue); 2: the wide second launch (the run queue)
  int cell_smem;                 // bytes of the TMA staging window in dynamic shared memory (0: the cell list stays in global memory)
  int overflow_odd;              // the CTAs on odd SM ids take searches that WAIT in the queue even while they stand back (AVP_OVERFLOW_ODD=0: off)
  int force_yield;               // development aid: let go of the scenario at EVERY quantum end (tests of the suspend / resume path on small batches)
};

// multi-producer / multi-consumer ring of non-negative ints; at most (mask + 1) entries are ever in flight, so a cell is
// empty again before its index comes round.  Called by ONE thread of a CTA.
__device__ __forceinline__ int ring_pop(int *head, int *tail, int32_t *buf, int mask) {
  for (;;) {
    const int h = *(volatile int *)head, t = *(volatile int *)tail;
    if (h - t >= 0) return -1;
    if (atomicCAS(head, h, h + 1) == h) {
      int v; long long spins = 0;
      while ((v = atomicExch(&buf[h & mask], -1)) == -1) { if (++spins > (1ll << 28)) return -3; }     // the producer has reserved the cell and writes it next (the bound turns a protocol bug into an error, not a hang)
      __threadfence();                                                // acquire: the producer's state is visible to this CTA
      return v;
    }
  }
}
__device__ __forceinline__ void ring_push(int *tail, int32_t *buf, int mask, int v) {
  __threadfence();                                                    // release: everything written for the consumer
  const int t = atomicAdd(tail, 1);
  long long spins = 0;
  while (atomicCAS(&buf[t & mask], -1, v) != -1) { if (++spins > (1ll << 28)) return; }
}

// ---- TMA (bulk asynchronous copy, cp.async.bulk + mbarrier) of a scenario's obstacle cell list and column starts into shared
// memory when a CTA takes the scenario over: every collision check of the following quantum (about 40 per pop) scans that
// list (collision_check.py:55-69), so it is staged once instead of being re-fetched through L1 beside the node / heap /
// table traffic.  Lists larger than AVP_CELL_SMEM stay in global memory (same code path, generic loads).
#ifndef AVP_CELL_SMEM
#define AVP_CELL_SMEM (48 * 1024)
#endif
#ifndef AVP_COURSE_BATCH
#define AVP_COURSE_BATCH 1      // course points per check item (1 .. AVP_MULTI_POSE); measured: 4 neighbouring points per item +3.5 % step time -- the
                                // points are checked by different warps at the same time and the first hit ends the scan, a batch only makes the items longer
#endif
#define AVP_CAND_SMEM ((int)sizeof(RsCandX) * AVP_NCHILD_MAX * RS_NINST)
#define AVP_CAND_IN_SMEM(block) ((block) >= 128)      // 64-thread CTAs (eight per SM) keep the candidates in global memory (L1 / L2 resident: 40 KB per CTA)
#define AVP_PLAN_DYN_SMEM(block, cell_smem) (12 * avp_sm_open(block) + (AVP_CAND_IN_SMEM(block) ? AVP_CAND_SMEM : 0) + (cell_smem))
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, unsigned bytes, unsigned long long *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
  asm volatile("{\n\t.reg .pred p;\n\tAVP_MBAR_WAIT:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@!p bra AVP_MBAR_WAIT;\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

__global__ void k_plan_init(PlanParams P) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i <= P.q_mask) P.queue[i] = (P.phase == 0 && i < P.K.n_work) ? (P.K.work_list ? P.K.work_list[i] : i) : -1;
  if (i <= P.slot_mask) P.slot_ring[i] = (i < P.n_slots) ? i : -1;
  if (i == 0) { PlanCtl c; c.q_head = 0; c.q_tail = (P.phase == 0) ? P.K.n_work : 0; c.fresh_next = 0; c.pad[0] = c.pad[1] = c.pad[2] = 0; c.finalised = 0; c.slot_head = 0; c.slot_tail = P.n_slots; c.n_suspends = 0; c.n_requeues = 0; c.error = 0; *P.ctl = c; *P.K.work_counter = 0; }
}

#define AVP_DIJ_WARPS 4
// one warp per scenario, persistent warps pulling scenarios (longest first) from an atomic counter
__global__ void __launch_bounds__(AVP_DIJ_WARPS * 32) k_dij_eager(PlanParams P) {
  __shared__ unsigned long long s_heap[AVP_DIJ_WARPS][AVP_SM_HEAP];
  __shared__ DijCtx s_D[AVP_DIJ_WARPS];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const KParams &K = P.K;
  for (;;) {
    int it = 0;
    if (lane == 0) it = atomicAdd(K.work_counter, 1);
    it = __shfl_sync(AVP_FULL_MASK, it, 0);
    if (it >= K.n_work) break;
    const int sc = K.work_list ? K.work_list[it] : it;
    const ScenDev &S = K.scen[sc];
    int32_t *hval = K.hval + S.id_off, *ost = K.ost + S.id_off;
    unsigned long long *gheap = K.dheap + (size_t)sc * K.dheap_cap;
    for (int i = lane; i < S.n_ids; i += 32) { hval[i] = -1; ost[i] = -1; }
    DijCtx &D = s_D[w];
    if (lane == 0) {
      D.S = &S; D.cost = K.cost + S.cost_off; D.hval = hval; D.ost = ost; D.gx = K.gx + S.id_off; D.gy = K.gy + S.id_off;
      D.sheap = s_heap[w]; D.gheap = gheap; D.gcap = K.dheap_cap; D.hn = 0; D.closed_len = 0; D.status = 0;
    }
    __syncwarp();
    int status = S.raster_error ? AVP_RASTER_AMBIGUOUS : 0, nhq = 0;
    if (!status) {
      long long term;
      const int d = dij_compute_path(D, s_heap[w], S.pose[0], S.pose[1], &term);
      __syncwarp();
      if (lane == 0 && K.hq_log) { int32_t *hql = K.hq_log + (size_t)sc * AVP_HQ_CAP * 3; hql[0] = (int)term; hql[1] = d; hql[2] = D.closed_len; }
      nhq = 1;
      if (d < 0) status = D.status ? D.status : AVP_H_UNREACHABLE;     // the reference never returns from this compute_path: no root node
    }
    const int hn = D.hn;
    for (int i = lane; i < hn && i < AVP_SM_HEAP; i += 32) gheap[i] = s_heap[w][i];
    if (lane == 0) {
      ScenState st; st.phase = 0; st.slot = -1; st.status = status; st.dij_hn = hn; st.dij_closed = D.closed_len;
      st.nhq = nhq; st.nhcalls = 0; st.G = 0; st.nclosed = 0; st.npops = 0; st.on = 0; st.cur = -1; st.in_radius = 0; st.best_ok = 0; st.quanta = 0; st.pad = 0;
      P.state[sc] = st;
    }
    __syncwarp();
  }
}

#ifdef AVP_PROFILE
#define PROF(...) __VA_ARGS__
#else
#define PROF(...)
#endif
// AVP_PROFILE_LIGHT: four shared-memory accumulators and nothing else (the full profile's counters live in local memory and
// perturb the commit warp): [0] evaluators waiting at barrier A, [1] barrier A -> barrier B as the evaluators see it (the
// serial section), [2] commit warp waiting at barrier A, [3] pops; written to prof[sc][0..3]
#ifdef AVP_PROFILE_LIGHT
#define LPROF(...) __VA_ARGS__
#else
#define LPROF(...)
#endif

template <int BLOCK>
__global__ void __launch_bounds__(BLOCK, (BLOCK >= 512 ? 1 : 512 / BLOCK)) k_plan(PlanParams PP) {
  static_assert(BLOCK >= 64 && BLOCK % 32 == 0, "one commit warp + at least one evaluator warp");
  const KParams &P = PP.K;
  constexpr int SMO = avp_sm_open(BLOCK);
  extern __shared__ __align__(16) unsigned char s_dyn[];
  double *s_of = reinterpret_cast<double *>(s_dyn);
  int32_t *s_oi = reinterpret_cast<int32_t *>(s_dyn + sizeof(double) * SMO);
  __shared__ unsigned long long s_heap[AVP_SM_HEAP];
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
  __shared__ int s_chit[AVP_NCHILD_MAX];
  __shared__ int s_scen, s_slot, s_phase, s_odd, s_limit;
  __shared__ int s_ctlA, s_ctlB, s_cur, s_do_commit, s_rb, s_nplan, s_npts, s_shot_coll, s_shot_bad, s_work;
  __shared__ int s_G, s_nclosed, s_npops, s_status, s_nhq, s_nhcalls, s_on, s_in_radius, s_best_ok;
  __shared__ int s_work2, s_work3, s_rs_done, s_course_rdy, s_q_rdy, s_sub_rdy, s_ins_done, s_cstride;   // the evaluators' queue: tail counter, finished rs items, course / rs queries / sub-step poses published, table inserts of the running commit done, stride of the course point order
  __shared__ VehGeom s_vg[BLOCK / 32][AVP_MULTI_POSE];   // per warp: the vehicle rectangles of the poses being checked (check_distance_multi_sm; [0]: check_distance_warp_sm)
  __shared__ DijCtx s_D;
#ifdef AVP_SCEN_SMEM
  __shared__ __align__(16) ScenDev s_S;
#endif
  __shared__ int s_gcnt[RS_NGROUP];                             // word instances of each ctype group evaluated so far (the warp that completes a group selects its winner)
  __shared__ int s_sift_n;                                      // do_pop: heap size before the pop whose sift the whole commit warp runs (0: none)
  __shared__ __align__(8) unsigned long long s_cell_bar;       // mbarrier of the staged cell list
  RsCandX (*s_cand)[RS_NINST] = reinterpret_cast<RsCandX (*)[RS_NINST]>(AVP_CAND_IN_SMEM(BLOCK) ? s_dyn + 12 * SMO : PP.K.cand_scratch + (size_t)blockIdx.x * AVP_CAND_SMEM);   // AVP_CAND_SMEM bytes: word candidates of the successors, with their arranged lengths
  unsigned char *s_cells = s_dyn + 12 * SMO + (AVP_CAND_IN_SMEM(BLOCK) ? AVP_CAND_SMEM : 0);   // PP.cell_smem bytes: double2 cells, then the int32 column starts
  unsigned cell_parity = 0;
  LPROF(__shared__ long long s_lp[12]; __shared__ long long s_lpe, s_lpc, s_lpt;)
#ifdef AVP_PROFILE
  __shared__ int s_trace_on;
  __shared__ long long s_ic[48];              // cycles per queue item (0..39: E1 items, 40/41: selections / course checks, 42/43: their counts)
  __shared__ long long s_wp[BLOCK / 32][8];   // per warp: cycles lane 0 spent working in each evaluator phase (barrier waits excluded)
  __shared__ long long s_subt[16];            // thread 0: sub-intervals of the serial section between barriers A and B
#endif

  const avp_config &cfg = P.cfg;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nchild = 2 * cfg.steering_angle_num;
  const double maxc = 1 / cfg.min_radius_turn;
  const int hmask = P.htab_size - 1;
  double *CX = P.course + (size_t)blockIdx.x * 3 * AVP_COURSE_CAP, *CY = CX + AVP_COURSE_CAP, *CYAW = CY + AVP_COURSE_CAP;
  int32_t *CDIR = P.course_dir + (size_t)blockIdx.x * AVP_COURSE_CAP;
  PlanCtl *ctl = PP.ctl;

  if (tid == 0) { unsigned sm_; asm("mov.u32 %0, %%smid;" : "=r"(sm_)); s_odd = (int)(sm_ & 1u); }
  if (tid == 0) mbar_init(&s_cell_bar, 1);
  __syncthreads();

  for (;;) {
    // ---- take the head of the run queue (thread 0).  The CTAs on odd SM ids stand back while few scenarios are live (see the
    //      header); if the live count has not moved for ~0.25 s they work anyway (no even CTA resident, e.g. a shared device).
    if (tid == 0) {
      int got = -1;
      long long t_last = clock64(); int last = -1; const long long t_idle = t_last; bool seen_waiting = false; long long t_wait = 0;
      if (PP.phase == 1) {                       // the narrow first launch: the next fresh scenario of the work list, or done
        const int k = atomicAdd(&ctl->fresh_next, 1);
        if (k < P.n_work) got = P.work_list ? P.work_list[k] : k;
      } else
      for (;;) {
        const int fin = *(volatile int *)&ctl->finalised;
        if (fin >= P.n_work) break;
        bool allowed = !(P.spread && s_odd) || (P.n_work - fin > PP.spread_max);
        if (!allowed) {
          // standing back -- unless searches have been WAITING in the queue for ~100 us: more searches are live than there are even
          // SMs, and a search that runs beside a busy neighbour (-25 %) beats one that does not run at all.  The odd CTA gives it
          // back at the end of the quantum (stand_back below) and the even CTAs, which let go of theirs whenever somebody waits,
          // pick it up within that time: the searches rotate through the shared pairs instead of staying there.
          const bool q = (*(volatile int *)&ctl->q_head - *(volatile int *)&ctl->q_tail) < 0;
          if (!q) seen_waiting = false;
          else if (!seen_waiting) { seen_waiting = true; t_wait = clock64(); }
          else if (PP.overflow_odd && clock64() - t_wait > 200000ll) allowed = true;
          if (fin != last) { last = fin; t_last = clock64(); }
          else if (clock64() - t_last > 500000000ll) allowed = true;
        }
        if (allowed) { got = ring_pop(&ctl->q_head, &ctl->q_tail, PP.queue, PP.q_mask); if (got >= 0) break; if (got == -3) { atomicExch(&ctl->error, 1); got = -1; break; } }
        if (clock64() - t_idle > (1ll << 38)) { atomicExch(&ctl->error, 2); got = -1; break; }      // ~2 minutes without work while searches are unfinished: give up (the host reports it)
        __nanosleep(allowed ? 500 : 5000);
      }
      int slot = -1, phase = 0;
      if (got >= 0) {
        const int4 st = __ldcg(reinterpret_cast<const int4 *>(&PP.state[got]));     // phase, slot, status, dij_hn
        phase = st.x; slot = st.y;
        if (phase == 0 && st.z == 0) {
          slot = ring_pop(&ctl->slot_head, &ctl->slot_tail, PP.slot_ring, PP.slot_mask);
          if (slot < 0) {       // every slot is held by a suspended search: this one goes to the back of the queue, a suspended one will come up
            ring_push(&ctl->q_tail, PP.queue, PP.q_mask, got); atomicAdd(&ctl->n_requeues, 1); got = -2; if (PP.phase != 1) __nanosleep(2000);
          }
        }
      }
      s_scen = got; s_slot = slot; s_phase = phase;
    }
    __syncthreads();
    const int sc = s_scen;
    if (sc == -1) break;
    if (sc == -2) { __syncthreads(); continue; }
    const int slot = s_slot;
    const bool fresh = (s_phase == 0);
    Node *nodes = P.nodes + (size_t)(slot < 0 ? 0 : slot) * P.node_cap;
    NodeShot *nshot = P.nshot + (size_t)(slot < 0 ? 0 : slot) * P.node_cap;
    int32_t *htab = P.htab + (size_t)(slot < 0 ? 0 : slot) * P.htab_stride;
    OEnt *oge = P.oheap + (size_t)(slot < 0 ? 0 : slot) * P.node_cap;
#ifndef AVP_SCEN_SMEM          // A/B build -DAVP_SCEN_SMEM: the scenario record copied to shared memory (measured 1 % slower than the L1-resident global record)
    const ScenDev &S = P.scen[sc];
#else
    // the scenario record in shared memory: col_range / lin_at / map_index / the Dijkstra read its fields on every call, and an L1
    // miss there (the nodes, heaps and h tables stream through L1) is a round trip to L2 in the middle of a dependent chain
    {
      const int32_t *src = reinterpret_cast<const int32_t *>(&P.scen[sc]); int32_t *dst = reinterpret_cast<int32_t *>(&s_S);
      for (int i = tid; i < (int)(sizeof(ScenDev) / 4); i += BLOCK) dst[i] = src[i];
    }
    __syncthreads();
    const ScenDev &S = s_S;
#endif
    const double2 *cells = P.cells + S.cell_off;
    const int32_t *col_start = P.col_start + S.col_off;
    const unsigned cell_bytes = (unsigned)S.n_obs * (unsigned)sizeof(double2), col_bytes = (((unsigned)S.nx + 1u) * 4u + 15u) & ~15u;
    const bool staged = slot >= 0 && S.n_obs > 0 && (int)(cell_bytes + col_bytes) <= PP.cell_smem;
    if (staged) {
      if (tid == 0) {
        mbar_expect_tx(&s_cell_bar, cell_bytes + col_bytes);
        bulk_g2s(s_cells, cells, cell_bytes, &s_cell_bar);
        bulk_g2s(s_cells + cell_bytes, col_start, col_bytes, &s_cell_bar);
      }
      cells = reinterpret_cast<const double2 *>(s_cells);
      col_start = reinterpret_cast<const int32_t *>(s_cells + cell_bytes);
    }
    int32_t *hval = P.hval + S.id_off, *ost = P.ost + S.id_off;
    unsigned long long *dgheap = P.dheap + (size_t)sc * P.dheap_cap;
    const double goal[3] = {S.pose[3], S.pose[4], pi_2_pi(S.pose[5])};
    int32_t *pops = P.pops ? P.pops + (size_t)sc * P.cap_pops : nullptr;
    int32_t *hql = P.hq_log ? P.hq_log + (size_t)sc * AVP_HQ_CAP * 3 : nullptr;
    int *dbg = P.dbg ? P.dbg + (size_t)sc * 8 : nullptr;
    const long long t_start = clock64();
    LPROF(unsigned long long gt_take = 0; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt_take));)
#ifdef AVP_PROFILE
    // cycle accumulators: thread 0 (commit warp) and thread 32 (evaluators) each keep their own
    long long pc[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, tp = t_start;
    long long wt = 0;
#define PIPE_TICK(who, k) do { if (tid == (who)) { const long long t_ = clock_ordered(); pc[k] += t_ - tp; tp = t_; } } while (0)
#define WP_START() do { if (lane == 0) wt = clock_ordered(); } while (0)
#define WP_ACC(k) do { if (lane == 0) { const long long t_ = clock_ordered(); s_wp[warp][k] += t_ - wt; wt = t_; } } while (0)
    if (lane == 0) for (int k = 0; k < 8; ++k) s_wp[warp][k] = 0;
    if (tid < 48) s_ic[tid] = 0;
    if (tid < 16) s_subt[tid] = 0;
    long long st_ = 0;
#define SUBT(k) do { if (tid == 0) { const long long t_ = clock_ordered(); s_subt[k] += t_ - st_; st_ = t_; } } while (0)
    long long *tsw = (P.wprof && warp < 16) ? P.wprof + ((size_t)sc * 16 + warp) * 24 + 8 : nullptr;
#define TS(k) do { __syncwarp(); if (lane == 0 && tsw && ((k) < 2 ? (s_npops == P.trace_pop) : s_trace_on)) tsw[k] = clock_ordered(); } while (0)
#else
#define PIPE_TICK(who, k) do { } while (0)
#define WP_START() do { } while (0)
#define WP_ACC(k) do { } while (0)
#define TS(k) do { } while (0)
#define SUBT(k) do { } while (0)
#endif

    // open_list.get() (path_planner.py:70) with the loop's exit tests; lane 0 of the commit warp decides and leaves the size of the
    // heap in s_sift_n, do_pop_warp then runs the heappop's sift (lane 0; the whole warp in the -DAVP_WARP_POP build)
    auto do_pop = [&]() {
      s_sift_n = 0;
      if (dbg) { dbg[0] = 2; dbg[1] = s_npops; dbg[2] = s_D.closed_len; dbg[3] = s_on; }
      if (P.watchdog_cycles > 0 && clock64() - t_start > P.watchdog_cycles && s_status == 0) s_status = AVP_CAPACITY;
      if (s_status != 0 || s_on == 0) { s_ctlA = CTL_EXIT; return; }
      if (s_npops >= cfg.max_pops) { s_status = AVP_CAPACITY; s_ctlA = CTL_EXIT; return; }
      if (s_npops >= s_limit) {
        // the quantum is used up: let go of the scenario if another one waits for an SM, or if this CTA sits on an odd SM id
        // and should stand back by now (an even CTA will take it); else run on without saving anything
        const int fin = *(volatile int *)&ctl->finalised;
        const bool waiting = (*(volatile int *)&ctl->q_head - *(volatile int *)&ctl->q_tail) < 0;
        const bool stand_back = P.spread && s_odd && (P.n_work - fin <= PP.spread_max);
        if (waiting || stand_back || PP.force_yield || PP.phase == 1) { s_status = AVP_PENDING; s_ctlA = CTL_EXIT; return; }
        s_limit = s_npops + PP.quantum;
      }
      const int ret = s_oi[0];
      s_cur = ret;
      if (pops && s_npops < P.cap_pops) {
        pops[s_npops] = ret;
        if (P.pop_fgh) { double *o = P.pop_fgh + ((size_t)sc * P.cap_pops + s_npops) * 3; o[0] = nodes[ret].f; o[1] = nodes[ret].g; o[2] = nodes[ret].h; }
      }
      s_npops++;
      s_sift_n = s_on; s_on = s_on - 1;
      s_ctlA = CTL_RUN;
    };
    auto do_pop_warp = [&](bool go) {           // called by every lane of warp 0
      if (lane == 0 && go) do_pop();
      __syncwarp();
      const int n_before = s_sift_n;
#if defined(AVP_WARP_POP)
      if (n_before > 0) oh_pop_fix_warp<SMO>(s_of, s_oi, oge, nodes, n_before, lane);
#elif defined(AVP_SERIAL_POP)
      if (n_before > 0 && lane == 0) { int n_ = n_before; oh_pop_fix<SMO>(s_of, s_oi, oge, nodes, n_); }
#else
      if (n_before > 0) oh_pop_fix_hybrid<SMO>(s_of, s_oi, oge, nodes, n_before, lane);
#endif
      __syncwarp();
    };

    // ---- take over the scenario: a fresh one starts at the root (the eager Dijkstra is done: k_dij_eager), a suspended one where
    //      it was left.  The shared-memory heads of the two heaps come from their save areas.
    const ScenState st0 = PP.state[sc];
    if (fresh && slot >= 0) for (int i = tid; i < P.htab_size; i += BLOCK) htab[i] = -1;
    for (int i = tid; i < st0.dij_hn && i < AVP_SM_HEAP; i += BLOCK) s_heap[i] = dgheap[i];
    if (!fresh) for (int i = tid; i < st0.on && i < SMO; i += BLOCK) { const int4 v = *reinterpret_cast<const int4 *>(&oge[i]); s_of[i] = __hiloint2double(v.y, v.x); s_oi[i] = v.z; }
    if (tid == 0) {
      s_D.S = &S; s_D.cost = P.cost + S.cost_off; s_D.hval = hval; s_D.ost = ost;
      s_D.gx = P.gx + S.id_off; s_D.gy = P.gy + S.id_off;
      s_D.sheap = s_heap; s_D.gheap = dgheap; s_D.gcap = P.dheap_cap;
      s_D.hn = st0.dij_hn; s_D.closed_len = st0.dij_closed; s_D.status = 0;
      s_on = st0.on; s_G = st0.G; s_nclosed = st0.nclosed; s_npops = st0.npops; s_nhq = st0.nhq; s_nhcalls = st0.nhcalls;
      s_status = st0.status;
      s_cur = st0.cur; s_in_radius = st0.in_radius; s_best_ok = st0.best_ok; s_shot_coll = 0; s_npts = 0; s_best.ok = 0;
      s_res[0].node = -1; s_res[1].node = -1; s_rb = 0; s_do_commit = 0; s_ctlA = CTL_RUN; s_ctlB = CTL_RUN;
      s_best.ct = 0; s_best.n = 0; s_best.inst = -1; s_best.degenerate = 0; s_best.L = 0.0;
      s_limit = st0.npops + PP.quantum;
      if (dbg) { dbg[0] = 1; dbg[1] = st0.npops; }
    }
    __syncthreads();
    if (fresh && s_status == 0) {
      // root node (hybrid_a_star.py:102-112); warp 1 meanwhile solves the root's rs word (the root's theta is a Python float: phi_np = 0)
      if (tid == 0) {
        Node r; r.x = S.pose[0]; r.y = S.pose[1]; r.theta = pi_2_pi(S.pose[2]); r.f = 0; r.g = 0; r.h = 0; r.parent = -1;
        r.forward = 1; r.steer_idx = 0; r.in_open = 1; r.in_closed = 0; r.hpos = 0;
        r.in_radius = sqrt(d_pow2(r.x - goal[0]) + d_pow2(r.y - goal[1])) < cfg.flag_radius;
        nodes[0] = r;
        htab_insert(htab, hmask, nodes, 0);
        { int n_ = s_on; oh_push<SMO>(s_of, s_oi, oge, nodes, n_, 0.0, 0); s_on = n_; }      // an empty heap: position 0 is in shared memory
      } else if (warp == 1) {
        const double q0[3] = {S.pose[0], S.pose[1], pi_2_pi(S.pose[2])};
        RsCand *rc = reinterpret_cast<RsCand *>(&s_cand[0][0]);          // scratch: nothing else uses s_cand before the first evaluation
        RsBest b; rs_length_warp(q0, goal, maxc, 1, 0, rc, b);
        if (lane == 0) {
          NodeShot w; w.t = 0.0; w.u = 0.0; w.v = 0.0; w.L = 0.0; w.inst = -1; w.ok = 0;
          if (b.ok && !b.degenerate) { w.t = rc[b.inst].t; w.u = rc[b.inst].u; w.v = rc[b.inst].v; w.L = b.L; w.inst = b.inst; w.ok = 1; }
          nshot[0] = w;
        }
      }
    }
    __syncthreads();
    if (staged) { mbar_wait(&s_cell_bar, cell_parity); cell_parity ^= 1u; }       // the cell list has landed
    if (warp == 0) { if (lane == 0) s_sift_n = 0; do_pop_warp(true); }     // a fresh search: the first get() returns the root
    PIPE_TICK(0, 0);
    PROF(if (tid == 32) tp = clock_ordered();)

    bool reached = false;
    LPROF(if (tid < 12) s_lp[tid] = 0; if (tid == 32) s_lpe = clock_ordered(); if (tid == 0) s_lpc = clock_ordered();)
    for (;;) {
      __syncthreads();                                 // ---- barrier A: the evaluators' result is complete, the next node is popped
      LPROF(if (tid == 32) { const long long t_ = clock_ordered(); s_lp[0] += t_ - s_lpe; s_lpe = t_; } if (tid == 0) { s_lp[2] += clock_ordered() - s_lpc; s_lp[3] += 1; })
      PIPE_TICK(0, 4);                           // commit warp waiting for the evaluators
      PROF(if (tid == 0) st_ = clock_ordered();)
      TS(0);
      // s_ctlA is written by do_pop (between B and A) and read here; s_ctlB is written between A and B and read
      // after B: no control word is written while another warp may still be reading it
      if (s_ctlA == CTL_EXIT) break;
      if (warp == 0) {
        SUBT(0);
        WP_START();
        const int wb = s_rb ^ 1;
        const int cur = s_cur;
        if (lane == 0) { s_ctlB = CTL_RUN; PROF(s_trace_on = (s_npops == P.trace_pop);) }
        __syncwarp();
        SUBT(1);
        if (s_res[wb].node == cur) {             // the result of the popped node is there
          PureRes &R = s_res[wb];
          const int shot_bad = R.in_radius ? s_shot_bad : 0, shot_coll = R.in_radius ? s_shot_coll : 0;   // of the evaluation just finished
          if (lane < nchild) R.coll[lane] = s_chit[lane];          // sub-step collision flags of the evaluation just finished
          if (lane == 0) { s_rb = wb; s_in_radius = R.in_radius; s_best_ok = R.shot_ok; PROF(pc[5]++;) }
          __syncwarp();
          SUBT(2);
          if (shot_bad) { if (lane == 0) { s_status = (shot_bad == 1) ? AVP_RS_DEGENERATE : AVP_CAPACITY; s_ctlB = CTL_EXIT; } }
          else if (R.in_radius && !shot_coll) { if (lane == 0) s_ctlB = CTL_FINISH; }     // path_planner.py:86-88
          else {
            // ---- lookups, g values, h-table prefetch (hybrid_a_star.py:154-172, :206-222): lanes 0..nchild-1.
            //      Node records and table inserts follow after barrier B (the evaluators do not need them).
            int found = -1, skip = 1, coll = 0; bool in_closed = false, oob_geo = false;
            const Node cn = nodes[cur];
            int hvp = -1;
            if (lane < nchild) {
              const int i = lane;
              const int id = R.hid[i];
              // probe and h value were read ahead by the evaluators (E2) after the table inserts of the commit that ran beside
              // them: no insert and no Dijkstra resume has happened since, except that an h value missing then may be there now
              hvp = R.hv[i];
              if (hvp < 0 && id >= 0) hvp = hval[id];
              const double x_ = R.cpose[i][0], y_ = R.cpose[i][1];
#ifdef AVP_NO_LOOKAHEAD
              found = htab_find(htab, hmask, nodes, x_, y_, th);
#else
              found = R.found[i];
#endif
              in_closed = found >= 0 && nodes[found].in_closed;
              oob_geo = (x_ > S.b[1] || x_ < S.b[0] || y_ > S.b[3] || y_ < S.b[2]);
              coll = R.coll[i];
            }
            SUBT(3);
            // closed_list is re-read for every successor (hybrid_a_star.py:155-163): while it is still empty (the root expansion) an
            // earlier sibling that collides is appended to it (:202) and switches the boundary test on for the later ones.  The first
            // such sibling cannot itself be skipped by the boundary test (nothing is closed before it), so it decides.
            {
              const unsigned appended = __ballot_sync(AVP_FULL_MASK, lane < nchild && !in_closed && found < 0 && coll);
              const int first = appended ? (__ffs(appended) - 1) : 64;
              const bool closed_nonempty = (s_nclosed > 0) || (first < lane);
              skip = (in_closed || (oob_geo && closed_nonempty)) ? 1 : 0;
            }
            if (lane < nchild) {
              const int i = lane;
              const double th = R.cpose[i][2];
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
            SUBT(4);
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
            SUBT(5);
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
          s_tgt = T; s_do_commit = 0; PROF(pc[6]++;)
        }
        // ---- the evaluators' queue for the target just published: counters, flags, header of the result buffer
        //      (the flags of the evaluation just finished were read above)
        __syncwarp();
        SUBT(6);
        if (s_ctlB == CTL_RUN) {
          if (lane < nchild) { s_valid[lane] = 0ull; s_chit[lane] = 0; }
          if (lane < RS_NGROUP) s_gcnt[lane] = 0;
          if (lane == 0) {
            s_nplan = 0; s_npts = 0; s_cstride = 1; s_work = 0; s_work2 = 0; s_work3 = 0; s_rs_done = 0; s_course_rdy = 0; s_q_rdy = 0; s_sub_rdy = 0;
            s_ins_done = 0; s_shot_coll = 0; s_shot_bad = 0;
            PureRes &Wn = s_res[s_rb ^ 1];
            Wn.node = s_tgt.valid ? s_tgt.node : -1; Wn.in_radius = s_tgt.in_radius; Wn.shot_ok = (s_tgt.in_radius && s_tgt.shot.ok) ? 1 : 0;
          }
        }
      }
      if (warp == 0) { SUBT(7); WP_ACC(0); }
      PIPE_TICK(0, 1);                           // accept + lookups + prediction
      SUBT(8);
      TS(1);
      __syncthreads();                                 // ---- barrier B: target published
      LPROF(if (tid == 32) { s_lp[1] += clock_ordered() - s_lpe; })
      LPROF(if (tid == 0) s_lpt = clock_ordered();)
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
        LPROF(if (tid == 0) { const long long t_ = clock_ordered(); s_lp[4] += t_ - s_lpt; s_lpt = t_; })
        if (s_do_commit && s_status == 0) {
          const PureRes &R = s_res[s_rb];
          const int cur = s_cur;
          // sequential commit in slot order (hybrid_a_star.py:154-239).  Lane 0 runs ahead over the successors
          // whose h value is already in the table; a miss resumes the Dijkstra search, which is warp-collective.
          int i = 0, n_miss = 0;
          int on = s_on;
          oh_prefetch_push<SMO>(oge, on, nchild, lane);            // the ancestors of the positions this commit pushes to
#ifndef AVP_SERIAL_COMMIT
          // FAST PATH (no successor misses its h value: all but ~0.3 % of the commits): what is independent per successor -- h, f,
          // the node fields, the counters -- is done by lanes 0..nchild-1 at once; only the heap operations, whose order is the
          // heap's structure, run one after the other (lane 0, slot order).  Same values, same order of heap operations.
          const unsigned miss_m = __ballot_sync(AVP_FULL_MASK, lane < nchild && s_need[lane] && s_hv[lane] < 0);
          if (miss_m == 0u) {
            int act = 0; double fv = 0.0;                        // 1 closed on creation, 2 push, 3 in-place update
            if (lane < nchild && !s_skip[lane]) {
              const int found = s_found[lane];
              if (found < 0 && R.coll[lane]) act = 1;
              else {
                const double h1 = s_h1[lane], h2 = R.rsL[lane];
                const double h = (h2 > h1) ? h2 : h1;                       // max(h_value_1, h_value_2) (:294-296)
                if (found < 0) {                                            // :206-216
                  Node &n = nodes[s_G + lane + 1];
                  fv = s_g[lane] + h;
                  n.h = h; n.f = fv; n.in_open = 1;
                  act = 2;
                } else {                                                    // :219-230 (in place, no re-heapify)
                  fv = h + s_g[lane];
                  if (fv < s_oldf[lane]) {
                    Node &n = nodes[found];
                    n.f = fv; n.g = s_g[lane]; n.h = h; n.parent = cur; n.forward = (lane < nchild / 2.0) ? 1 : 0; n.steer_idx = (uint8_t)(lane % cfg.steering_angle_num);
                    act = 3;
                  } else act = 4;
                }
              }
            }
            const unsigned closed_m = __ballot_sync(AVP_FULL_MASK, act == 1), h_m = __ballot_sync(AVP_FULL_MASK, act >= 2);
            const unsigned push_m = __ballot_sync(AVP_FULL_MASK, act == 2), upd_m = __ballot_sync(AVP_FULL_MASK, act == 3);
            if (lane == 0) { s_nclosed += __popc(closed_m); s_nhcalls += __popc(h_m); }
            unsigned todo = push_m | upd_m;
            while (todo) {
              const int k = __ffs(todo) - 1; todo &= todo - 1;
              const double fk = shfl_d(fv, k);
              if ((push_m >> k) & 1u) {                 // every lane: the sift is warp-collective (oh_siftdown_warp)
                LPROF(long long tl_ = 0; if (lane == 0) tl_ = clock_ordered();)
                oh_push_warp<SMO>(s_of, s_oi, oge, nodes, on, fk, s_G + k + 1, lane);
                LPROF(if (lane == 0) { s_lp[8] += clock_ordered() - tl_; s_lp[9] += 1; })
              } else { if (lane == 0) oh_set_key<SMO>(s_of, oge, nodes[s_found[k]].hpos, fk); __syncwarp(); }
            }
            i = nchild;
          }
#endif
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
                  PROF(const long long t_ = clock64();)
                  LPROF(const long long tl_ = clock_ordered();)
                  oh_push<SMO>(s_of, s_oi, oge, nodes, on, f, child);
                  LPROF(s_lp[8] += clock_ordered() - tl_; s_lp[9] += 1;)
                  PROF(pc[8] += clock64() - t_; pc[9]++;)
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
            PROF(const long long td_ = clock64();)
            LPROF(long long tdl_ = 0; if (tid == 0) tdl_ = clock_ordered();)
            const int d = dij_compute_path(s_D, s_heap, R.cpose[stop][0], R.cpose[stop][1], &term);
            LPROF(if (tid == 0) { const long long t_ = clock_ordered(); s_lp[6] += t_ - tdl_; s_lpt += t_ - tdl_; })
            PROF(if (lane == 0) { pc[10] += clock64() - td_; pc[11]++; })
            ++n_miss;
            if (lane == 0) {
              if (hql && s_nhq < AVP_HQ_CAP) { hql[3 * s_nhq] = (int)term; hql[3 * s_nhq + 1] = d; hql[3 * s_nhq + 2] = s_D.closed_len; }
              s_nhq++;
              if (d < 0) { s_status = s_D.status ? s_D.status : AVP_H_UNREACHABLE; s_nhcalls++; }      // the call was entered (hybrid_a_star.py:261); the reference never returns from it
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
        LPROF(if (tid == 0) { const long long t_ = clock_ordered(); s_lp[5] += t_ - s_lpt; s_lpt = t_; })
        if (lane == 0) s_sift_n = 0;
        do_pop_warp(s_do_commit || s_status != 0);
#ifndef AVP_NO_AB_PREFETCH
        // the serial section after the next barrier A reads the popped node's record and, when the node after it is the heap's new
        // root, that node's record and stored word: both ids are known now, the loads land while the evaluation finishes
        if (lane < 4 && s_ctlA == CTL_RUN) {
          const int a = (lane < 2) ? s_cur : ((s_on > 0) ? s_oi[0] : -1);
          if (a >= 0) { if (lane & 1) prefetch_l1(&nshot[a]); else prefetch_l1(&nodes[a]); }
        }
#endif
        LPROF(if (tid == 0) { const long long t_ = clock_ordered(); s_lp[7] += t_ - s_lpt; s_lpt = t_; })
        WP_ACC(2);
        TS(4);
        PIPE_TICK(0, 3);                         // heappop
        // ---- the commit warp joins the evaluators' queue (initialised before barrier B)
        __syncwarp();
        T = s_tgt;
        if (!T.valid) { LPROF(if (tid == 0) s_lpc = clock_ordered();) continue; }
        WP_START();
      } else {
        // =========================== EVALUATORS ===========================
        T = s_tgt;
        if (!T.valid) { LPROF(if (tid == 32) s_lpe = clock_ordered();) continue; }
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
#ifdef AVP_SUB_FINE
        const int n_items = 4 + RS_NITEM + nchild * nsubs;      // A/B build: one sub-step check per item
#else
        const int n_items = 4 + RS_NITEM + nchild;
#endif
        VehGeom *vg = s_vg[warp];
        for (;;) {
          int it = 0;
          if (lane == 0) it = atomicAdd(&s_work, 1);
          it = __shfl_sync(AVP_FULL_MASK, it, 0);
          if (it >= n_items) break;
          PROF(const long long ti_ = clock64();)
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
                { double st_, ct_; d_sincos(-T.theta, st_, ct_); s_tcs[0] = ct_; s_tcs[1] = st_; }
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
              const int nq = (npts + AVP_COURSE_BATCH - 1) / AVP_COURSE_BATCH;      // the unit of the order is a batch of neighbouring points
              if (nq > 4) { st = (nq * 49 + 64) >> 7; for (;;) { int a = nq, b = st; while (b) { const int t_ = a % b; a = b; b = t_; } if (a == 1) break; ++st; } }
#endif
              s_npts = npts; s_cstride = st; __threadfence_block(); st_release_cta(&s_course_rdy, 1);
            }
          } else if (it == 1) {
#ifdef AVP_QUERY_SERIAL      // A/B build: two sin/cos evaluations one after the other on lanes 0..nchild-1
            if (lane < nchild) {
              const int c = lane;
              double q0[3];
              const double tn = cfg.tan_steer[c % cfg.steering_angle_num];
              const double speed = (c < nchild / 2.0) ? cfg.max_v : -cfg.max_v;
              const double td = speed * cfg.dt;
              q0[2] = pi_2_pi(T.theta + (cfg.max_v * tn) / cfg.lw * cfg.dt);
              double cs, sn; d_sincos(q0[2], sn, cs);
              q0[0] = T.x + td * cs; q0[1] = T.y + td * sn;
              W.cpose[c][0] = q0[0]; W.cpose[c][1] = q0[1]; W.cpose[c][2] = q0[2];
              rs_query_cs(q0, cs, sn, goal, maxc, s_Q[c]);
            }
#else
            // successor c: lane c evaluates sin / cos of its heading, lane 16 + c sin / cos of (goal heading - its heading) at the same
            // time (one call of the same function on 2 * nchild lanes instead of two calls one after the other: the start of every
            // evaluation waits for this item); same arguments, same bits
            {
              const int c = lane & 15, half = lane >> 4;
              double th = 0.0, sv = 0.0, cv = 1.0;
              if (c < nchild) {
                const double tn = cfg.tan_steer[c % cfg.steering_angle_num];
                th = pi_2_pi(T.theta + (cfg.max_v * tn) / cfg.lw * cfg.dt);
                d_sincos(half ? (goal[2] - th) : th, sv, cv);
              }
              const double sp = shfl_d(sv, (lane & 15) + 16), cp = shfl_d(cv, (lane & 15) + 16);
              if (lane < nchild) {
                double q0[3];
                const double speed = (c < nchild / 2.0) ? cfg.max_v : -cfg.max_v;
                const double td = speed * cfg.dt;
                q0[2] = th; q0[0] = T.x + td * cv; q0[1] = T.y + td * sv;
                W.cpose[c][0] = q0[0]; W.cpose[c][1] = q0[1]; W.cpose[c][2] = q0[2];
                rs_query_cs2(q0, cv, sv, goal, maxc, sp, cp, s_Q[c]);
              }
            }
#endif
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
              double cs, sn; d_sincos(th_i, sn, cs);
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
#ifdef AVP_RS_FINE
          } else if (it >= 4 + nchild) {
            const int rs_it = it - 4 - nchild;
#else
          } else if (it <= 3 + RS_NITEM) {
            const int rs_it = it - 4;
#endif
            const int k = lane / nchild, row = lane - k * nchild;
            const int inst = (k < 3) ? rs_item_inst[rs_it][k] : -1;
#ifdef AVP_RS_OWNQ
            // A/B build: every rs item derives the normalised query of its rows itself (the same operations as item 1, so the same bits)
            // instead of waiting for item 1 to publish them: the eleven warps that start with an rs item do not idle through item 1
            RsQuery Qown;
            if (inst >= 0) {
              double q0[3];
              const double tn = cfg.tan_steer[row % cfg.steering_angle_num];
              const double speed = (row < nchild / 2.0) ? cfg.max_v : -cfg.max_v;
              const double td = speed * cfg.dt;
              q0[2] = pi_2_pi(T.theta + (cfg.max_v * tn) / cfg.lw * cfg.dt);
              double cs, sn; d_sincos(q0[2], sn, cs);
              q0[0] = T.x + td * cs; q0[1] = T.y + td * sn;
              rs_query_cs(q0, cs, sn, goal, maxc, Qown);
            }
            const RsQuery &Qrow = Qown;
#else
            if (!wait_ge_cta(&s_q_rdy, 1)) { if (lane == 0) s_status = AVP_CAPACITY; break; }
            const RsQuery &Qrow = s_Q[row];
#endif
            if (inst >= 0) {
              double t, u, v;
              if (rs_eval_instance(inst, Qrow, t, u, v)) {
                RsCandX &c = s_cand[row][inst];                       // arranged once, here (rs_curve.py:200-534), and kept for the selection
                double l[5] = {0.0, 0.0, 0.0, 0.0, 0.0}, a[5]; int ct; unsigned mask;
                const int n = rs_arrange_inl(inst, t, u, v, 1, 1, l, ct, mask);
#pragma unroll
                for (int q = 0; q < 5; ++q) { a[q] = fabs(l[q]); if (q < n) c.len[q] = l[q]; }
                c.t = t; c.u = u; c.v = v; c.n = n; c.ct = ct; c.mask = mask;
                c.L = py_sum(a, n, mask);                              // rs_curve.py:148
                atomicOr(&s_valid[row], 1ull << inst);
              }
            }
            __syncwarp();
#ifdef AVP_SEL_GROUPED       // A/B build (measured +6 % step time: 10 lanes per group selection instead of 22, and the selection code a second time in the hot path)
            // set_path + calc_optimal_path per ctype group (rs_curve.py:137-156, :99-110) as soon as the group's last instance is there:
            // the warp that completes a group (a counter per group) selects its winner for every successor (lanes = successors).  The
            // selection used to follow ALL rs items (11 k cycles at the end of the evaluation, most warps idle); what is left for
            // the end is the combination of the 11 group winners per successor.
            {
              unsigned sel = 0u;
              if (lane == 0) {
                __threadfence_block();
                int gprev = -1, cnt = 0;
                for (int k2 = 0; k2 <= 3; ++k2) {
                  const int in2 = (k2 < 3) ? rs_item_inst[rs_it][k2] : -1;
                  int g2 = -1;
                  if (in2 >= 0) { g2 = 0; while (in2 >= rs_grp_begin[g2 + 1]) ++g2; }
                  if (g2 != gprev) {
                    if (gprev >= 0) { const int old = atomicAdd(&s_gcnt[gprev], cnt); if (old + cnt == rs_grp_begin[gprev + 1] - rs_grp_begin[gprev]) sel |= 1u << gprev; }
                    gprev = g2; cnt = 0;
                  }
                  if (g2 >= 0) ++cnt;
                }
                __threadfence_block();
              }
              sel = __shfl_sync(AVP_FULL_MASK, sel, 0);
              __syncwarp();
              while (sel) {
                const int g = __ffs(sel) - 1; sel &= sel - 1;
                if (lane < nchild) rs_select_group_x(s_cand[lane], *(volatile unsigned long long *)&s_valid[lane], g, maxc, s_grp[lane][g]);
              }
              __syncwarp();
            }
#endif
            if (lane == 0) { __threadfence_block(); add_release_cta(&s_rs_done, 1); }
          } else {
            if (!wait_ge_cta(&s_sub_rdy, 1)) { if (lane == 0) s_status = AVP_CAPACITY; break; }
#ifdef AVP_RS_FINE
            const int i = it - 4;
#else
            const int i = it - 4 - RS_NITEM;
#endif
#if !defined(AVP_SUB_COARSE) && !defined(AVP_SUB_FINE)
            // the sub-steps of successor i in ONE pass over the cell list (check_distance_multi_sm): their rectangles are computed side by
            // side by lanes 4p .. 4p+3, any hit decides (hybrid_a_star.py:185-204 breaks at the first one; the flag is an OR)
            {
              int coll = 0;
              if (cfg.collision_mode == 1) {
                for (int k = nsubs - 1; k >= 0 && !coll; --k) {
                  const double tn = cfg.tan_steer[i % cfg.steering_angle_num];
                  const double speed = (i < nchild / 2.0) ? cfg.max_v : -cfg.max_v;
                  const double td_i = speed * cfg.ddt * (k + 1);
                  const double th_i = pi_2_pi(T.theta + (cfg.max_v * tn) / cfg.lw * cfg.ddt * (k + 1));
                  double cs, sn; d_sincos(th_i, sn, cs);
                  coll = check_circle_warp(cfg, S, cells, T.x + td_i * cs, T.y + td_i * sn, cs, sn) ? 1 : 0;
                }
              } else {
                const VehDims vd = veh_dims(cfg);
                for (int k0 = 0; k0 < nsubs && !coll; k0 += AVP_MULTI_POSE) {
                  const int np = (nsubs - k0 < AVP_MULTI_POSE) ? nsubs - k0 : AVP_MULTI_POSE;
                  const int k = k0 + (lane >> 2);
                  double px = 0.0, py = 0.0, pcs = 1.0, psn = 0.0;
                  if ((lane >> 2) < np) {
                    if (k < 4) { px = s_sub[i][k][0]; py = s_sub[i][k][1]; pcs = s_sub[i][k][2]; psn = s_sub[i][k][3]; }
                    else {
                      const double tn = cfg.tan_steer[i % cfg.steering_angle_num];
                      const double speed = (i < nchild / 2.0) ? cfg.max_v : -cfg.max_v;
                      const double td_i = speed * cfg.ddt * (k + 1);
                      const double th_i = pi_2_pi(T.theta + (cfg.max_v * tn) / cfg.lw * cfg.ddt * (k + 1));
                      d_sincos(th_i, psn, pcs);
                      px = T.x + td_i * pcs; py = T.y + td_i * psn;
                    }
                  }
                  coll = check_distance_multi_sm(vd, S, cells, col_start, np, px, py, pcs, psn, vg) ? 1 : 0;
                }
              }
              if (lane == 0) s_chit[i] = coll;
            }
#elif defined(AVP_SUB_FINE)
            // one (successor, sub-step) pair per item, the farthest sub-steps of every successor first (any hit decides,
            // hybrid_a_star.py:185-204 breaks at the first one; the parent pose is collision free, so the sub-step next to it is the
            // least likely to hit); a successor that has its hit is not checked again
            {
              const int j = i, kk0 = j / nchild, ii = j - kk0 * nchild, k = nsubs - 1 - kk0;
              const int known = __shfl_sync(AVP_FULL_MASK, *(volatile int *)&s_chit[ii], 0);
              if (!known) {
                bool hit;
                if (k < 4) hit = check_pose_cs_warp_sm(cfg, S, cells, col_start, s_sub[ii][k][0], s_sub[ii][k][1], s_sub[ii][k][2], s_sub[ii][k][3], vg);
                else {
                  const double tn = cfg.tan_steer[ii % cfg.steering_angle_num];
                  const double speed = (ii < nchild / 2.0) ? cfg.max_v : -cfg.max_v;
                  const double td_i = speed * cfg.ddt * (k + 1);
                  const double th_i = pi_2_pi(T.theta + (cfg.max_v * tn) / cfg.lw * cfg.ddt * (k + 1));
                  double cs, sn; d_sincos(th_i, sn, cs);
                  hit = check_pose_cs_warp_sm(cfg, S, cells, col_start, T.x + td_i * cs, T.y + td_i * sn, cs, sn, vg);
                }
                if (hit && lane == 0) s_chit[ii] = 1;
              }
            }
#else
            int coll = 0;
            for (int kk = 0; kk < nsubs; ++kk) {
              // any hit decides (hybrid_a_star.py:185-204 breaks at the first one): the sub-steps are checked farthest first -- the parent pose
              // is collision free, so the sub-step next to it is the least likely to hit
#ifdef AVP_SUB_FORWARD
              const int k = kk;
#else
              const int k = nsubs - 1 - kk;
#endif
              bool hit;
              if (k < 4) hit = check_pose_cs_warp_sm(cfg, S, cells, col_start, s_sub[i][k][0], s_sub[i][k][1], s_sub[i][k][2], s_sub[i][k][3], vg);
              else {
                const double tn = cfg.tan_steer[i % cfg.steering_angle_num];
                const double speed = (i < nchild / 2.0) ? cfg.max_v : -cfg.max_v;
                const double td_i = speed * cfg.ddt * (k + 1);
                const double th_i = pi_2_pi(T.theta + (cfg.max_v * tn) / cfg.lw * cfg.ddt * (k + 1));
                double cs, sn; d_sincos(th_i, sn, cs);
                hit = check_pose_cs_warp_sm(cfg, S, cells, col_start, T.x + td_i * cs, T.y + td_i * sn, cs, sn, vg);
              }
              if (hit) { coll = 1; break; }
            }
            if (lane == 0) s_chit[i] = coll;
#endif
          }
#ifdef AVP_PROFILE
          if (lane == 0 && it < 40) atomicAdd(reinterpret_cast<unsigned long long *>(&s_ic[it]), (unsigned long long)(clock64() - ti_));
#endif
        }
        if (warp != 0) { WP_ACC(1); TS(5); WP_START(); PIPE_TICK(32, 12); }
        // ---- E2: the items that consume other items' results:
        //   course point checks (hybrid_a_star.py:334-347), after s_course_rdy, in strided order, skipped once one of them hit
        //   word selection, one successor per warp item (lanes 0..10 = ctype groups; -DAVP_SEL_PAIRED: two, lanes 0..10 / 16..26): set_path de-duplication + minimum
        //   per group, then calc_optimal_path (combine the groups) and the word kept for the successor's own shot; after every
        //   rs item is finished (s_rs_done).  A warp takes a selection whenever the rs items are finished, else a course point.
        // the commit warp only helps while E1 items are left: it arrives late, and a selection or a course point taken then
        // would make it the last warp at barrier A
        if (warp == 0) { WP_ACC(3); LPROF(if (tid == 0) s_lpc = clock_ordered();) continue; }
        if (!wait_ge_cta(&s_course_rdy, 1)) { if (lane == 0) s_status = AVP_CAPACITY; continue; }
        const int npts = s_npts, cstride = s_cstride;
        const int nbat = (npts + AVP_COURSE_BATCH - 1) / AVP_COURSE_BATCH;
#if defined(AVP_SEL_GROUPED)
        const int n_sel = 1;                        // the group winners are there (E1): one item combines them, lanes = successors
#elif !defined(AVP_SEL_PAIRED)      // one successor per selection item (lanes = ctype groups); -DAVP_SEL_PAIRED: two successors per item (measured 2 % slower on C2 together with the 2048-entry heap head)
        const int n_sel = nchild;
#else
        const int n_sel = (nchild + 1) / 2;
#endif
        bool sel_left = true, chk_left = npts > 0;
        for (;;) {
          PROF(const long long ti_ = clock64();)
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
            if (it >= nbat) { chk_left = false; continue; }
            kind = 1;
          } else break;
          if (kind == 1) {
            const int bq = (int)(((long long)it * cstride) % nbat);
            const int stop = __shfl_sync(AVP_FULL_MASK, *(volatile int *)&s_shot_coll, 0);   // warp-uniform early exit
            if (stop) { chk_left = false; continue; }
            // AVP_COURSE_BATCH neighbouring course points in one pass over the cell list (lanes 4p .. 4p+3: point p of the batch)
            const int j0 = bq * AVP_COURSE_BATCH, np = (npts - j0 < AVP_COURSE_BATCH) ? npts - j0 : AVP_COURSE_BATCH;
            const int j = j0 + (lane >> 2);
            double gx_ = 0.0, gy_ = 0.0, gsn = 0.0, gcs = 1.0;
            if ((lane >> 2) < np) {
              const double ix = CX[j], iy = CY[j], cm = s_tcs[0], sm = s_tcs[1];
              gx_ = cm * ix + sm * iy + T.x; gy_ = -sm * ix + cm * iy + T.y;      // rs_curve.py:124-130
              const double gyaw = pi_2_pi(CYAW[j] + T.theta);
              const double gth = pi_2_pi(gyaw);
              d_sincos(gth, gsn, gcs);
            }
            bool chit = false;
            if (cfg.collision_mode == 1) {
              for (int q = 0; q < np && !chit; ++q)
                chit = check_circle_warp(cfg, S, cells, shfl_d(gx_, 4 * q), shfl_d(gy_, 4 * q), shfl_d(gcs, 4 * q), shfl_d(gsn, 4 * q));
            }
#if AVP_COURSE_BATCH == 1 && defined(AVP_COURSE_SINGLE)     // A/B build: the single-pose function (a second copy of the cell predicate in the hot code: +4 % step time)
            else chit = check_distance_warp_sm(veh_dims(cfg), S, cells, col_start, shfl_d(gx_, 0), shfl_d(gy_, 0), shfl_d(gcs, 0), shfl_d(gsn, 0), vg);
#else
            else chit = check_distance_multi_sm(veh_dims(cfg), S, cells, col_start, np, gx_, gy_, gcs, gsn, vg);
#endif
            if (chit) { if (lane == 0) s_shot_coll = 1; }
          } else {
#if defined(AVP_SEL_GROUPED)
            const int gl = 0, i = lane;
#elif !defined(AVP_SEL_PAIRED)      // one successor per selection item (lanes = ctype groups); -DAVP_SEL_PAIRED: two successors per item (measured 2 % slower on C2 together with the 2048-entry heap head)
            const int gl = lane, i = (lane < 16) ? it : nchild;
#else
            const int half = lane >> 4, gl = lane & 15, i = 2 * it + half;
#endif
#ifndef AVP_SEL_GROUPED
            if (i < nchild && gl < RS_NGROUP) rs_select_group_x(s_cand[i], s_valid[i], gl, maxc, s_grp[i][gl]);
            __syncwarp();
#endif
            if (i < nchild && gl == 0) {
              // calc_optimal_path over the group winners in order (rs_curve.py:99-110: the last word with L <= min wins)
              int bi = -1, degenerate = 0; double minL = 0.0;
              for (int g = 0; g < RS_NGROUP; ++g) {
                degenerate |= s_grp[i][g].degenerate;
                if (s_grp[i][g].inst < 0) continue;
                if (bi < 0 || s_grp[i][g].Lm <= minL) { bi = s_grp[i][g].inst; minL = s_grp[i][g].Lm; }
              }
              const int ok = (bi >= 0 && !degenerate) ? 1 : 0;
              W.rsok[i] = ok;
              NodeShot w; w.t = 0.0; w.u = 0.0; w.v = 0.0; w.L = 0.0; w.inst = -1; w.ok = 0;
              if (bi >= 0) { const RsCandX &c = s_cand[i][bi]; W.rsL[i] = c.L / maxc; w.t = c.t; w.u = c.u; w.v = c.v; w.L = c.L; w.inst = bi; w.ok = ok; }
              else W.rsL[i] = 0.0;
              W.shot[i] = w;
            }
          }
#ifdef AVP_PROFILE
          if (lane == 0) { const int k_ = 40 + kind; atomicAdd(reinterpret_cast<unsigned long long *>(&s_ic[k_]), (unsigned long long)(clock64() - ti_)); atomicAdd(reinterpret_cast<unsigned long long *>(&s_ic[k_ + 2]), 1ull); }
#endif
        }
        WP_ACC(2); TS(7); PIPE_TICK(32, 13);
        LPROF(if (tid == 32) s_lpe = clock_ordered();)
      }
    }
    __syncthreads();
    LPROF(if (tid < 10 && P.prof) P.prof[(size_t)sc * 16 + tid] += s_lp[tid];)
    LPROF(if (tid == 0 && P.prof) { long long *o = P.prof + (size_t)sc * 16; unsigned long long gt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
                                     o[10] += clock64() - t_start;                 // cycles this scenario held an SM
                                     if (st0.quanta == 0 && fresh) o[11] = (long long)gt_take;        // wall clock (ns) of its first take-over
                                     o[12] = (long long)gt;                        // ... of its latest release (the finish, in the end)
                                     o[13] += 1; })
#ifdef AVP_PROFILE
    if (lane == 0 && P.wprof && warp < 16) { long long *o = P.wprof + ((size_t)sc * 16 + warp) * 24; for (int k = 0; k < 8; ++k) o[k] += s_wp[warp][k]; }
    if (tid < 48 && P.wprof) P.wprof[((size_t)sc * 16 + (tid >> 3)) * 24 + 16 + (tid & 7)] += s_ic[tid];
    if (tid < 16 && P.wprof) P.wprof[((size_t)sc * 16 + 8 + (tid >> 3)) * 24 + 16 + (tid & 7)] += s_subt[tid];
    if (tid == 32 && P.prof) { long long *o = P.prof + (size_t)sc * 16; o[7] += pc[7]; o[12] += pc[12]; o[13] += pc[13]; { unsigned sm_; asm("mov.u32 %0, %%smid;" : "=r"(sm_)); o[14] = ((long long)sm_ << 40) | (t_start & 0xffffffffffll); } o[15] += pc[15]; }
    if (tid == 0 && P.prof) { long long *o = P.prof + (size_t)sc * 16; for (int k = 0; k < 7; ++k) o[k] += pc[k]; for (int k = 8; k < 12; ++k) o[k] += pc[k]; }
#endif

    if (s_status == AVP_PENDING) {
      // ---- let go of the scenario: heap heads back to their save areas, counters to ScenState, the id to the tail of the run queue
      for (int i = tid; i < s_D.hn && i < AVP_SM_HEAP; i += BLOCK) dgheap[i] = s_heap[i];
      for (int i = tid; i < s_on && i < SMO; i += BLOCK) { int4 v; v.x = __double2loint(s_of[i]); v.y = __double2hiint(s_of[i]); v.z = s_oi[i]; v.w = 0; *reinterpret_cast<int4 *>(&oge[i]) = v; }
      __threadfence();
      __syncthreads();
      if (tid == 0) {
        ScenState st; st.phase = 1; st.slot = slot; st.status = 0; st.dij_hn = s_D.hn; st.dij_closed = s_D.closed_len;
        st.nhq = s_nhq; st.nhcalls = s_nhcalls; st.G = s_G; st.nclosed = s_nclosed; st.npops = s_npops; st.on = s_on; st.cur = s_cur;
        st.in_radius = s_in_radius; st.best_ok = s_best_ok; st.quanta = st0.quanta + 1; st.pad = 0;
        PP.state[sc] = st;
        atomicAdd(&ctl->n_suspends, 1);
        ring_push(&ctl->q_tail, PP.queue, PP.q_mask, sc);
      }
      __syncthreads();
      continue;
    }

    // ---- finish: summary + finish_path (hybrid_a_star.py:351-389) + rs tail (path_planner.py:100-108)
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
      if (s_cur >= 0 && slot >= 0) { R.last_pose[0] = nodes[s_cur].x; R.last_pose[1] = nodes[s_cur].y; R.last_pose[2] = nodes[s_cur].theta; }
      else { R.last_pose[0] = 0.0; R.last_pose[1] = 0.0; R.last_pose[2] = 0.0; }
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
      PP.state[sc].phase = 2;
      if (slot >= 0) ring_push(&ctl->slot_tail, PP.slot_ring, PP.slot_mask, slot);
      __threadfence();
      atomicAdd(&ctl->finalised, 1);
    }
    __syncthreads();
  }
#undef PIPE_TICK
#undef WP_START
#undef WP_ACC
#undef TS
#undef SUBT
}
