// avp_pipe_defs.cuh -- the pipelined pop of the search kernel (k_plan, avp_plan.cuh): description and shared definitions.
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
// Warp 0 (the COMMIT warp) runs the
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
#if !defined(AVP_RS_BYCOST) && !defined(AVP_RS_FINE)
// The items of one word family next to each other in the queue (two warps run the same formula code at the same time: the
// instruction lines they fetch are shared), the left-over fourth instance of each family as an item of its own.  Measured against
// the cost-ordered 19-item table below (-DAVP_RS_BYCOST): C3 -2.3 %, C2 within the noise.
#define RS_NITEM 23
__device__ __constant__ int8_t rs_item_inst[RS_NITEM][3] = {
  {6, 7, 8}, {9, -1, -1}, {26, 27, 28}, {29, -1, -1}, {34, 35, 36}, {37, -1, -1}, {30, 31, 32}, {33, -1, -1}, {38, 39, 40}, {41, -1, -1},
  {42, 43, 44}, {45, -1, -1}, {10, 11, 12}, {13, -1, -1}, {14, 15, 16}, {17, -1, -1}, {2, 3, 4}, {5, -1, -1},
  {20, 21, -1}, {18, 19, -1}, {22, 23, -1}, {24, 25, -1}, {0, 1, -1}};
#elif !defined(AVP_RS_FINE)
#define RS_NITEM 19
__device__ __constant__ int8_t rs_item_inst[RS_NITEM][3] = {
  {45, 13, 17}, {0, 1, -1}, {6, 7, 8}, {5, 33, 41}, {20, 21, -1}, {22, 23, -1}, {10, 11, 12}, {14, 15, 16}, {26, 27, 28},
  {34, 35, 36}, {24, 25, -1}, {9, -1, -1}, {29, -1, -1}, {37, -1, -1}, {2, 3, 4}, {38, 39, 40}, {30, 31, 32}, {42, 43, 44}, {18, 19, -1}};
#else
// A/B variant: one word instance x all successors per warp item (10 of 32 lanes, no divergence between instances: the item is as
// long as ONE formula), the sub-step chains (the longest items) first, one selection per successor: a shorter critical path
// through the evaluation for the same work
#define RS_NITEM 46
__device__ __constant__ int8_t rs_item_inst[RS_NITEM][3] = {
  {45, -1, -1}, {13, -1, -1}, {17, -1, -1}, {0, -1, -1}, {1, -1, -1}, {6, -1, -1}, {7, -1, -1}, {8, -1, -1}, {5, -1, -1}, {33, -1, -1}, {41, -1, -1},
  {20, -1, -1}, {21, -1, -1}, {22, -1, -1}, {23, -1, -1}, {10, -1, -1}, {11, -1, -1}, {12, -1, -1}, {14, -1, -1}, {15, -1, -1}, {16, -1, -1},
  {26, -1, -1}, {27, -1, -1}, {28, -1, -1}, {34, -1, -1}, {35, -1, -1}, {36, -1, -1}, {24, -1, -1}, {25, -1, -1}, {9, -1, -1}, {29, -1, -1}, {37, -1, -1},
  {2, -1, -1}, {3, -1, -1}, {4, -1, -1}, {38, -1, -1}, {39, -1, -1}, {40, -1, -1}, {30, -1, -1}, {31, -1, -1}, {32, -1, -1}, {42, -1, -1}, {43, -1, -1}, {44, -1, -1},
  {18, -1, -1}, {19, -1, -1}};
#endif

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
