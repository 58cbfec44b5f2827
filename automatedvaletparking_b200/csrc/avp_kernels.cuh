// avp_kernels.cuh -- kernels of the hybrid-A* hot path (sm_100a).
//
//   k_raster / k_count_cols / k_fill_cells : Map.detect_obstacle_edge (costmap.py:197-261) and the
//                                            np.where(cost_map == 255) cell list (collision_check.py:55-57)
//   k_check_batch                          : distance_checker / two_circle_checker .check
//   k_expand_pure                          : pure part of hybrid_a_star.expand_node + rs length
//   k_rs_optimal                           : rs_curve.calc_optimal_path
// plus the containers (heapq emulations, exact-pose table) and the Dijkstra of the search kernels in avp_plan.cuh.
#pragma once
#include "avp_dev.cuh"

#ifndef AVP_SM_HEAP
#define AVP_SM_HEAP 1024      // Dijkstra heap entries kept in shared memory (the rest spills to L2/HBM)
#endif
#ifndef AVP_SM_OPEN
#define AVP_SM_OPEN 1024      // open-list heap entries kept in shared memory
#endif
#define AVP_COURSE_CAP 1024
#define AVP_HQ_CAP 256
#ifndef AVP_NCHILD_MAX
#define AVP_NCHILD_MAX 10      // 2 * steering_angle_num supported by the search kernel's shared-memory layout
#endif

struct __align__(16) Node {
  double x, y, theta, f, g, h;
  int32_t parent;
  uint8_t forward, steer_idx, in_open, in_closed;
  int32_t hpos;            // position of this node's entry in the open heap (valid while in_open)
  int32_t in_radius;       // distance to the goal < flag_radius (hybrid_a_star.py:308-310), evaluated when the node is created
};
static_assert(sizeof(Node) == 64, "Node must be 64 bytes");

// The rs word selected by calc_optimal_path(pose -> goal) when a node was scored (calc_node_heuristic,
// hybrid_a_star.py:286-292).  try_rs_curve (:326-332) repeats exactly that call when the node is popped,
// so the pipelined search kernel keeps the word instead of solving it again.
struct __align__(8) NodeShot { double t, u, v, L; int32_t inst; int32_t ok; };

struct KParams {
  avp_config cfg;
  int n_scen;
  const ScenDev *scen;
  const uint8_t *cost;
  const double2 *cells;
  const int32_t *col_start;
  // per-id arrays, offset by ScenDev.id_off
  int32_t *hval;   // distance of the first closedlist entry with this grid id, -1 = none
  int32_t *ost;    // -1 unseen, -2 popped, >= 0 current distance while in the open heap
  double *gx, *gy; // lattice coordinates of the Grid object holding this id
  unsigned long long *dheap; int dheap_cap;   // Dijkstra queue, dheap_cap entries per SCENARIO
  // per-slot workspaces (a slot belongs to one search from its first pop to its last)
  Node *nodes; int node_cap;
  NodeShot *nshot;                // node_cap per slot
  struct OEnt *oheap;             // open heap (node_cap entries per slot; [0, SMO) = save area of the shared-memory head)
  int32_t *htab; int htab_size;   // power of two
  int htab_stride;                // entries between two slots' tables
  double *course;                 // 3*AVP_COURSE_CAP doubles per CTA (scratch)
  int32_t *course_dir;
  unsigned char *cand_scratch;      // per CTA: the word candidates of the successors (AVP_CAND_SMEM bytes) for CTAs too narrow to keep them in shared memory
  // results per scenario
  avp_plan_summary *sums;
  double *paths; int cap_path;
  int32_t *pops; int cap_pops;
  double *pop_fgh;                // n * cap_pops * 3: f, g, h of every popped node at the time of its pop (avp_trace_fgh), may be NULL
  int32_t *hq_log;                // n * AVP_HQ_CAP * 3, may be NULL
  int *work_counter;
  const int32_t *work_list;       // processing order (longest start-goal distance first); NULL: 0..n_work-1
  int n_work;
  long long *prof;                // n * 16 SM-cycle accumulators / counters per scenario (thread 0): phases of the main loop, may be NULL
  int trace_pop;                  // pipelined kernel: the pop whose per-warp timeline is recorded in wprof (development aid)
  int spread;                     // the grid covers every SM and the CTAs on odd SM ids stand back while few scenarios are live (avp_plan.cuh)
  long long *wprof;               // n * 16 * 24: per warp (16) and phase (8) work cycles of the pipelined kernel, may be NULL
  int *dbg;                       // n * 8 ints of progress checkpoints (development aid), may be NULL
  long long watchdog_cycles;      // 0 = off; a scenario running longer aborts with AVP_CAPACITY
};

// ------------------------------------------------------------------------------------------
// rasterisation

// np.add.reduce over <= AVP_MAX_VERT strided doubles (pairwise_sum, loops_utils.h)
__device__ __forceinline__ double np_sum_dev(const double *a, int n) {
  if (n < 8) { double r = -0.0; for (int i = 0; i < n; ++i) r += a[i]; return r; }
  double r[8]; int i;
  for (i = 0; i < 8; ++i) r[i] = a[i];
  for (i = 8; i < n - (n % 8); i += 8) for (int j = 0; j < 8; ++j) r[j] += a[i + j];
  double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
  for (; i < n; ++i) res += a[i];
  return res;
}

// one CTA per scenario; each thread rasterises whole polygons (costmap.py:203-261)
__global__ void k_raster(int n_scen, ScenDev *scen, const int32_t *nv, const int32_t *vert_off, const double *verts,
                         uint8_t *cost) {
  const int s = blockIdx.x;
  if (s >= n_scen) return;
  ScenDev &S = scen[s];
  uint8_t *cm = cost + S.cost_off;
  for (int o = S.obs_begin + threadIdx.x; o < S.obs_end; o += blockDim.x) {
    int n0 = nv[o];
    if (n0 > AVP_MAX_VERT) { S.raster_error = 2; continue; }
    double px[AVP_MAX_VERT], py[AVP_MAX_VERT], ang[AVP_MAX_VERT];
    int ord[AVP_MAX_VERT];
    const double *v = verts + 2 * (size_t)vert_off[o];
    for (int i = 0; i < n0; ++i) { px[i] = v[2 * i]; py[i] = v[2 * i + 1]; }
    // np.unique(axis=0): lexicographic sort + drop duplicate rows (costmap.py:206)
    for (int i = 1; i < n0; ++i) {
      double tx = px[i], ty = py[i]; int j = i;
      while (j > 0 && (px[j - 1] > tx || (px[j - 1] == tx && py[j - 1] > ty))) { px[j] = px[j - 1]; py[j] = py[j - 1]; --j; }
      px[j] = tx; py[j] = ty;
    }
    int n = 0;
    for (int i = 0; i < n0; ++i)
      if (n == 0 || px[i] != px[n - 1] || py[i] != py[n - 1]) { px[n] = px[i]; py[n] = py[i]; ++n; }
    const double cx = np_sum_dev(px, n) / n, cy = np_sum_dev(py, n) / n;           // :210-211
    for (int i = 0; i < n; ++i) { ang[i] = d_atan2(py[i] - cy, px[i] - cx) + AVP_PI; ord[i] = i; }   // :215
    for (int i = 1; i < n; ++i) { int t = ord[i], j = i; while (j > 0 && ang[ord[j - 1]] > ang[t]) { ord[j] = ord[j - 1]; --j; } ord[j] = t; }
    for (int j = 0; j < n; ++j) {
      const int a = ord[j], b = ord[(j + 1 == n) ? 0 : j + 1];
      const double p1x = px[a], p1y = py[a];
      const double vx = px[b] - p1x, vy = py[b] - p1y;
      const double ra = d_atan2(vy, vx), c = d_cos(ra), sn = d_sin(ra);             // :229-232
      const double len = __fma_rn(c, vx, sn * vy);     // np.dot(rotation_matrix, v)[0] (BLAS gemv)
      const int points_num = (int)floor(len / S.dx);                                   // :240-241
      const double lstep = (points_num > 1) ? len / (points_num - 1) : 0.0;
      for (int k = 0; k < points_num; ++k) {
        double pxk;
        if (points_num == 1) pxk = 0.0 * len + 0.0;
        else if (k == points_num - 1) pxk = len;
        else pxk = (lstep == 0.0) ? ((double)k / (points_num - 1)) * len + 0.0 : (double)k * lstep + 0.0;
        const double ox = c * pxk + p1x, oy = sn * pxk + p1y;                          // :246-251
        int ixm = -1, iym = -1, nxm = 0, nym = 0;
        int i0 = (int)floor((ox - S.b[0]) / S.stepx);
        for (int i = max(i0 - 3, 0); i <= min(i0 + 3, S.nx - 1); ++i) {
          const double xi = lin_at(S.b[0], S.b[1], S.stepx, S.nx, i);
          if (xi < ox && xi > ox - S.dx) { if (!nxm) ixm = i; ++nxm; }               // :253-254
        }
        i0 = (int)floor((oy - S.b[2]) / S.stepy);
        for (int i = max(i0 - 3, 0); i <= min(i0 + 3, S.ny - 1); ++i) {
          const double yi = lin_at(S.b[2], S.b[3], S.stepy, S.ny, i);
          if (yi < oy && yi > oy - S.dy) { if (!nym) iym = i; ++nym; }               // :256-257
        }
        if (nxm > 0 && nym > 0) {
          if (nxm > 1 || nym > 1) S.raster_error = 1;
          cm[(size_t)ixm * S.ny + iym] = 255;                                          // :259-261
        }
      }
    }
  }
}

// per scenario: col_start[ix] = number of obstacle cells in columns < ix (np.where order)
__global__ void k_count_cols(int n_scen, ScenDev *scen, const uint8_t *cost, int32_t *col_start) {
  const int s = blockIdx.x;
  if (s >= n_scen) return;
  ScenDev &S = scen[s];
  const uint8_t *cm = cost + S.cost_off;
  int32_t *cs = col_start + S.col_off;
  for (int ix = threadIdx.x; ix < S.nx; ix += blockDim.x) {
    int c = 0;
    const uint8_t *row = cm + (size_t)ix * S.ny;
    for (int iy = 0; iy < S.ny; ++iy) c += (row[iy] == 255);
    cs[ix + 1] = c;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int acc = 0; cs[0] = 0;
    for (int ix = 0; ix < S.nx; ++ix) { acc += cs[ix + 1]; cs[ix + 1] = acc; }
    S.n_obs = acc;
  }
}

// cell_off / cell_cap of every scenario = exclusive prefix sum of the (even-padded) obstacle counts, on the device (one CTA);
// total[0] = entries needed, total[1] = 1 if that exceeds the capacity of the cell list (the host then grows it and repeats)
__global__ void __launch_bounds__(1024) k_scan_cells(int n_scen, ScenDev *scen, long long cap, long long *total) {
  __shared__ long long s_part[1024];
  const int t = threadIdx.x, per = (n_scen + 1023) / 1024, beg = t * per, end = min(beg + per, n_scen);
  long long acc = 0;
  for (int i = beg; i < end; ++i) acc += (scen[i].n_obs + 1) & ~1;
  s_part[t] = acc;
  __syncthreads();
  for (int o = 1; o < 1024; o <<= 1) { const long long v = (t >= o) ? s_part[t - o] : 0; __syncthreads(); s_part[t] += v; __syncthreads(); }
  long long off = s_part[t] - acc;
  for (int i = beg; i < end; ++i) { scen[i].cell_off = off; scen[i].cell_cap = scen[i].n_obs; off += (scen[i].n_obs + 1) & ~1; }
  if (t == 1023) { total[0] = s_part[1023]; total[1] = (s_part[1023] > cap) ? 1 : 0; }
}

__global__ void k_fill_cells(int n_scen, const ScenDev *scen, const uint8_t *cost, const int32_t *col_start, double2 *cells, const long long *total) {
  if (total[1]) return;            // the cell list is too small: nothing is written, the host grows it and launches again
  const int s = blockIdx.x;
  if (s >= n_scen) return;
  const ScenDev &S = scen[s];
  const uint8_t *cm = cost + S.cost_off;
  const int32_t *cs = col_start + S.col_off;
  double2 *out = cells + S.cell_off;
  for (int ix = threadIdx.x; ix < S.nx; ix += blockDim.x) {
    int w = cs[ix];
    const uint8_t *row = cm + (size_t)ix * S.ny;
    const double x = lin_at(S.b[0], S.b[1], S.stepx, S.nx, ix);
    for (int iy = 0; iy < S.ny; ++iy)
      if (row[iy] == 255) out[w++] = make_double2(x, lin_at(S.b[2], S.b[3], S.stepy, S.ny, iy));
  }
}

// ------------------------------------------------------------------------------------------
// API kernels (single-step drop-in methods and kernel-level parity tests)

// one warp per pose
__global__ void __launch_bounds__(128) k_check_batch(avp_config cfg, const ScenDev *scen, int s, const double2 *cells, const int32_t *col_start,
                              int m, const double *poses, uint8_t *out) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (w >= m) return;
  const ScenDev &S = scen[s];
  const bool hit = check_pose_warp(cfg, S, cells + S.cell_off, col_start + S.col_off, poses[3 * w], poses[3 * w + 1], poses[3 * w + 2]);
  if ((threadIdx.x & 31) == 0) out[w] = hit ? 1 : 0;
}

// the same check for every loaded scenario's own start and goal pose (headings wrapped by pi_2_pi as the search does,
// hybrid_a_star.py:105,109): one warp per (scenario, start | goal); out[2*s + which]
__global__ void __launch_bounds__(128) k_check_start_goal(avp_config cfg, const ScenDev *scen, int n, const double2 *cells, const int32_t *col_start, uint8_t *out) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (w >= 2 * n) return;
  const ScenDev &S = scen[w >> 1];
  const double *p = S.pose + 3 * (w & 1);
  const bool hit = check_pose_warp(cfg, S, cells + S.cell_off, col_start + S.col_off, p[0], p[1], pi_2_pi(p[2]));
  if ((threadIdx.x & 31) == 0) out[w] = hit ? 1 : 0;
}

// ------------------------------------------------------------------------------------------
// PathPlanner.split_path (path_planner.py:112-192), SURVEY 8f row 1: gear-change detection and collision-checked
// extension points.  One warp per path.  Gear change at i  <=>  1 - scipy.spatial.distance.cosine(v1, v2) < 0 with
// v1 = p[i+1] - p[i], v2 = p[i+2] - p[i+1]  (:126-136); scipy 1.18 evaluates  dist = 1.0 - uv / sqrt(uu * vv)  with
// np.dot (2-element ddot: fma(a1, b1, a0 * b0), pinned by tests/golden/leaf_split.npz), clips it to [0, 2], and a zero
// displacement gives NaN (no gear change).  The lanes evaluate 32 consecutive i at once; the changes are then handled in
// order: segment k = [the previous change's accepted extension points, last first (:141-147)] + rows start..i+1 +
// [up to extended_num points rolled forward from p[i+1] with the gear of step i, each kept only if collision free
// (:150-174)].  Output: the segments back to back (= out_final_path of path_planning, :52) and their lengths.
// info[4] = status, number of segments, change_gear, number of points.
enum { AVP_SPLIT_OK = 0, AVP_SPLIT_NO_GEAR_CHANGE = 1 /* reference: IndexError at :181 */, AVP_SPLIT_NO_PATH = 2, AVP_SPLIT_CAPACITY = 3 };
#define AVP_EXT_MAX 8

__device__ __forceinline__ bool split_gear_change(const double *p, int i) {
  const double u0 = p[3 * (i + 1)] - p[3 * i], u1 = p[3 * (i + 1) + 1] - p[3 * i + 1];
  const double v0 = p[3 * (i + 2)] - p[3 * (i + 1)], v1 = p[3 * (i + 2) + 1] - p[3 * (i + 1) + 1];
  const double uv = __fma_rn(u1, v1, u0 * v0), uu = __fma_rn(u1, u1, u0 * u0), vv = __fma_rn(v1, v1, v0 * v0);
  double dist = 1.0 - uv / sqrt(uu * vv);
  if (dist < 0.0) dist = 0.0; else if (dist > 2.0) dist = 2.0;      // np.clip (NaN stays NaN)
  return (1 - dist) < 0;
}

__device__ __forceinline__ void split_path_warp(const avp_config &cfg, const ScenDev &S, const double2 *cells, const int32_t *col_start,
                                                const double *p, int nf, double *out, int cap_pts, int32_t *seg_len, int cap_seg, int32_t *info) {
  const int lane = threadIdx.x & 31;
  int start = 0, change = 0, have = 0, nseg = 0, npts = 0; bool over = false;
  double ext[AVP_EXT_MAX][3];                       // accepted extension points of the latest gear change (same values on every lane)
  const int next = cfg.extended_num < AVP_EXT_MAX ? cfg.extended_num : AVP_EXT_MAX;
  auto put = [&](int at, double x, double y, double t) { if (at < cap_pts) { out[3 * at] = x; out[3 * at + 1] = y; out[3 * at + 2] = t; } else over = true; };
  for (int base = 0; base < nf - 2; base += 32) {
    const int ii = base + lane;
    unsigned m = __ballot_sync(AVP_FULL_MASK, ii < nf - 2 && split_gear_change(p, ii));
    while (m) {
      const int i = base + __ffs(m) - 1; m &= m - 1;
      change++;
      const int end = i + 2;
      int pre = 0;
      if (change > 1 && have > 0) {                 // :141-147: insert(0, ...) one by one reverses their order
        for (int j = 0; j < have; ++j) if (lane == 0) put(npts + j, ext[have - 1 - j][0], ext[have - 1 - j][1], ext[have - 1 - j][2]);
        pre = have; have = 0;
      }
      for (int r = lane; r < end - start; r += 32) put(npts + pre + r, p[3 * (start + r)], p[3 * (start + r) + 1], p[3 * (start + r) + 2]);
      int len = pre + (end - start);
      for (int j = 0; j < next; ++j) {              // :150-174
        const double xi = p[3 * i], thi = p[3 * i + 2], xn = p[3 * (i + 1)], yn = p[3 * (i + 1) + 1], thn = p[3 * (i + 1) + 2];
        const bool f1 = (xn > xi) && (thi > -AVP_PI / 2 && thi < AVP_PI / 2);
        const bool f2 = (xn < xi) && ((thi > AVP_PI / 2 && thi < AVP_PI) || (thi > -AVP_PI && thi < -AVP_PI / 2));
        const double speed = (f1 || f2) ? cfg.max_v : -cfg.max_v;
        const double td = speed * cfg.ddt * (j + 1);
        const double cs = d_cos(thn), sn = d_sin(thn);
        const double xj = xn + td * cs, yj = yn + td * sn;
        if (!check_pose_cs_warp(cfg, S, cells, col_start, xj, yj, cs, sn)) {
          if (lane == 0) put(npts + len, xj, yj, thn);
          ext[have][0] = xj; ext[have][1] = yj; ext[have][2] = thn; ++have; ++len;
        }
      }
      if (lane == 0) { if (nseg < cap_seg) seg_len[nseg] = len; else over = true; }
      ++nseg; npts += len; start = i + 1;
    }
  }
  int status = AVP_SPLIT_OK;
  if (nseg == 0) status = AVP_SPLIT_NO_GEAR_CHANGE;            // :181 pre_path = split_path[-1] on an empty list
  else {
    int pre = 0;
    if (have > 0) { for (int j = 0; j < have; ++j) if (lane == 0) put(npts + j, ext[have - 1 - j][0], ext[have - 1 - j][1], ext[have - 1 - j][2]); pre = have; }
    for (int r = lane; r < nf - start; r += 32) put(npts + pre + r, p[3 * (start + r)], p[3 * (start + r) + 1], p[3 * (start + r) + 2]);
    const int len = pre + (nf - start);
    if (lane == 0) { if (nseg < cap_seg) seg_len[nseg] = len; else over = true; }
    ++nseg; npts += len;
  }
  if (__any_sync(AVP_FULL_MASK, over)) status = AVP_SPLIT_CAPACITY;
  if (lane == 0) { info[0] = status; info[1] = nseg; info[2] = change; info[3] = npts; }
}

// every finished plan of the batch (results resident on the device), one warp per scenario
__global__ void __launch_bounds__(128) k_split_batch(avp_config cfg, const ScenDev *scen, int n, const double2 *cells, const int32_t *col_start,
                                                     const avp_plan_summary *sums, const double *paths, int cap_path,
                                                     double *out, int cap_pts, int32_t *seg_len, int cap_seg, int32_t *info) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (w >= n) return;
  const int lane = threadIdx.x & 31;
  int32_t *inf = info + 4 * (size_t)w;
  if (sums[w].status != AVP_OK) { if (lane == 0) { inf[0] = AVP_SPLIT_NO_PATH; inf[1] = 0; inf[2] = 0; inf[3] = 0; } return; }
  if (sums[w].n_final > cap_path) { if (lane == 0) { inf[0] = AVP_SPLIT_CAPACITY; inf[1] = 0; inf[2] = 0; inf[3] = 0; } return; }
  const ScenDev &S = scen[w];
  split_path_warp(cfg, S, cells + S.cell_off, col_start + S.col_off, paths + (size_t)w * cap_path * 3, sums[w].n_final,
                  out + (size_t)w * cap_pts * 3, cap_pts, seg_len + (size_t)w * cap_seg, cap_seg, inf);
}
// one caller-supplied path against scenario s's raster (the drop-in PathPlanner.split_path)
__global__ void __launch_bounds__(32) k_split_one(avp_config cfg, const ScenDev *scen, int s, const double2 *cells, const int32_t *col_start,
                                                  const double *path, int nf, double *out, int cap_pts, int32_t *seg_len, int cap_seg, int32_t *info) {
  const ScenDev &S = scen[s];
  split_path_warp(cfg, S, cells + S.cell_off, col_start + S.col_off, path, nf, out, cap_pts, seg_len, cap_seg, info);
}

// Corridor extraction: path_opti.compute_collision_H (optimization/path_optimazition.py:221-658)
// and its copy ocp_optimization.compute_collision_H (optimization/ocp_optimization.py:36-480).
// One warp per path point; the lanes scan the obstacle cells of the raster columns covered by the AABB of the
// inflated vehicle rectangle + expand_dis (the cell list is sorted by column, np.where order), classify each
// cell into the first of the four areas (right, front, left, rear) whose heading-dependent box contains it
// (:375-645) and keep per lane the minima of the horizontal / vertical distances to that edge; the four minima
// are order independent, so the warp reduction reproduces the reference's sequential loop exactly.
// out4[w] = {x_max, y_max, x_min, y_min} (each <= expand_dis), status[w] = 1 for a heading outside [-pi, pi]
// (the reference leaves `case` unbound there).
__global__ void __launch_bounds__(128) k_corridor(avp_config cfg, const ScenDev *scen, int s, const double2 *cells_all, const int32_t *col_all,
                                                  double expand_dis, int m, const double *poses, double *out4, int32_t *status) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= m) return;
  const ScenDev &S = scen[s];
  const double2 *cells = cells_all + S.cell_off;
  const int32_t *col_start = col_all + S.col_off;
  const double x = poses[3 * w], y = poses[3 * w + 1], th = poses[3 * w + 2], e = expand_dis;
  int shift;
  if (th >= -AVP_PI && th < -AVP_PI / 2) shift = 2;          // case 3
  else if (th >= -AVP_PI / 2 && th < 0) shift = 3;           // case 4
  else if (th >= 0 && th < AVP_PI / 2) shift = 0;            // case 1
  else if (th >= AVP_PI / 2 && th <= AVP_PI) shift = 1;      // case 2
  else { if (lane < 4) out4[4 * w + lane] = NAN; if (lane == 0) status[w] = 1; return; }
  const double sn = d_sin(th), cs = d_cos(th);
  VehGeom g;
  veh_geom(cfg, x, y, cs, sn, g);                            // corners, k, b, sqrt(1 + k*k) of the four edges
  const double bx_max = g.x_max + e, bx_min = g.x_min - e, by_max = g.y_max + e, by_min = g.y_min - e;   // :255-258
  double a0[4], a1[4], a2[4], a3[4];                         // get_area_boundary (:289-294) with the box of (case, area) applied
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const double p1x = g.vb[k][0], p1y = g.vb[k][1], p2x = g.vb[k + 1][0], p2y = g.vb[k + 1][1];     // vb[4] == vb[0]
    const int q = (k + shift) & 3;
    const int sx = (q < 2) ? 1 : -1, sy = (q == 1 || q == 2) ? 1 : -1;
    const double xl = (p2x < p1x) ? p2x : p1x, xh = (p2x > p1x) ? p2x : p1x, yl = (p2y < p1y) ? p2y : p1y, yh = (p2y > p1y) ? p2y : p1y;
    a0[k] = (sx < 0) ? xl - e : xl; a1[k] = (sx > 0) ? xh + e : xh;
    a2[k] = (sy < 0) ? yl - e : yl; a3[k] = (sy > 0) ? yh + e : yh;
  }
  const double as = fabs(sn), ac = fabs(cs);
  double x_max = e, x_min = e, y_max = e, y_min = e;
  int lo, hi;
  col_range(S, bx_min, bx_max, lo, hi);
  if (lo <= hi) {
    const int beg = col_start[lo], end = col_start[hi + 1];
    for (int i = beg + lane; i < end; i += 32) {
      const double2 p = cells[i];
      if (!(p.x >= bx_min && p.x <= bx_max && p.y >= by_min && p.y <= by_max)) continue;          // :266-275
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (p.x > a0[k] && p.x < a1[k] && p.y > a2[k] && p.y < a3[k]) {
          const double sd = fabs(g.lk[k] * p.x + g.lb[k] - p.y) / g.ls[k];                         // compute_distance (:296-298)
          const double ver = sd / ac, hor = sd / as;                                                   // :303-305
          const int q = (k + shift) & 3;
          if (q < 2) { if (hor < x_max) x_max = hor; } else { if (hor < x_min) x_min = hor; }
          if (q == 1 || q == 2) { if (ver < y_max) y_max = ver; } else { if (ver < y_min) y_min = ver; }
          break;
        }
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    double v;
    v = shfl_dbl(x_max, lane ^ o); if (v < x_max) x_max = v;
    v = shfl_dbl(y_max, lane ^ o); if (v < y_max) y_max = v;
    v = shfl_dbl(x_min, lane ^ o); if (v < x_min) x_min = v;
    v = shfl_dbl(y_min, lane ^ o); if (v < y_min) y_min = v;
  }
  if (lane == 0) { out4[4 * w] = x_max; out4[4 * w + 1] = y_max; out4[4 * w + 2] = x_min; out4[4 * w + 3] = y_min; status[w] = 0; }
}

// rs length (normalised-by-maxc L/maxc) of pose -> goal, warp-collective; cand: 46 entries of shared memory
__device__ __forceinline__ void rs_length_warp(const double q0[3], const double q1[3], double maxc, int xy_np, int phi_np,
                                               RsCand *cand, RsBest &best) {
  const int lane = threadIdx.x & 31;
  RsQuery Q; rs_query(q0, q1, maxc, Q);
  unsigned long long valid = 0ull;
  for (int base = 0; base < RS_NINST; base += 32) {
    const int inst = base + lane;
    bool ok = false;
    if (inst < RS_NINST) { double t, u, v; ok = rs_eval_instance(inst, Q, t, u, v); if (ok) { cand[inst].t = t; cand[inst].u = u; cand[inst].v = v; cand[inst].L = rs_cand_L(inst, cand[inst], xy_np, phi_np); } }
    valid |= (unsigned long long)__ballot_sync(AVP_FULL_MASK, ok) << base;
  }
  __syncwarp();
  rs_select(cand, valid, xy_np, phi_np, maxc, best);   // every lane computes the same result
}

// one CTA, one warp per successor (hybrid_a_star.py:133-151,185-204 + rs length)
__global__ void __launch_bounds__(AVP_NCHILD_MAX * 32) k_expand_pure(avp_config cfg, const ScenDev *scen, int s, const double2 *cells, const int32_t *col_start,
                              double px, double py, double pth, double *out_pose, int32_t *out_flags, double *out_rsL) {
  __shared__ RsCand cand[AVP_NCHILD_MAX][RS_NINST];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ns = cfg.steering_angle_num;
  if (w >= 2 * ns) return;
  const ScenDev &S = scen[s];
  const double tn = cfg.tan_steer[w % ns];
  const bool fwd = w < ns;
  const double speed = fwd ? cfg.max_v : -cfg.max_v;
  const double th = pi_2_pi(pth + (cfg.max_v * tn) / cfg.lw * cfg.dt);
  const double td = speed * cfg.dt;
  const double x_ = px + td * d_cos(th), y_ = py + td * d_sin(th);
  int fl = 0;
  for (int k = 0; k < cfg.n_substeps; ++k) {
    const double th_i = pi_2_pi(pth + (cfg.max_v * tn) / cfg.lw * cfg.ddt * (k + 1));
    const double td_i = speed * cfg.ddt * (k + 1);
    if (check_pose_warp(cfg, S, cells + S.cell_off, col_start + S.col_off, px + td_i * d_cos(th_i), py + td_i * d_sin(th_i), th_i)) { fl |= 1; break; }
  }
  if (x_ > S.b[1] || x_ < S.b[0] || y_ > S.b[3] || y_ < S.b[2]) fl |= 2;
  const double q0[3] = {x_, y_, th}, q1[3] = {S.pose[3], S.pose[4], pi_2_pi(S.pose[5])};
  const double maxc = 1 / cfg.min_radius_turn;
  RsBest best; rs_length_warp(q0, q1, maxc, 1, 1, cand[w], best);
  if (lane == 0) {
    out_pose[3 * w] = x_; out_pose[3 * w + 1] = y_; out_pose[3 * w + 2] = th;
    out_flags[w] = fl;
    out_rsL[w] = best.ok ? best.L / maxc : NAN;
  }
}

// one warp per query (rs_curve.py:99-134), selected word + its course
__global__ void __launch_bounds__(128) k_rs_optimal(int m, const double *q, double maxc, double step_size, int xy_np, int phi_np,
                             double *lengths, char *ctypes, int32_t *nseg, double *L, int cap_pts,
                             double *x, double *y, double *yaw, int32_t *dir, int32_t *n_pts) {
  __shared__ RsCand cand[4][RS_NINST];
  const int wl = threadIdx.x >> 5, w = blockIdx.x * (blockDim.x >> 5) + wl, lane = threadIdx.x & 31;
  if (w >= m) return;
  const double q0[3] = {q[6 * w], q[6 * w + 1], q[6 * w + 2]}, q1[3] = {q[6 * w + 3], q[6 * w + 4], q[6 * w + 5]};
  RsBest best; rs_length_warp(q0, q1, maxc, xy_np, phi_np, cand[wl], best);
  if (lane == 0) {
    if (!best.ok) { nseg[w] = best.degenerate ? -2 : -1; n_pts[w] = 0; return; }
    const int np_ = rs_course(best, maxc, step_size, q0, cap_pts, x + (size_t)w * cap_pts, y + (size_t)w * cap_pts,
                              yaw + (size_t)w * cap_pts, dir + (size_t)w * cap_pts);
    n_pts[w] = np_;
    nseg[w] = best.n;
    for (int i = 0; i < 5; ++i) lengths[5 * w + i] = (i < best.n) ? best.len[i] / maxc : 0.0;
    for (int i = 0; i < 8; ++i) ctypes[8 * w + i] = rs_ct_names[best.ct][i];
    L[w] = best.L / maxc;
  }
}

// ------------------------------------------------------------------------------------------
// Dijkstra (path_plan/compute_h.py), verbatim CPython heapq emulation, run by ONE warp.
// Heap entries are 64-bit keys (distance << 32 | grid_id): Grid.__lt__ (compute_h.py:33-38)
// is an unsigned 64-bit compare.  The heap is the array gheap[0 .. gcap) in global memory; while a kernel works
// on it the first AVP_SM_HEAP entries live in shared memory (sheap) and gheap[0 .. AVP_SM_HEAP) is their save area
// (dij_heap_load / dij_heap_store), so a search can be suspended on one SM and resumed on another.

// AVP_PREFETCH_MODE (A/B builds): 0 none, 1 prefetch.global.L1, 2 a load whose result is never used (default: the only mode that
// measured faster than none, profiles/experiments_r02.md), 3 prefetch.global.L2
#ifndef AVP_PREFETCH_MODE
#define AVP_PREFETCH_MODE 2
#endif
__device__ __forceinline__ void prefetch_l1(const void *p) {
#if AVP_PREFETCH_MODE == 1
  asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
#elif AVP_PREFETCH_MODE == 2
  unsigned sink; asm volatile("ld.global.ca.u32 %0, [%1];" : "=r"(sink) : "l"(p));
#elif AVP_PREFETCH_MODE == 3
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#endif
}

struct DijCtx {
  const ScenDev *S; const uint8_t *cost;
  int32_t *hval, *ost; double *gx, *gy;
  unsigned long long *sheap, *gheap; int gcap;
  int hn; int closed_len; int status;
};

__device__ __forceinline__ double shfl_d(double v, int src) {
  int lo = __double2loint(v), hi = __double2hiint(v);
  lo = __shfl_sync(AVP_FULL_MASK, lo, src); hi = __shfl_sync(AVP_FULL_MASK, hi, src);
  return __hiloint2double(hi, lo);
}
__device__ __forceinline__ unsigned long long shfl_u64(unsigned long long v, int src) {
  unsigned lo = (unsigned)v, hi = (unsigned)(v >> 32);
  lo = __shfl_sync(AVP_FULL_MASK, lo, src); hi = __shfl_sync(AVP_FULL_MASK, hi, src);
  return ((unsigned long long)hi << 32) | lo;
}

// floor(a / d) for d > 0, exactly as IEEE division followed by floor: the quotient estimated with the
// reciprocal decides unless it lies within 1e-9 of an integer, where the true division is performed.
__device__ __forceinline__ long long floor_div_exact(double a, double d, double inv_d) {
  const double q = a * inv_d, f = floor(q), r = q - f;
  if (r > 1e-9 && r < 1.0 - 1e-9 && fabs(q) < 1e5) return (long long)f;   // |q - a/d| <= 4 ulp(q) < 1e-10 here
  return (long long)floor(a / d);
}

// Dijkstra.compute_path (compute_h.py:198-214); warp-collective.  Returns the popped distance of the
// target cell, or -1 if the queue ran dry (reference: blocks forever) / capacity.
// `sheap` must be the kernel's __shared__ heap array (the function is inlined so that the
// compiler keeps the shared address space); the queue persists in D / sheap / D.gheap across calls.
__device__ __forceinline__ int dij_compute_path(DijCtx &D, unsigned long long *sheap, double node_x, double node_y, long long *term_out) {
  const ScenDev &S = *D.S;
  const int lane = threadIdx.x & 31;
  // loop invariants in registers
  const uint8_t *cost = D.cost; int32_t *hval = D.hval, *ost = D.ost; double *gxa = D.gx, *gya = D.gy;
  unsigned long long *gheap = D.gheap; const int gcap = D.gcap;
  const double b0 = S.b[0], b1 = S.b[1], b2 = S.b[2], b3 = S.b[3], dx = S.dx, dy = S.dy;
  const int nx = S.nx, ny = S.ny, mx = S.mx, my = S.my, stride = S.stride, n_ids = S.n_ids;
  int hn = D.hn, closed_len = D.closed_len, status = 0;
  const double inv_dx = 1.0 / dx, inv_dy = 1.0 / dy;
  // heappop walks five levels per round (below): lane -> sibling pair q of level d under the current position (lane 31: none), the
  // lanes holding the pair's ancestors inside the round (pl_anc) and those of them that have to choose their RIGHT child for the pair
  // to lie on the path (pl_want) -- constants of the lane
  const int pl_d = 32 - __clz(lane + 1), pl_q = lane + 1 - (1 << (pl_d - 1));
  unsigned pl_anc = 0u, pl_want = 0u;
  for (int k = 1; k < pl_d && pl_d <= 5; ++k) { const int a = (1 << (k - 1)) - 1 + (pl_q >> (pl_d - k)); pl_anc |= 1u << a; if ((pl_q >> (pl_d - k - 1)) & 1) pl_want |= 1u << a; }

#define HP_GET(i) ((i) < AVP_SM_HEAP ? sheap[(i)] : gheap[(i)])
#define HP_SET(i, v) do { if ((i) < AVP_SM_HEAP) sheap[(i)] = (v); else gheap[(i)] = (v); } while (0)

  const long long term = map_index(S, node_x, node_y);
  if (term_out) *term_out = term;
  double cur_x = S.pose[3], cur_y = S.pose[4];
  int cur_dist = 0;
  const long long gid = map_index(S, cur_x, cur_y);
  if (lane == 0) { if (gid >= 0 && gid < n_ids && hval[gid] < 0) hval[gid] = 0; }   // initial_map (:50-72)
  closed_len++;
  __syncwarp();
  int result = -1;
  for (long long guard = 0;; ++guard) {
    if (guard > 4ll * n_ids + 1024) { status = AVP_CAPACITY; break; }   // cannot happen: every id is pushed at most twice
    // update_openlist (compute_h.py:84-195): the 8 neighbours evaluated by lanes 0..7
    int action = 0, prio = 0; int nid = 0; unsigned long long key = 0ull;
    if (lane < 8) {
      const int ddx = (lane == 0 || lane == 3 || lane == 5) ? -1 : ((lane == 2 || lane == 4 || lane == 7) ? 1 : 0);
      const int ddy = (lane < 3) ? 1 : ((lane < 5) ? 0 : -1);
      const double ngx = ddx < 0 ? cur_x - dx : (ddx > 0 ? cur_x + dx : cur_x);
      const double ngy = ddy < 0 ? cur_y - dy : (ddy > 0 ? cur_y + dy : cur_y);
      const long long qx = floor_div_exact(ngx - b0, dx, inv_dx);             // shared by is_obstacle and convert_position_to_index
      // is_obstacle (compute_h.py:237-255)
      long long xi = qx - 1, yi = floor_div_exact(ngy - b2, dy, inv_dy) - 1;
      if (xi >= mx) xi = mx - 1; if (yi >= my) yi = my - 1;
      if (xi < 0) xi += nx; if (yi < 0) yi += ny;             // python negative indexing
      const bool in_map = xi >= 0 && xi < nx && yi >= 0 && yi < ny;
      const long long id = qx + floor_div_exact(b3 - ngy, dy, inv_dy) * (long long)stride;   // costmap.py:319-329
      const bool id_ok = id >= 0 && id < n_ids;
      // both loads are issued before either result is needed (one memory round trip instead of two)
      const uint8_t cv = in_map ? cost[(size_t)xi * ny + yi] : (uint8_t)0;
      const int st = id_ok ? ost[(int)id] : -3;
      if (cv != 255) {
        bool ok = true;
        if (ddx < 0 && !(ngx >= b0)) ok = false; if (ddx > 0 && !(ngx <= b1)) ok = false;
        if (ddy > 0 && !(ngy <= b3)) ok = false; if (ddy < 0 && !(ngy >= b2)) ok = false;
        if (ok && id_ok) {
          nid = (int)id;
          prio = cur_dist + ((ddx && ddy) ? 14 : 10);
          key = ((unsigned long long)(unsigned)prio << 32) | (unsigned)nid;
          if (st == -1) {                    // first visit (add_grid_to_openlist :229-235): this lane records the Grid
            action = 1; ost[nid] = prio; gxa[nid] = ngx; gya[nid] = ngy;
          } else if (st >= 0 && st > prio) action = 2;      // queued with a larger distance (:222-228)
        }
      }
    }
    unsigned pm = __ballot_sync(AVP_FULL_MASK, action == 1) & 0xffu, dm = __ballot_sync(AVP_FULL_MASK, action == 2) & 0xffu;
    unsigned todo = pm | dm;
    while (todo) {                            // in the reference's neighbour order
      const int k = __ffs(todo) - 1; todo &= todo - 1;
      const unsigned long long key_k = shfl_u64(key, k);
      if ((pm >> k) & 1u) {                   // heappush (hn and status are tracked by every lane: no broadcast after the loop)
        if (hn >= gcap) status = AVP_CAPACITY;
        else {
          if (lane == 0) {
            int pos = hn;
            while (pos > 0) {                 // heapq._siftdown
              const int parent = (pos - 1) >> 1;
              const unsigned long long pv = HP_GET(parent);
              if (key_k < pv) { HP_SET(pos, pv); pos = parent; continue; }
              break;
            }
            HP_SET(pos, key_k);
          }
          hn++;
        }
      } else {                                // overwrite IN PLACE, no re-sift (:222-228)
        const int hn_b = hn;
        __syncwarp();
        int pos = 0x7fffffff;
        for (int i = lane; i < hn_b; i += 32) if ((unsigned)HP_GET(i) == (unsigned)key_k) { pos = i; break; }
        for (int o = 16; o > 0; o >>= 1) pos = min(pos, __shfl_xor_sync(AVP_FULL_MASK, pos, o));
        if (lane == 0 && pos != 0x7fffffff) { HP_SET(pos, key_k); ost[(int)(unsigned)key_k] = (int)(key_k >> 32); }
        __syncwarp();
      }
    }
    __syncwarp();                               // lane 0's heap writes are visible to the warp
    if (status) break;
    if (hn == 0) { status = AVP_H_UNREACHABLE; break; }
    // update_closedlist (:74-82): heappop
    const unsigned long long top = sheap[0];      // every lane reads the root (a shared-memory broadcast)
    const int cur_id = (int)(unsigned)top;
    const double nxt_x = gxa[cur_id], nxt_y = gya[cur_id];     // issued before the sift so that the latency overlaps it
    // the popped cell is expanded next: its neighbours' state words and cost-map bytes are prefetched while lane 0 sifts (the
    // positions are estimated from the id -- ids and raster indices are floors of the same lattice coordinates; a wrong guess
    // only costs the prefetch)
    if (lane >= 1 && lane <= 6) {      // (prefetching one pop further ahead -- the root's two children -- measured 1.5 % slower: not done)
      const int i1 = cur_id / stride, i0 = cur_id - i1 * stride;
      if (lane <= 3) { const int q = (i1 + lane - 2) * stride + i0 - 1; if (q >= 0 && q + 2 < n_ids) { prefetch_l1(&ost[q]); prefetch_l1(&ost[q + 2]); } }
      else { const int xi = i0 - 1 + (lane - 5), yi = my - i1 - 3; if (xi >= 0 && xi < nx && yi >= 0 && yi + 2 < ny) { prefetch_l1(&cost[(size_t)xi * ny + yi]); prefetch_l1(&cost[(size_t)xi * ny + yi + 2]); } }
    }
#ifdef AVP_DIJ_SERIAL_POP       // A/B build: the serial sift of round 1
    if (lane == 0) {
      const unsigned long long item = HP_GET(hn - 1);
      const int n = hn - 1;
      if (n > 0) {                            // heapq._siftup
        int pos = 0, child = 1;
        while (child < n) {
          const int right = child + 1;
          unsigned long long cv = HP_GET(child);
          if (right < n) { const unsigned long long rv = HP_GET(right); if (!(cv < rv)) { child = right; cv = rv; } }
          HP_SET(pos, cv); pos = child; child = 2 * pos + 1;
        }
        while (pos > 0) {                     // heapq._siftdown(heap, 0, pos)
          const int parent = (pos - 1) >> 1;
          const unsigned long long pv = HP_GET(parent);
          if (item < pv) { HP_SET(pos, pv); pos = parent; continue; }
          break;
        }
        HP_SET(pos, item);
      }
    }
#else
    // heapq.heappop: the last entry replaces the root and _siftup moves the smaller child up, level by level, to a leaf.  Which child
    // is smaller depends on the heap only, not on the item, so FIVE levels are walked at once: the 31 sibling pairs below the
    // current position are compared by 31 lanes (one load round), a ballot holds every outcome, and the path is read off it with
    // bit operations; the lanes on the path move their winner up.  Same compares, same moves as the serial loop.
    {
      const int n = hn - 1;
      unsigned long long item = 0ull;
      if (lane == 0) item = HP_GET(n);
      if (n > 0) {
        int pos = 0;
        for (;;) {
          const int L = ((pos + 1) << pl_d) - 1 + 2 * pl_q;                              // left entry of this lane's pair
          const bool have = lane < 31 && L < n, have_r = have && L + 1 < n;
          unsigned long long lv = 0ull, rv = 0ull;
          if (have) lv = HP_GET(L);
          if (have_r) rv = HP_GET(L + 1);
          const bool right = have_r && !(lv < rv);
          const unsigned long long cv = right ? rv : lv;
          const unsigned rm = __ballot_sync(AVP_FULL_MASK, right);
          // the pair lies on the path iff it exists and every ancestor pair chose the child it hangs below: one mask compare per lane
          // instead of walking the five levels one after the other (an existing pair's ancestors exist)
          const bool on = have && ((rm & pl_anc) == pl_want);
          const unsigned om = __ballot_sync(AVP_FULL_MASK, on);
          if (on) HP_SET(((pos + 1) << (pl_d - 1)) - 1 + pl_q, cv);                     // the winner moves up to the pair's parent
          __syncwarp();
          if (om == 0u) break;                                                           // no child: a leaf
          const int t = 31 - __clz(om);                                                  // the deepest pair on the path (lane ids grow with the level)
          const int dt = 32 - __clz(t + 1), qt = t + 1 - (1 << (dt - 1));
          pos = ((pos + 1) << dt) - 1 + 2 * qt + (int)((rm >> t) & 1u);
          if (dt < 5) break;                                                             // the path ended inside the round: a leaf
        }
        if (lane == 0) {
          while (pos > 0) {                     // heapq._siftdown(heap, 0, pos)
            const int parent = (pos - 1) >> 1;
            const unsigned long long pv = HP_GET(parent);
            if (item < pv) { HP_SET(pos, pv); pos = parent; continue; }
            break;
          }
          HP_SET(pos, item);
        }
      }
    }
#endif
    if (lane == 0) {
      const int id = (int)(unsigned)top;
      ost[id] = -2;
      // closedlist: the first entry per id wins (hybrid_a_star.py:272-283).  Only the goal cell is ever closed twice
      // (initial_map puts it there with distance 0 and a neighbour re-pushes it once), so every other
      // id is written unconditionally: no load of hval on the pop path
      if ((long long)id != gid) hval[id] = (int)(top >> 32);
    }
    hn -= 1; closed_len++;
    cur_dist = (int)(top >> 32);
    __syncwarp();
    if ((long long)cur_id == term) { result = cur_dist; break; }
    cur_x = nxt_x; cur_y = nxt_y;
  }
  if (lane == 0) { D.hn = hn; D.closed_len = closed_len; if (status) D.status = status; }
  __syncwarp();
  return result;
#undef HP_GET
#undef HP_SET
}

// Stand-alone resumable Dijkstra query (drop-in for compute_h.Dijkstra.compute_path, single-step API):
// the queue of scenario `s` persists between launches in `save` (shared-memory part), gheap and st[].
struct DijPersist { int hn, closed_len, status, inited; };

__global__ void __launch_bounds__(32) k_dij_query(const ScenDev *scen, int s, const uint8_t *cost, int32_t *hval, int32_t *ost, double *gx, double *gy,
                                                  unsigned long long *gheap, int gcap, DijPersist *st,
                                                  int reset, double x, double y, int32_t *out3) {
  __shared__ unsigned long long s_heap[AVP_SM_HEAP];
  __shared__ DijCtx D;
  const ScenDev &S = scen[s];
  const int lane = threadIdx.x;
  if (reset || !st->inited) {
    for (int i = lane; i < S.n_ids; i += 32) { hval[S.id_off + i] = -1; ost[S.id_off + i] = -1; }
    if (lane == 0) { st->hn = 0; st->closed_len = 0; st->status = 0; st->inited = 1; }
    __syncwarp();
  }
  const int hn0 = st->hn;
  for (int i = lane; i < hn0 && i < AVP_SM_HEAP; i += 32) s_heap[i] = gheap[i];
  if (lane == 0) {
    D.S = &S; D.cost = cost + S.cost_off; D.hval = hval + S.id_off; D.ost = ost + S.id_off; D.gx = gx + S.id_off; D.gy = gy + S.id_off;
    D.sheap = s_heap; D.gheap = gheap; D.gcap = gcap; D.hn = hn0; D.closed_len = st->closed_len; D.status = 0;
  }
  __syncwarp();
  long long term;
  const int d = dij_compute_path(D, s_heap, x, y, &term);
  __syncwarp();
  const int hn1 = D.hn;
  for (int i = lane; i < hn1 && i < AVP_SM_HEAP; i += 32) gheap[i] = s_heap[i];
  if (lane == 0) { st->hn = hn1; st->closed_len = D.closed_len; st->status = D.status; out3[0] = d; out3[1] = D.closed_len; out3[2] = (int)term; out3[3] = D.status; }
}

// ------------------------------------------------------------------------------------------
// hybrid A* containers (per slot)

#define AVP_PENDING 7   // internal: the search let go of its SM at the end of a quantum (k_plan) and waits in the run queue

__device__ __forceinline__ unsigned long long pose_hash(double x, double y, double t) {
  const unsigned long long a = (unsigned long long)__double_as_longlong(x + 0.0), b = (unsigned long long)__double_as_longlong(y + 0.0),
                           c = (unsigned long long)__double_as_longlong(t + 0.0);
  unsigned long long h = a * 0x9E3779B97F4A7C15ULL; h ^= (h >> 29); h += b * 0xBF58476D1CE4E5B9ULL; h ^= (h >> 31);
  h += c * 0x94D049BB133111EBULL; h ^= (h >> 30);
  return (h * 0xD6E8FEB86659FD93ULL) >> 20;
}
// exact-pose lookup: replaces the == scans over closed_list / open_list.queue (hybrid_a_star.py:155-172)
__device__ __forceinline__ int htab_find(const int32_t *htab, int mask, const Node *nodes, double x, double y, double t) {
  unsigned long long p = pose_hash(x, y, t);
  for (int probes = 0; probes <= mask; ++probes, ++p) {
    const int e = htab[p & mask];
    if (e < 0) return -1;
    const Node &n = nodes[e];
    if (n.x == x && n.y == y && n.theta == t) return e;
  }
  return -1;
}
// the same probe issued by a warp other than the inserting one (the pipelined kernel's evaluators): the slot is read at L2
// (ld.global.cg), where the inserts' atomicCAS operations are performed
__device__ __forceinline__ int htab_find_cg(const int32_t *htab, int mask, const Node *nodes, double x, double y, double t) {
  unsigned long long p = pose_hash(x, y, t);
  for (int probes = 0; probes <= mask; ++probes, ++p) {
    const int e = __ldcg(&htab[p & mask]);
    if (e < 0) return -1;
    const Node &n = nodes[e];
    if (n.x == x && n.y == y && n.theta == t) return e;
  }
  return -1;
}
// concurrent insert (distinct poses): claim the first empty slot of the probe sequence
__device__ __forceinline__ void htab_insert(int32_t *htab, int mask, const Node *nodes, int idx) {
  const Node &n = nodes[idx];
  unsigned long long p = pose_hash(n.x, n.y, n.theta);
  for (int probes = 0; probes <= mask; ++probes, ++p)
    if (atomicCAS((int *)&htab[p & mask], -1, idx) == -1) return;
}

// open_list: CPython heapq of Node objects ordered by Node.__lt__ (f only, hybrid_a_star.py:61-68).
// Entries carry a copy of f (kept in sync on the reference's in-place updates through Node.hpos);
// the heap is the array ge[0 .. node_cap) of 16-byte records in global memory (ONE load per entry: a level of a sift costs one
// memory round trip, not two); while a kernel works on it the first SMO entries live in shared memory (separate key / index
// arrays) and ge[0 .. SMO) is their save area, so a search can be suspended on one SM and resumed on another.
// The functions are force-inlined and take the __shared__ arrays themselves so that the compiler keeps the
// shared address space.
struct __align__(16) OEnt { double f; int32_t idx; int32_t pad; };
template <int SMO>
__device__ __forceinline__ void oh_get(const double *sf, const int32_t *si, const OEnt *ge, int i, double &f, int &idx) {
  if (i < SMO) { f = sf[i]; idx = si[i]; }
  else { const int4 v = *reinterpret_cast<const int4 *>(&ge[i]); f = __hiloint2double(v.y, v.x); idx = v.z; }
}
#define OH_SET(i, f_, idx_) do { if ((i) < SMO) { sf[(i)] = (f_); si[(i)] = (idx_); } else { int4 v_; v_.x = __double2loint(f_); v_.y = __double2hiint(f_); v_.z = (idx_); v_.w = 0; *reinterpret_cast<int4 *>(&ge[(i)]) = v_; } nodes[(idx_)].hpos = (i); } while (0)
// The sifts are chains of dependent loads: below the shared-memory head every level is a DRAM / L2 round trip (the heaps of the
// long searches do not fit the L2 together).  The addresses of the next levels are known before their data is needed, so they
// are prefetched (no effect on any result):
//   push  the ancestors of the next k positions [n, n + k) are contiguous per level: one or two lines per level, all of them at once
//   pop   the four levels below a position are four contiguous ranges (32, 64, 128, 256 bytes): one round trip per four levels
template <int SMO>
__device__ __forceinline__ void oh_prefetch_push(const OEnt *ge, int n, int k, int lane) {
  // lane l: the level-l ancestors ((p + 1) >> l) - 1 of p in [n, n + k): at most k / 2^l + 1 adjacent entries
  if (lane >= 1 && lane < 24) {
    const int lo = ((n + 1) >> lane) - 1, hi = ((n + k) >> lane) - 1;
    if (hi >= SMO && lo >= 0) { prefetch_l1(&ge[lo < SMO ? SMO : lo]); prefetch_l1(&ge[hi]); }
  }
}
template <int SMO>
__device__ __forceinline__ void oh_prefetch_subtree(const OEnt *ge, int pos, int last) {
  int lo = 2 * pos + 1, cnt = 2;               // level d below pos: entries [lo, lo + cnt), 16 bytes each
#pragma unroll
  for (int d = 0; d < 4; ++d) {
    if (lo >= last) break;
    int hi = lo + cnt - 1; if (hi >= last) hi = last - 1;
    if (hi >= SMO) {
      prefetch_l1(&ge[lo < SMO ? SMO : lo]); prefetch_l1(&ge[hi]);
      if (cnt == 16 && lo + 8 <= hi && lo + 8 >= SMO) prefetch_l1(&ge[lo + 8]);
    }
    lo = 2 * lo + 1; cnt *= 2;
  }
}
template <int SMO>
__device__ __forceinline__ void oh_siftdown(double *sf, int32_t *si, OEnt *ge, Node *nodes, int pos, double fi, int item) {   // heapq._siftdown(heap, 0, pos)
  while (pos > 0) {
    const int parent = (pos - 1) >> 1; double pf; int pi; oh_get<SMO>(sf, si, ge, parent, pf, pi);
    if (fi < pf) { OH_SET(pos, pf, pi); pos = parent; continue; }
    break;
  }
  OH_SET(pos, fi, item);
}
template <int SMO>
__device__ __forceinline__ void oh_push(double *sf, int32_t *si, OEnt *ge, Node *nodes, int &n, double f, int idx) {
  const int pos = n++;
  oh_siftdown<SMO>(sf, si, ge, nodes, pos, f, idx);
}
// heapq.heappop after the root has been read: move the last entry to the root and _siftup
template <int SMO>
__device__ __forceinline__ void oh_pop_fix(double *sf, int32_t *si, OEnt *ge, Node *nodes, int &n) {
  const int last = n - 1;
  double fi; int item; oh_get<SMO>(sf, si, ge, last, fi, item);
  n = last;
  if (last == 0) return;
  int pos = 0, child = 1, lvl = 0;
  while (child < last) {
    const int right = child + 1;
    if (2 * child + 2 >= SMO && (lvl++ & 3) == 0) oh_prefetch_subtree<SMO>(ge, pos, last);      // the next four levels in one round trip
    double cf; int ci; oh_get<SMO>(sf, si, ge, child, cf, ci);
    if (right < last) { double rf; int ri; oh_get<SMO>(sf, si, ge, right, rf, ri); if (!(cf < rf)) { child = right; cf = rf; ci = ri; } }
    OH_SET(pos, cf, ci); pos = child; child = 2 * pos + 1;
  }
  oh_siftdown<SMO>(sf, si, ge, nodes, pos, fi, item);
}
// (A/B build -DAVP_WARP_POP; measured 2-3 % SLOWER than the serial oh_pop_fix with prefetches on the bench workload: 62 scattered
// 16-byte loads per round instead of 2 -- profiles/experiments_r02.md)
// heapq.heappop after the root has been read, WARP-COLLECTIVE: five levels of _siftup per round (see dij_compute_path): the 31 sibling
// pairs below the position are loaded and compared by 31 lanes at once -- below the shared-memory head that is ONE DRAM round trip
// for five levels instead of five --, the path is read off the ballot, the lanes on the path move their winner up.  The final
// _siftdown (the moved item bubbling up) is lane 0's.
template <int SMO>
__device__ __forceinline__ void oh_pop_fix_warp(double *sf, int32_t *si, OEnt *ge, Node *nodes, int n_before, int lane) {
  const int last = n_before - 1;
  double fi = 0.0; int item = 0;
  if (lane == 0) oh_get<SMO>(sf, si, ge, last, fi, item);
  if (last == 0) return;
  int pos = 0;
  for (;;) {
    const int d = 32 - __clz(lane + 1), q = lane + 1 - (1 << (d - 1));
    const long long L = (((long long)pos + 1) << d) - 1 + 2 * q;
    const bool have = lane < 31 && L < last, have_r = have && L + 1 < last;
    double lf = 0.0, rf = 0.0; int li = 0, ri = 0;
    if (have) oh_get<SMO>(sf, si, ge, (int)L, lf, li);
    if (have_r) oh_get<SMO>(sf, si, ge, (int)L + 1, rf, ri);
    const bool right = have_r && !(lf < rf);
    const double cf = right ? rf : lf; const int ci = right ? ri : li;
    const unsigned vm = __ballot_sync(AVP_FULL_MASK, have), rm = __ballot_sync(AVP_FULL_MASK, right);
    int node = pos, qq = 0, depth = 0, my_parent = -1;
#pragma unroll
    for (int dd = 1; dd <= 5; ++dd) {
      const int lid = (1 << (dd - 1)) - 1 + qq;
      if (depth != dd - 1 || !((vm >> lid) & 1u)) break;
      if (lid == lane) my_parent = node;
      const int r = (rm >> lid) & 1u;
      node = (int)((((long long)pos + 1) << dd) - 1 + 2 * qq + r); qq = 2 * qq + r; depth = dd;
    }
    if (my_parent >= 0) OH_SET(my_parent, cf, ci);
    __syncwarp();
    pos = node;
    if (depth < 5) break;
  }
  if (lane == 0) oh_siftdown<SMO>(sf, si, ge, nodes, pos, fi, item);
  __syncwarp();
}
// heapq._siftdown(heap, 0, pos), WARP-COLLECTIVE: the ancestors of `pos` are known before any of them is read, so lane l reads
// ancestor l + 1 -- every level of the sift in ONE round trip --, the ballot of (item < ancestor) gives the number k of levels the
// item rises (the serial loop stops at the first ancestor that is not larger), lanes 0..k-1 move their ancestor one level down
// and lane k writes the item.  Same reads (each ancestor before it is overwritten), same writes as the serial loop.
// `fi` / `item` / `pos` must be the same on every lane; every lane of the warp calls.
#ifndef AVP_SERIAL_PUSH
template <int SMO>
__device__ __forceinline__ void oh_siftdown_warp(double *sf, int32_t *si, OEnt *ge, Node *nodes, int pos, double fi, int item, int lane) {
  const int anc = (lane < 30) ? (((pos + 1) >> (lane + 1)) - 1) : -1;
  double pf = 0.0; int pi = 0;
  if (anc >= 0) oh_get<SMO>(sf, si, ge, anc, pf, pi);
  const unsigned up = __ballot_sync(AVP_FULL_MASK, anc >= 0 && fi < pf);
  const int k = __ffs(~up) - 1;                 // lanes 30, 31 never vote: ~up != 0
  const int dst = ((pos + 1) >> lane) - 1;      // ancestor `lane` of pos (ancestor 0 = pos itself)
  if (lane < k) { OH_SET(dst, pf, pi); }
  else if (lane == k) { OH_SET(dst, fi, item); }
  __syncwarp();
}
#else
template <int SMO>
__device__ __forceinline__ void oh_siftdown_warp(double *sf, int32_t *si, OEnt *ge, Node *nodes, int pos, double fi, int item, int lane) {
  if (lane == 0) oh_siftdown<SMO>(sf, si, ge, nodes, pos, fi, item);
  __syncwarp();
}
#endif
template <int SMO>
__device__ __forceinline__ void oh_push_warp(double *sf, int32_t *si, OEnt *ge, Node *nodes, int &n, double f, int idx, int lane) {
  const int pos = n++;
  oh_siftdown_warp<SMO>(sf, si, ge, nodes, pos, f, idx, lane);
}
// heapq.heappop after the root has been read, HYBRID: while the five levels below the position lie in the shared-memory head the
// warp takes them in one round (31 sibling pairs compared at once, as oh_pop_fix_warp: two rounds cover the 11 levels of a
// 2048-entry head); below the head lane 0 descends level by level behind the four-level prefetch (there the 62 scattered loads
// of a warp round cost more than they save); the final _siftdown (the moved item rising again -- in an A* open list it is a
// recent push with a small f and rises far) is the warp-collective one.
template <int SMO>
__device__ __forceinline__ void oh_pop_fix_hybrid(double *sf, int32_t *si, OEnt *ge, Node *nodes, int n_before, int lane) {
  const int last = n_before - 1;
  if (last == 0) return;
  double fi = 0.0; int item = 0;
  if (lane == 0) oh_get<SMO>(sf, si, ge, last, fi, item);
  int pos = 0;
  bool bottom = false;
#ifndef AVP_SERIAL_POP_HEAD
  // lane -> sibling pair q of level d below the position; the pair lies on the path iff it exists and every ancestor pair of the round
  // chose the child it hangs below: one mask compare per lane (as in dij_compute_path)
  const int d = 32 - __clz(lane + 1), q = lane + 1 - (1 << (d - 1));
  unsigned anc = 0u, want = 0u;
  for (int k = 1; k < d && d <= 5; ++k) { const int a = (1 << (k - 1)) - 1 + (q >> (d - k)); anc |= 1u << a; if ((q >> (d - k - 1)) & 1) want |= 1u << a; }
  while (((pos + 1) << 5) + 30 < SMO) {
    const int L = ((pos + 1) << d) - 1 + 2 * q;
    const bool have = lane < 31 && L < last, have_r = have && L + 1 < last;
    double lf = 0.0, rf = 0.0; int li = 0, ri = 0;
    if (have) { lf = sf[L]; li = si[L]; }
    if (have_r) { rf = sf[L + 1]; ri = si[L + 1]; }
    const bool right = have_r && !(lf < rf);
    const double cf = right ? rf : lf; const int ci = right ? ri : li;
    const unsigned rm = __ballot_sync(AVP_FULL_MASK, right);
    const bool on = have && ((rm & anc) == want);
    const unsigned om = __ballot_sync(AVP_FULL_MASK, on);
    if (on) { const int par = ((pos + 1) << (d - 1)) - 1 + q; sf[par] = cf; si[par] = ci; nodes[ci].hpos = par; }
    __syncwarp();
    if (om == 0u) { bottom = true; break; }
    const int t = 31 - __clz(om), dt = 32 - __clz(t + 1), qt = t + 1 - (1 << (dt - 1));
    pos = ((pos + 1) << dt) - 1 + 2 * qt + (int)((rm >> t) & 1u);
    if (dt < 5) { bottom = true; break; }
  }
#endif
  if (!bottom) {
    if (lane == 0) {
      int child = 2 * pos + 1, lvl = 0;
      while (child < last) {
        const int right = child + 1;
        if (2 * child + 2 >= SMO && (lvl++ & 3) == 0) oh_prefetch_subtree<SMO>(ge, pos, last);
        double cf; int ci; oh_get<SMO>(sf, si, ge, child, cf, ci);
        if (right < last) { double rf; int ri; oh_get<SMO>(sf, si, ge, right, rf, ri); if (!(cf < rf)) { child = right; cf = rf; ci = ri; } }
        OH_SET(pos, cf, ci); pos = child; child = 2 * pos + 1;
      }
    }
    pos = __shfl_sync(AVP_FULL_MASK, pos, 0);
  }
  fi = shfl_d(fi, 0); item = __shfl_sync(AVP_FULL_MASK, item, 0);
  oh_siftdown_warp<SMO>(sf, si, ge, nodes, pos, fi, item, lane);
}
// in-place key update of an entry (hybrid_a_star.py:224-230: no re-heapify)
template <int SMO>
__device__ __forceinline__ void oh_set_key(double *sf, OEnt *ge, int hpos, double f) { if (hpos < SMO) sf[hpos] = f; else ge[hpos].f = f; }

// hybrid_a_star.py:243-259
__device__ __forceinline__ double node_cost(const avp_config &c, bool gear, double theta, double father_theta, bool father_gear) {
  double cost_gear = 0; if (gear != father_gear) cost_gear = c.cost_gear;
  const double cost_heading = fabs(theta - father_theta);
  const double cost = cost_gear + c.cost_heading_change * cost_heading;
  return c.cost_scale * cost;
}

// successor i of a node (hybrid_a_star.py:134-151)
__device__ __forceinline__ void child_pose(const avp_config &cfg, const Node &cn, int i, int nchild, double &x_, double &y_, double &th) {
  const double tn = cfg.tan_steer[i % cfg.steering_angle_num];
  const double speed = (i < nchild / 2.0) ? cfg.max_v : -cfg.max_v;
  const double td = speed * cfg.dt;
  th = pi_2_pi(cn.theta + (cfg.max_v * tn) / cfg.lw * cfg.dt);
  x_ = cn.x + td * d_cos(th); y_ = cn.y + td * d_sin(th);
}

// ------------------------------------------------------------------------------------------
enum { CTL_RUN = 0, CTL_EXIT = 1 };

// shared-memory entries of the open heap per CTA width (dynamic shared memory: 12 bytes per entry)
#ifndef AVP_SMO_WIDE
#define AVP_SMO_WIDE 1024      // (2048: C2 +2 % step time together with two successors per selection item; 8192: +3 %)
#endif
__host__ __device__ constexpr int avp_sm_open(int block) { return block >= 256 ? AVP_SMO_WIDE : (block >= 128 ? 1024 : 512); }   // more shared memory here costs L1 hit rate (libm tables, nodes)
