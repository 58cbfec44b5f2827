// avp_kernels.cuh -- kernels of the hybrid-A* hot path (sm_100a).
//
//   k_raster / k_count_cols / k_fill_cells : Map.detect_obstacle_edge (costmap.py:197-261) and the
//                                            np.where(cost_map == 255) cell list (collision_check.py:55-57)
//   k_check_batch                          : distance_checker / two_circle_checker .check
//   k_expand_pure                          : pure part of hybrid_a_star.expand_node + rs length
//   k_rs_optimal                           : rs_curve.calc_optimal_path
//   k_search                               : PathPlanner.a_star_plan, whole search per scenario
//                                            (persistent CTAs pulling scenarios from an atomic counter)
#pragma once
#include "avp_dev.cuh"

#ifndef AVP_BLOCK
#define AVP_BLOCK 64          // threads per search CTA
#endif
#ifndef AVP_SM_HEAP
#define AVP_SM_HEAP 1024      // Dijkstra heap entries kept in shared memory (the rest spills to L2/HBM)
#endif
#define AVP_NWARPS (AVP_BLOCK / 32)
#define AVP_COURSE_CAP 1024
#define AVP_HQ_CAP 256
#ifndef AVP_NCHILD_MAX
#define AVP_NCHILD_MAX 10      // 2 * steering_angle_num supported by the search kernel's shared-memory layout
#endif

struct __align__(16) Node {
  double x, y, theta, f, g, h;
  int32_t parent;
  uint8_t forward, steer_idx, in_open, in_closed;
  int32_t pad0, pad1;
};
static_assert(sizeof(Node) == 64, "Node must be 64 bytes");

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
  // per-slot (persistent CTA) workspaces
  unsigned long long *dheap; int dheap_cap;
  Node *nodes; int node_cap;
  int32_t *oheap;
  int32_t *htab; int htab_size;   // power of two
  double *course;                 // 3*AVP_COURSE_CAP doubles per slot
  int32_t *course_dir;
  // results per scenario
  avp_plan_summary *sums;
  double *paths; int cap_path;
  int32_t *pops; int cap_pops;
  int32_t *hq_log;                // n * AVP_HQ_CAP * 3, may be NULL
  int *work_counter;
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

__global__ void k_fill_cells(int n_scen, const ScenDev *scen, const uint8_t *cost, const int32_t *col_start, double2 *cells) {
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
__global__ void k_check_batch(avp_config cfg, const ScenDev *scen, int s, const double2 *cells, const int32_t *col_start,
                              int m, const double *poses, uint8_t *out) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (w >= m) return;
  const ScenDev &S = scen[s];
  const bool hit = check_pose_warp(cfg, S, cells + S.cell_off, col_start + S.col_off, poses[3 * w], poses[3 * w + 1], poses[3 * w + 2]);
  if ((threadIdx.x & 31) == 0) out[w] = hit ? 1 : 0;
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
    if (inst < RS_NINST) { double t, u, v; ok = rs_eval_instance(inst, Q, t, u, v); if (ok) { cand[inst].t = t; cand[inst].u = u; cand[inst].v = v; } }
    valid |= (unsigned long long)__ballot_sync(AVP_FULL_MASK, ok) << base;
  }
  __syncwarp();
  rs_select(cand, valid, xy_np, phi_np, maxc, best);   // every lane computes the same result
}

// one CTA, one warp per successor (hybrid_a_star.py:133-151,185-204 + rs length)
__global__ void k_expand_pure(avp_config cfg, const ScenDev *scen, int s, const double2 *cells, const int32_t *col_start,
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
__global__ void k_rs_optimal(int m, const double *q, double maxc, double step_size, int xy_np, int phi_np,
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
// is an unsigned 64-bit compare.  The first AVP_SM_HEAP entries live in shared memory.

struct DijCtx {
  const ScenDev *S; const uint8_t *cost;
  int32_t *hval, *ost; double *gx, *gy;
  unsigned long long *sheap, *gheap; int gcap;
  int hn; int closed_len; int status;
};

__device__ __forceinline__ unsigned long long hp_get(const DijCtx &D, int i) { return i < AVP_SM_HEAP ? D.sheap[i] : D.gheap[i - AVP_SM_HEAP]; }
__device__ __forceinline__ void hp_set(DijCtx &D, int i, unsigned long long v) { if (i < AVP_SM_HEAP) D.sheap[i] = v; else D.gheap[i - AVP_SM_HEAP] = v; }

// heapq._siftdown (Lib/heapq.py:207-219)
__device__ __forceinline__ void hp_siftdown(DijCtx &D, int start, int pos) {
  const unsigned long long item = hp_get(D, pos);
  while (pos > start) {
    const int parent = (pos - 1) >> 1;
    const unsigned long long p = hp_get(D, parent);
    if (item < p) { hp_set(D, pos, p); pos = parent; continue; }
    break;
  }
  hp_set(D, pos, item);
}
// heapq._siftup (Lib/heapq.py:260-278)
__device__ __forceinline__ void hp_siftup(DijCtx &D, int pos) {
  const int n = D.hn, start = pos;
  const unsigned long long item = hp_get(D, pos);
  int child = 2 * pos + 1;
  while (child < n) {
    const int right = child + 1;
    unsigned long long cv = hp_get(D, child);
    if (right < n) { const unsigned long long rv = hp_get(D, right); if (!(cv < rv)) { child = right; cv = rv; } }
    hp_set(D, pos, cv); pos = child; child = 2 * pos + 1;
  }
  hp_set(D, pos, item);
  hp_siftdown(D, start, pos);
}

// compute_h.py:237-255
__device__ __forceinline__ bool dij_is_obstacle(const ScenDev &S, const uint8_t *cm, double gx, double gy) {
  long long xi = (long long)floor((gx - S.b[0]) / S.dx) - 1, yi = (long long)floor((gy - S.b[2]) / S.dy) - 1;
  if (xi >= S.mx) xi = S.mx - 1; if (yi >= S.my) yi = S.my - 1;
  if (xi < 0) xi += S.nx; if (yi < 0) yi += S.ny;           // python negative indexing
  if (xi < 0 || xi >= S.nx || yi < 0 || yi >= S.ny) return false;
  return cm[(size_t)xi * S.ny + yi] == 255;
}

__device__ __forceinline__ double shfl_d(double v, int src) {
  int lo = __double2loint(v), hi = __double2hiint(v);
  lo = __shfl_sync(AVP_FULL_MASK, lo, src); hi = __shfl_sync(AVP_FULL_MASK, hi, src);
  return __hiloint2double(hi, lo);
}

// Dijkstra.compute_path (compute_h.py:198-214); warp-collective.  Returns the popped distance
// of the target cell, or -1 if the queue ran dry (reference: blocks forever) / capacity.
__device__ __noinline__ int dij_compute_path(DijCtx &D, double node_x, double node_y, long long *term_out) {
  const ScenDev &S = *D.S;
  const int lane = threadIdx.x & 31;
  const long long term = map_index(S, node_x, node_y);
  if (term_out) *term_out = term;
  double cur_x = S.pose[3], cur_y = S.pose[4];
  int cur_dist = 0;
  const long long gid = map_index(S, cur_x, cur_y);
  if (lane == 0) { D.closed_len++; if (gid >= 0 && gid < S.n_ids && D.hval[gid] < 0) D.hval[gid] = 0; }   // initial_map (:50-72)
  __syncwarp();
  for (long long guard = 0;; ++guard) {
    if (guard > 4ll * S.n_ids + 1024) { if (lane == 0) D.status = AVP_CAPACITY; __syncwarp(); return -1; }   // cannot happen: every id is pushed at most twice
    // update_openlist (compute_h.py:84-195): the 8 neighbours evaluated by lanes 0..7
    bool valid = false; int prio = 0, st = 0; long long nid = 0; double ngx = 0.0, ngy = 0.0;
    if (lane < 8) {
      const int ddx = (lane == 0 || lane == 3 || lane == 5) ? -1 : ((lane == 2 || lane == 4 || lane == 7) ? 1 : 0);
      const int ddy = (lane < 3) ? 1 : ((lane < 5) ? 0 : -1);
      ngx = ddx < 0 ? cur_x - S.dx : (ddx > 0 ? cur_x + S.dx : cur_x);
      ngy = ddy < 0 ? cur_y - S.dy : (ddy > 0 ? cur_y + S.dy : cur_y);
      if (!dij_is_obstacle(S, D.cost, ngx, ngy)) {
        bool ok = true;
        if (ddx < 0 && !(ngx >= S.b[0])) ok = false; if (ddx > 0 && !(ngx <= S.b[1])) ok = false;
        if (ddy > 0 && !(ngy <= S.b[3])) ok = false; if (ddy < 0 && !(ngy >= S.b[2])) ok = false;
        if (ok) {
          nid = map_index(S, ngx, ngy);
          if (nid >= 0 && nid < S.n_ids) { valid = true; st = D.ost[nid]; prio = cur_dist + ((ddx && ddy) ? 14 : 10); }
        }
      }
    }
    const unsigned vmask = __ballot_sync(AVP_FULL_MASK, valid) & 0xffu;
    for (int k = 0; k < 8; ++k) {             // add_grid_to_openlist (:216-235), in the reference's order
      if (!((vmask >> k) & 1u)) continue;
      const int st_k = __shfl_sync(AVP_FULL_MASK, st, k), prio_k = __shfl_sync(AVP_FULL_MASK, prio, k);
      const int nid_k = (int)__shfl_sync(AVP_FULL_MASK, (int)nid, k);
      if (st_k == -1) {                        // first visit: heappush
        const double x_k = shfl_d(ngx, k), y_k = shfl_d(ngy, k);
        if (lane == 0) {
          if (D.hn >= AVP_SM_HEAP + D.gcap) D.status = AVP_CAPACITY;
          else {
            hp_set(D, D.hn, ((unsigned long long)(unsigned)prio_k << 32) | (unsigned)nid_k);
            D.hn++; hp_siftdown(D, 0, D.hn - 1);
            D.ost[nid_k] = prio_k; D.gx[nid_k] = x_k; D.gy[nid_k] = y_k;
          }
        }
      } else if (st_k >= 0 && st_k > prio_k) {  // still queued with a larger distance: overwrite IN PLACE, no re-sift (:222-228)
        __syncwarp();
        int pos = 0x7fffffff;
        for (int i = lane; i < D.hn; i += 32) if ((unsigned)hp_get(D, i) == (unsigned)nid_k) { pos = i; break; }
        for (int o = 16; o > 0; o >>= 1) pos = min(pos, __shfl_xor_sync(AVP_FULL_MASK, pos, o));
        if (lane == 0 && pos != 0x7fffffff) {
          hp_set(D, pos, ((unsigned long long)(unsigned)prio_k << 32) | (unsigned)nid_k);
          D.ost[nid_k] = prio_k;
        }
      }
      __syncwarp();
    }
    const int st_now = __shfl_sync(AVP_FULL_MASK, D.status, 0), hn_now = __shfl_sync(AVP_FULL_MASK, D.hn, 0);
    if (st_now) return -1;
    if (hn_now == 0) { if (lane == 0) D.status = AVP_H_UNREACHABLE; __syncwarp(); return -1; }
    // update_closedlist (:74-82): heappop
    unsigned long long top = 0ull;
    if (lane == 0) {
      top = hp_get(D, 0);
      const unsigned long long last = hp_get(D, D.hn - 1);
      D.hn--;
      if (D.hn > 0) { hp_set(D, 0, last); hp_siftup(D, 0); }
      const int id = (int)(unsigned)top, dist = (int)(top >> 32);
      D.ost[id] = -2; D.closed_len++;
      if (D.hval[id] < 0) D.hval[id] = dist;
    }
    const int cur_id = (int)__shfl_sync(AVP_FULL_MASK, (unsigned)top, 0);
    cur_dist = (int)__shfl_sync(AVP_FULL_MASK, (unsigned)(top >> 32), 0);
    __syncwarp();
    if ((long long)cur_id == term) return cur_dist;
    cur_x = D.gx[cur_id]; cur_y = D.gy[cur_id];
  }
}

// ------------------------------------------------------------------------------------------
// hybrid A* containers (per slot, global memory)

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
__device__ __forceinline__ void htab_insert(int32_t *htab, int mask, const Node *nodes, int idx) {
  const Node &n = nodes[idx];
  unsigned long long p = pose_hash(n.x, n.y, n.theta);
  for (int probes = 0; probes <= mask; ++probes, ++p) if (htab[p & mask] < 0) { htab[p & mask] = idx; return; }
}
// open_list: heapq of node indices ordered by Node.__lt__ (f only, hybrid_a_star.py:61-68)
__device__ __forceinline__ void open_siftdown(int32_t *heap, const Node *nodes, int start, int pos) {
  const int item = heap[pos]; const double fi = nodes[item].f;
  while (pos > start) {
    const int parent = (pos - 1) >> 1; const int pv = heap[parent];
    if (fi < nodes[pv].f) { heap[pos] = pv; pos = parent; continue; }
    break;
  }
  heap[pos] = item;
}
__device__ __forceinline__ void open_siftup(int32_t *heap, const Node *nodes, int n, int pos) {
  const int start = pos, item = heap[pos];
  int child = 2 * pos + 1;
  while (child < n) {
    const int right = child + 1;
    int cv = heap[child];
    if (right < n) { const int rv = heap[right]; if (!(nodes[cv].f < nodes[rv].f)) { child = right; cv = rv; } }
    heap[pos] = cv; pos = child; child = 2 * pos + 1;
  }
  heap[pos] = item;
  open_siftdown(heap, nodes, start, pos);
}

// hybrid_a_star.py:243-259
__device__ __forceinline__ double node_cost(const avp_config &c, bool gear, double theta, double father_theta, bool father_gear) {
  double cost_gear = 0; if (gear != father_gear) cost_gear = c.cost_gear;
  const double cost_heading = fabs(theta - father_theta);
  const double cost = cost_gear + c.cost_heading_change * cost_heading;
  return c.cost_scale * cost;
}

// ------------------------------------------------------------------------------------------
// the search kernel: PathPlanner.a_star_plan (path_planner.py:58-110) for one scenario per CTA.

enum { CTL_RUN = 0, CTL_EXIT = 1 };

__global__ void __launch_bounds__(AVP_BLOCK) k_search(KParams P) {
  __shared__ unsigned long long s_heap[AVP_SM_HEAP];
  __shared__ RsCand s_cand[AVP_NCHILD_MAX + 1][RS_NINST];
  __shared__ unsigned long long s_valid[AVP_NCHILD_MAX + 1];
  __shared__ double s_cpose[AVP_NCHILD_MAX][3];
  __shared__ double s_rsL[AVP_NCHILD_MAX];
  __shared__ int s_found[AVP_NCHILD_MAX], s_coll[AVP_NCHILD_MAX], s_need[AVP_NCHILD_MAX], s_rsok[AVP_NCHILD_MAX], s_skip[AVP_NCHILD_MAX];
  __shared__ int s_scen, s_ctl, s_cur, s_in_radius, s_npts, s_shot_coll, s_shot_bad;
  __shared__ int s_G, s_nclosed, s_npops, s_on, s_status, s_nhq, s_nhcalls;
  __shared__ RsBest s_best;
  __shared__ DijCtx s_D;

  const avp_config &cfg = P.cfg;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int slot = blockIdx.x;
  const int nchild = 2 * cfg.steering_angle_num;
  const double maxc = 1 / cfg.min_radius_turn;
  Node *nodes = P.nodes + (size_t)slot * P.node_cap;
  int32_t *oheap = P.oheap + (size_t)slot * P.node_cap;
  int32_t *htab = P.htab + (size_t)slot * P.htab_size;
  const int hmask = P.htab_size - 1;
  double *CX = P.course + (size_t)slot * 3 * AVP_COURSE_CAP, *CY = CX + AVP_COURSE_CAP, *CYAW = CY + AVP_COURSE_CAP;
  int32_t *CDIR = P.course_dir + (size_t)slot * AVP_COURSE_CAP;

  for (;;) {
    if (tid == 0) s_scen = atomicAdd(P.work_counter, 1);
    __syncthreads();
    const int sc = s_scen;
    if (sc >= P.n_scen) break;
    const ScenDev &S = P.scen[sc];
    const double2 *cells = P.cells + S.cell_off;
    const int32_t *col_start = P.col_start + S.col_off;
    int32_t *hval = P.hval + S.id_off, *ost = P.ost + S.id_off;
    const double goal[3] = {S.pose[3], S.pose[4], pi_2_pi(S.pose[5])};
    int32_t *pops = P.pops ? P.pops + (size_t)sc * P.cap_pops : nullptr;
    int32_t *hql = P.hq_log ? P.hq_log + (size_t)sc * AVP_HQ_CAP * 3 : nullptr;
    int *dbg = P.dbg ? P.dbg + (size_t)sc * 8 : nullptr;
    const long long t_start = clock64();
    if (dbg && tid == 0) { dbg[0] = 1; dbg[1] = 0; }

    // ---- per-scenario initialisation (all threads)
    for (int i = tid; i < S.n_ids; i += AVP_BLOCK) { hval[i] = -1; ost[i] = -1; }
    for (int i = tid; i < P.htab_size; i += AVP_BLOCK) htab[i] = -1;
    if (tid == 0) {
      s_D.S = &S; s_D.cost = P.cost + S.cost_off; s_D.hval = hval; s_D.ost = ost;
      s_D.gx = P.gx + S.id_off; s_D.gy = P.gy + S.id_off;
      s_D.sheap = s_heap; s_D.gheap = P.dheap + (size_t)slot * P.dheap_cap; s_D.gcap = P.dheap_cap;
      s_D.hn = 0; s_D.closed_len = 0; s_D.status = 0;
      s_G = 0; s_nclosed = 0; s_npops = 0; s_on = 0; s_nhq = 0; s_nhcalls = 0;
      s_status = S.raster_error ? AVP_RASTER_AMBIGUOUS : 0;
      s_cur = -1; s_in_radius = 0; s_shot_coll = 0; s_npts = 0; s_best.ok = 0;
    }
    __syncthreads();

    // ---- hybrid_a_star.__init__: eager Dijkstra to the start cell (hybrid_a_star.py:89-91), root node (:102-112)
    if (warp == 0 && s_status == 0) {
      long long term;
      const int d = dij_compute_path(s_D, S.pose[0], S.pose[1], &term);
      if (lane == 0) {
        if (hql && s_nhq < AVP_HQ_CAP) { hql[3 * s_nhq] = (int)term; hql[3 * s_nhq + 1] = d; hql[3 * s_nhq + 2] = s_D.closed_len; }
        s_nhq++;
        if (d < 0) s_status = s_D.status ? s_D.status : AVP_H_UNREACHABLE;
        Node r; r.x = S.pose[0]; r.y = S.pose[1]; r.theta = pi_2_pi(S.pose[2]); r.f = 0; r.g = 0; r.h = 0; r.parent = -1;
        r.forward = 1; r.steer_idx = 0; r.in_open = 1; r.in_closed = 0; r.pad0 = 0; r.pad1 = 0;
        nodes[0] = r;
        htab_insert(htab, hmask, nodes, 0);
        oheap[0] = 0; s_on = 1;
      }
    }

    // ---- main loop (path_planner.py:68-98)
    bool reached = false;
    if (dbg && tid == 0) dbg[0] = 2;
    for (;;) {
      __syncthreads();
      if (tid == 0) {
        if (dbg) { dbg[1] = s_npops; dbg[2] = s_D.closed_len; dbg[3] = s_on; }
        if (P.watchdog_cycles > 0 && clock64() - t_start > P.watchdog_cycles && s_status == 0) s_status = AVP_CAPACITY;
        if (s_status != 0 || s_on == 0) s_ctl = CTL_EXIT;
        else if (s_npops >= cfg.max_pops) { s_status = AVP_CAPACITY; s_ctl = CTL_EXIT; }
        else {
          // open_list.get() == heapq.heappop
          const int last = oheap[--s_on]; int ret = last;
          if (s_on) { ret = oheap[0]; oheap[0] = last; open_siftup(oheap, nodes, s_on, 0); }
          s_cur = ret;
          if (pops && s_npops < P.cap_pops) pops[s_npops] = ret;
          s_npops++;
          const Node &cn = nodes[ret];
          const double distance = sqrt(d_pow2(cn.x - goal[0]) + d_pow2(cn.y - goal[1]));   // hybrid_a_star.py:308-309 (** 2 == libm pow)
          s_in_radius = distance < cfg.flag_radius;
          s_shot_coll = 0; s_shot_bad = 0; s_npts = 0; s_best.ok = 0;
          s_ctl = CTL_RUN;
        }
        for (int i = 0; i <= nchild; ++i) s_valid[i] = 0ull;
      }
      __syncthreads();
      if (s_ctl == CTL_EXIT) break;
      if (dbg && tid == 0) dbg[0] = 3;
      const int cur = s_cur;
      const Node cn = nodes[cur];
      const int phi_np = cur != 0;           // root theta is a Python float (see oracle generate_path)

      // phase 1: rs word instances of the goal shot; successor poses + closed/open lookups
      if (s_in_radius) {
        const double q0[3] = {cn.x, cn.y, cn.theta};
        RsQuery Q; rs_query(q0, goal, maxc, Q);
        for (int inst = tid; inst < RS_NINST; inst += AVP_BLOCK) {
          double t, u, v;
          if (rs_eval_instance(inst, Q, t, u, v)) { s_cand[nchild][inst].t = t; s_cand[nchild][inst].u = u; s_cand[nchild][inst].v = v; atomicOr(&s_valid[nchild], 1ull << inst); }
        }
      }
      for (int i = tid; i < nchild; i += AVP_BLOCK) {            // hybrid_a_star.py:134-165
        const double tn = cfg.tan_steer[i % cfg.steering_angle_num];
        const double speed = (i < nchild / 2.0) ? cfg.max_v : -cfg.max_v;
        const double td = speed * cfg.dt;
        double th = cn.theta + (cfg.max_v * tn) / cfg.lw * cfg.dt;
        th = pi_2_pi(th);
        const double x_ = cn.x + td * d_cos(th), y_ = cn.y + td * d_sin(th);
        s_cpose[i][0] = x_; s_cpose[i][1] = y_; s_cpose[i][2] = th;
        const int found = htab_find(htab, hmask, nodes, x_, y_, th);
        const bool in_closed = found >= 0 && nodes[found].in_closed;
        const bool oob = (s_nclosed > 0) && (x_ > S.b[1] || x_ < S.b[0] || y_ > S.b[3] || y_ < S.b[2]);
        s_found[i] = found; s_skip[i] = (in_closed || oob) ? 1 : 0; s_coll[i] = 0; s_need[i] = 0; s_rsok[i] = 0;
      }
      __syncthreads();

      if (dbg && tid == 0) dbg[0] = 4;
      // phase 2: thread 0 selects the shot word and lays out its course; meanwhile the other
      // lanes/warps collision-check the sub-steps of new successors (hybrid_a_star.py:185-204)
      if (tid == 0 && s_in_radius) {
        RsBest b; rs_select(s_cand[nchild], s_valid[nchild], 1, phi_np, maxc, b);
        if (!b.ok || b.degenerate) s_shot_bad = 1;
        else {
          const double q0[3] = {cn.x, cn.y, cn.theta};
          const int np_ = rs_course(b, maxc, 0.5, q0, AVP_COURSE_CAP, CX, CY, CYAW, CDIR);
          if (np_ < 0) s_shot_bad = 2; else s_npts = np_;
          s_best = b;
        }
      }
      {
        const int w0 = (AVP_NWARPS > 1) ? warp - 1 : 0, nw = (AVP_NWARPS > 1) ? AVP_NWARPS - 1 : 1;
        if (AVP_NWARPS == 1 || warp >= 1) {
          for (int i = w0; i < nchild; i += nw) {
            if (s_skip[i] || s_found[i] >= 0) continue;
            const double tn = cfg.tan_steer[i % cfg.steering_angle_num];
            const double speed = (i < nchild / 2.0) ? cfg.max_v : -cfg.max_v;
            int coll = 0;
            for (int k = 0; k < cfg.n_substeps; ++k) {
              const double td_i = speed * cfg.ddt * (k + 1);
              double th_i = cn.theta + (cfg.max_v * tn) / cfg.lw * cfg.ddt * (k + 1);
              th_i = pi_2_pi(th_i);
              const double x_i = cn.x + td_i * d_cos(th_i), y_i = cn.y + td_i * d_sin(th_i);
              if (check_pose_warp(cfg, S, cells, col_start, x_i, y_i, th_i)) { coll = 1; break; }
            }
            if (lane == 0) s_coll[i] = coll;
          }
        }
      }
      __syncthreads();
      if (s_shot_bad) { if (tid == 0) s_status = (s_shot_bad == 1) ? AVP_RS_DEGENERATE : AVP_CAPACITY; continue; }

      if (dbg && tid == 0) dbg[0] = 5;
      // phase 3: collision check of the shot's course (hybrid_a_star.py:334-347)
      if (s_in_radius) {
        const int npts = s_npts;
        for (int i = warp; i < npts; i += AVP_NWARPS) {
          const int stop = __shfl_sync(AVP_FULL_MASK, *(volatile int *)&s_shot_coll, 0);   // warp-uniform early exit
          if (stop) break;
          if (check_pose_warp(cfg, S, cells, col_start, CX[i], CY[i], pi_2_pi(CYAW[i]))) { if (lane == 0) s_shot_coll = 1; }
        }
      }
      __syncthreads();
      if (s_in_radius && !s_shot_coll) { reached = true; break; }     // path_planner.py:86-88

      if (dbg && tid == 0) dbg[0] = 6;
      // phase 4: rs lengths of the successors that will be scored (hybrid_a_star.py:286-294)
      for (int i = tid; i < nchild; i += AVP_BLOCK) {
        const int f = s_found[i];
        s_need[i] = (!s_skip[i]) && ((f < 0 && !s_coll[i]) || (f >= 0));
      }
      __syncthreads();
      for (int item = tid; item < nchild * RS_NINST; item += AVP_BLOCK) {
        const int i = item / RS_NINST, inst = item - i * RS_NINST;
        if (!s_need[i]) continue;
        const double q0[3] = {s_cpose[i][0], s_cpose[i][1], s_cpose[i][2]};
        RsQuery Q; rs_query(q0, goal, maxc, Q);
        double t, u, v;
        if (rs_eval_instance(inst, Q, t, u, v)) { s_cand[i][inst].t = t; s_cand[i][inst].u = u; s_cand[i][inst].v = v; atomicOr(&s_valid[i], 1ull << inst); }
      }
      __syncthreads();
      for (int i = tid; i < nchild; i += AVP_BLOCK) {
        if (!s_need[i]) continue;
        RsBest b; rs_select(s_cand[i], s_valid[i], 1, 1, maxc, b);
        s_rsok[i] = (b.ok && !b.degenerate) ? 1 : 0;
        s_rsL[i] = b.ok ? b.L / maxc : 0.0;
      }
      __syncthreads();

      if (dbg && tid == 0) dbg[0] = 7;
      // phase 5: sequential commit in slot order by warp 0 (hybrid_a_star.py:154-239)
      if (warp == 0) {
        for (int i = 0; i < nchild; ++i) {
          if (s_skip[i]) continue;
          if (__shfl_sync(AVP_FULL_MASK, s_status, 0)) break;
          const int found = s_found[i];
          const double x_ = s_cpose[i][0], y_ = s_cpose[i][1], th = s_cpose[i][2];
          const bool fwd = i < nchild / 2.0;
          int child = found, miss = 0; int hv = -1;
          if (lane == 0) {
            if (found < 0) {
              child = s_G + i + 1;
              if (child >= P.node_cap) { s_status = AVP_CAPACITY; }
              else {
                Node n; n.x = x_; n.y = y_; n.theta = th; n.f = 0; n.g = 0; n.h = 0; n.parent = cur;
                n.forward = fwd ? 1 : 0; n.steer_idx = (uint8_t)(i % cfg.steering_angle_num); n.in_open = 0;
                n.in_closed = s_coll[i] ? 1 : 0; n.pad0 = 0; n.pad1 = 0;
                nodes[child] = n;
                htab_insert(htab, hmask, nodes, child);
                if (s_coll[i]) s_nclosed++;
              }
            }
            if (!s_status && !(found < 0 && s_coll[i])) {
              if (!s_rsok[i]) s_status = AVP_RS_DEGENERATE;
              else {
                s_nhcalls++;
                const long long id = map_index(S, x_, y_);               // calc_node_heuristic (:261-283)
                hv = (id >= 0 && id < S.n_ids) ? hval[id] : -1;
                miss = hv < 0;
              }
            }
          }
          __syncwarp();
          const int st = __shfl_sync(AVP_FULL_MASK, s_status, 0);
          if (st) break;
          if (found < 0 && s_coll[i]) continue;
          miss = __shfl_sync(AVP_FULL_MASK, miss, 0);
          if (miss) {
            long long term;
            const int d = dij_compute_path(s_D, x_, y_, &term);
            if (lane == 0) {
              if (hql && s_nhq < AVP_HQ_CAP) { hql[3 * s_nhq] = (int)term; hql[3 * s_nhq + 1] = d; hql[3 * s_nhq + 2] = s_D.closed_len; }
              s_nhq++;
              if (d < 0) s_status = s_D.status ? s_D.status : AVP_H_UNREACHABLE;
              hv = d;
            }
            __syncwarp();
            if (__shfl_sync(AVP_FULL_MASK, s_status, 0)) break;
          }
          if (lane == 0) {
            const double h1 = hv / 100.0, h2 = s_rsL[i];
            const double h = (h2 > h1) ? h2 : h1;                         // max(h_value_1, h_value_2) (:294-296)
            child = (found < 0) ? s_G + i + 1 : found;
            Node &n = nodes[child];
            if (found < 0) {                                              // :206-216
              n.g = node_cost(cfg, fwd, th, cn.theta, cn.forward != 0);
              n.h = h; n.f = n.g + n.h;
              n.in_open = 1;
              oheap[s_on++] = child; open_siftdown(oheap, nodes, 0, s_on - 1);
            } else {                                                      // :219-230 (in-place, no re-heapify)
              const double new_g = node_cost(cfg, n.forward != 0, n.theta, cn.theta, cn.forward != 0);
              const double new_f = h + new_g;
              if (new_f < n.f) { n.f = new_f; n.g = new_g; n.h = h; n.parent = cur; n.forward = fwd ? 1 : 0; n.steer_idx = (uint8_t)(i % cfg.steering_angle_num); }
            }
          }
          __syncwarp();
        }
        if (lane == 0 && !s_status) { nodes[cur].in_closed = 1; nodes[cur].in_open = 0; s_nclosed++; s_G += nchild; }   // :235-239
      }
    }
    __syncthreads();

    if (dbg && tid == 0) dbg[0] = 8;
    // ---- finish: summary + finish_path (hybrid_a_star.py:351-389) + rs tail (path_planner.py:100-108)
    if (tid == 0) {
      avp_plan_summary &R = P.sums[sc];
      int status = s_status;
      if (!status && !reached) status = (s_in_radius && s_best.ok) ? AVP_OPEN_EXHAUSTED_RS : AVP_OPEN_EXHAUSTED;
      R.status = status; R.n_pops = s_npops; R.global_index = s_G; R.n_closed = s_nclosed; R.n_open = s_on;
      R.last_index = s_cur; R.n_hq = s_nhq; R.h_closed = s_D.closed_len; R.nx = S.nx; R.ny = S.ny; R.n_obs = S.n_obs;
      R.n_hcalls = s_nhcalls; R.pitch[0] = S.dx; R.pitch[1] = S.dy;
      for (int i = 0; i < 4; ++i) R.boundary[i] = S.b[i];
      R.origin[0] = S.b[0]; R.origin[1] = S.b[2];
      R.n_astar = 0; R.n_rs = 0; R.n_final = 0; R.rs_nseg = 0; R.rs_L = 0.0;
      for (int i = 0; i < 5; ++i) R.rs_lengths[i] = 0.0;
      for (int i = 0; i < 8; ++i) R.rs_ctypes[i] = 0;
      if (status == AVP_OK || status == AVP_OPEN_EXHAUSTED_RS) {
        double *fp = P.paths + (size_t)sc * P.cap_path * 3;
        int np_ = 0;
        // parent walk; the chain is emitted root-first, so first measure the depth
        int depth = 0; for (int k = s_cur; k != 0; k = nodes[k].parent) ++depth;
        auto push = [&](double px, double py, double pt) { if (np_ < P.cap_path) { fp[3 * np_] = px; fp[3 * np_ + 1] = py; fp[3 * np_ + 2] = pt; } ++np_; };
        push(nodes[0].x, nodes[0].y, nodes[0].theta);
        for (int lvl = 1; lvl <= depth; ++lvl) {
          int ch = s_cur; for (int k = 0; k < depth - lvl; ++k) ch = nodes[ch].parent;
          const Node &c = nodes[ch]; const Node &par = nodes[c.parent];
          for (int j = 0; j < cfg.n_substeps; ++j) {
            const double speed = c.forward ? cfg.max_v : -cfg.max_v;
            const double td_j = speed * cfg.ddt * (j + 1);
            double th_j = par.theta + (cfg.max_v * cfg.tan_steer[c.steer_idx]) / cfg.lw * cfg.ddt * (j + 1);
            th_j = pi_2_pi(th_j);
            push(par.x + td_j * d_cos(th_j), par.y + td_j * d_sin(th_j), th_j);
          }
        }
        R.n_astar = np_;
        for (int i = 1; i < s_npts; ++i) push(CX[i], CY[i], CYAW[i]);
        if (dbg) dbg[0] = 9;
        R.n_final = np_; R.n_rs = s_npts; R.rs_nseg = s_best.n; R.rs_L = s_best.L / maxc;
        for (int i = 0; i < s_best.n; ++i) R.rs_lengths[i] = s_best.len[i] / maxc;
        for (int i = 0; i < 8; ++i) R.rs_ctypes[i] = rs_ct_names[s_best.ct][i];
      }
    }
    __syncthreads();
  }
}
