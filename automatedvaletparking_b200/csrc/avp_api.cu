// avp_api.cu -- host side of the C ABI declared in include/avp_b200.h.
// Plain CUDA runtime; no torch types.  There is no CPU fallback: every entry point that
// computes launches a kernel on the context's device.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <string>
#include <vector>
#include <algorithm>
#include <utility>
#include <time.h>
#include "avp_plan.cuh"

struct avp_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  avp_config cfg;
  std::string err;
  int64_t launches = 0;
  int n_sm = 0, slots = 0;
  // scenarios
  int n = 0;
  std::vector<ScenDev> h_scen;
  ScenDev *d_scen = nullptr;
  int32_t *d_nv = nullptr, *d_vert_off = nullptr; double *d_verts = nullptr;
  uint8_t *d_cost = nullptr; int64_t cost_bytes = 0;
  int32_t *d_col = nullptr; int64_t col_count = 0;
  double2 *d_cells = nullptr; int64_t cell_count = 0; long long *d_cell_total = nullptr;
  bool rasterised = false;
  // byte capacities of the scenario arrays: uploads of a same-sized batch reuse the allocations (cudaMalloc/cudaFree are
  // synchronising and cost far more than the H2D copies of a batch)
  size_t cap_scen = 0, cap_nv = 0, cap_vert_off = 0, cap_verts = 0, cap_cost = 0, cap_col = 0, cap_cells = 0, cap_ids = 0, cap_order = 0;
  // per-id arrays
  int64_t id_count = 0;
  int32_t *d_hval = nullptr, *d_ost = nullptr; double *d_gx = nullptr, *d_gy = nullptr;
  // search workspaces: per scenario (Dijkstra queue, ScenState), per slot (nodes, open heap, exact-pose table), per CTA (course scratch)
  int ws_n = 0, ws_slots = 0, ws_ctas = 0, node_cap = 0, htab_size = 0, dheap_cap = 0;
  NodeShot *d_nshot = nullptr;
  Node *d_nodes = nullptr; OEnt *d_oheap = nullptr; int32_t *d_htab = nullptr; unsigned long long *d_dheap = nullptr;
  ScenState *d_state = nullptr; PlanCtl *d_ctl = nullptr; int32_t *d_queue = nullptr, *d_slot_ring = nullptr; int q_mask = 0, slot_mask = 0;
  int32_t *d_order = nullptr;      // processing order: expensive scenarios (far start-goal pairs) first
  double *d_course = nullptr; int32_t *d_course_dir = nullptr; unsigned char *d_cand = nullptr;
  int narrow_block = 64, narrow64_per_sm = 1;
  int plan_block = 512, plan_grid = 0, ctas_per_sm[2] = {1, 2}, narrow_per_sm = 1; float narrow_ms = 0.f; cudaEvent_t evN = nullptr;
  // results
  avp_plan_summary *d_sums = nullptr; double *d_paths = nullptr; int32_t *d_pops = nullptr, *d_hq = nullptr;
  int cap_path = 0, cap_pops = 0; int res_n = 0;
  double *d_pop_fgh = nullptr; size_t cap_fgh = 0; bool trace_fgh = false;
  long long *d_prof = nullptr, *d_wprof = nullptr;
  int *d_counter = nullptr; int *d_dbg = nullptr; long long watchdog_cycles = 0;
  float pass_ms[2] = {0.f, 0.f}; int n_suspends = 0;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, evM = nullptr, evA = nullptr, evB = nullptr;
  // persistent single-scenario Dijkstra query state (drop-in for compute_h.Dijkstra)
  int dq_scen = -1; unsigned long long *d_dq_gheap = nullptr; DijPersist *d_dq_state = nullptr; int32_t *d_dq_out = nullptr;
  // scratch for the small API kernels
  void *d_scratch = nullptr; size_t scratch_bytes = 0;
};

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_); return -1; } } while (0)
#define FAIL(msg) do { ctx->err = (msg); return -2; } while (0)

static void free_dev(void *p) { if (p) cudaFree(p); }

// grow-only device buffer: reallocates when `bytes` exceeds the capacity (or is far below it)
template <class T>
static cudaError_t ensure_dev(T **p, size_t *cap, size_t bytes) {
  if (*p && bytes <= *cap && bytes * 4 >= *cap) return cudaSuccess;
  if (*p) cudaFree(*p);
  *p = nullptr; *cap = 0;
  cudaError_t e = cudaMalloc((void **)p, bytes ? bytes : 16);
  if (e == cudaSuccess) *cap = bytes ? bytes : 16;
  return e;
}

static int ensure_scratch(avp_ctx *ctx, size_t bytes) {
  if (bytes <= ctx->scratch_bytes) return 0;
  free_dev(ctx->d_scratch); ctx->d_scratch = nullptr; ctx->scratch_bytes = 0;
  CK(cudaMalloc(&ctx->d_scratch, bytes));
  ctx->scratch_bytes = bytes;
  return 0;
}

extern "C" int avp_create(int device_id, const avp_config *cfg, avp_ctx **out) {
  if (!cfg || !out) return -3;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) return -4;      // no CUDA device: fail loudly
  if (device_id < 0 || device_id >= ndev) return -5;
  if (cfg->steering_angle_num < 1 || 2 * cfg->steering_angle_num > AVP_NCHILD_MAX) return -6;
  avp_ctx *ctx = new avp_ctx();
  ctx->device = device_id; ctx->cfg = *cfg;
  if (cudaSetDevice(device_id) != cudaSuccess) { delete ctx; return -7; }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device_id) != cudaSuccess) { delete ctx; return -8; }
  ctx->n_sm = prop.multiProcessorCount;
  if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return -9; }
  cudaFuncSetAttribute(k_plan<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, AVP_PLAN_DYN_SMEM(512, AVP_CELL_SMEM));
  cudaFuncSetAttribute(k_plan<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, AVP_PLAN_DYN_SMEM(256, AVP_CELL_SMEM));
  cudaFuncSetAttribute(k_plan<640>, cudaFuncAttributeMaxDynamicSharedMemorySize, AVP_PLAN_DYN_SMEM(640, AVP_CELL_SMEM));
  cudaFuncSetAttribute(k_plan<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, AVP_PLAN_DYN_SMEM(128, 0));
  cudaFuncSetAttribute(k_plan<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, AVP_PLAN_DYN_SMEM(64, 0));
  {
    int o = 0;
    ctx->ctas_per_sm[0] = (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, k_plan<512>, 512, AVP_PLAN_DYN_SMEM(512, AVP_CELL_SMEM)) == cudaSuccess && o > 0) ? o : 1;
    ctx->ctas_per_sm[1] = (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, k_plan<256>, 256, AVP_PLAN_DYN_SMEM(256, AVP_CELL_SMEM)) == cudaSuccess && o > 0) ? o : 1;
    ctx->narrow_per_sm = (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, k_plan<128>, 128, AVP_PLAN_DYN_SMEM(128, 0)) == cudaSuccess && o > 0) ? o : 1;
    ctx->narrow64_per_sm = (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, k_plan<64>, 64, AVP_PLAN_DYN_SMEM(64, 0)) == cudaSuccess && o > 0) ? o : 1;
  }
  ctx->slots = ctx->n_sm * ctx->ctas_per_sm[0];               // persistent grids: multiples of the SM count
  cudaEventCreate(&ctx->ev0); cudaEventCreate(&ctx->ev1); cudaEventCreate(&ctx->evM); cudaEventCreate(&ctx->evN);
  if (cudaMalloc(&ctx->d_counter, sizeof(int)) != cudaSuccess) { delete ctx; return -10; }
  *out = ctx;
  return 0;
}

static void free_scenarios(avp_ctx *ctx) {
  free_dev(ctx->d_scen); free_dev(ctx->d_nv); free_dev(ctx->d_vert_off); free_dev(ctx->d_verts); free_dev(ctx->d_cost);
  free_dev(ctx->d_col); free_dev(ctx->d_cells); free_dev(ctx->d_hval); free_dev(ctx->d_ost); free_dev(ctx->d_gx); free_dev(ctx->d_gy);
  ctx->d_scen = nullptr; ctx->d_nv = ctx->d_vert_off = nullptr; ctx->d_verts = nullptr; ctx->d_cost = nullptr; ctx->d_col = nullptr;
  ctx->d_cells = nullptr; ctx->d_hval = ctx->d_ost = nullptr; ctx->d_gx = ctx->d_gy = nullptr;
  ctx->cap_scen = ctx->cap_nv = ctx->cap_vert_off = ctx->cap_verts = ctx->cap_cost = ctx->cap_col = ctx->cap_cells = ctx->cap_ids = 0;
  ctx->n = 0; ctx->rasterised = false; ctx->dq_scen = -1;
}
static void free_results(avp_ctx *ctx) {
  free_dev(ctx->d_sums); free_dev(ctx->d_paths); free_dev(ctx->d_pops); free_dev(ctx->d_hq); free_dev(ctx->d_dbg); ctx->d_dbg = nullptr; free_dev(ctx->d_prof); ctx->d_prof = nullptr; free_dev(ctx->d_wprof); ctx->d_wprof = nullptr;
  ctx->d_sums = nullptr; ctx->d_paths = nullptr; ctx->d_pops = nullptr; ctx->d_hq = nullptr; ctx->res_n = 0;
}
static void free_ws(avp_ctx *ctx) {
  free_dev(ctx->d_nshot); free_dev(ctx->d_nodes); free_dev(ctx->d_oheap); free_dev(ctx->d_htab); free_dev(ctx->d_dheap);
  free_dev(ctx->d_state); free_dev(ctx->d_ctl); free_dev(ctx->d_queue); free_dev(ctx->d_slot_ring); free_dev(ctx->d_course); free_dev(ctx->d_course_dir); free_dev(ctx->d_cand); ctx->d_cand = nullptr;
  ctx->d_nshot = nullptr; ctx->d_nodes = nullptr; ctx->d_oheap = nullptr; ctx->d_htab = nullptr; ctx->d_dheap = nullptr; ctx->d_state = nullptr; ctx->d_ctl = nullptr;
  ctx->d_queue = ctx->d_slot_ring = nullptr; ctx->d_course = nullptr; ctx->d_course_dir = nullptr; ctx->ws_n = ctx->ws_slots = ctx->ws_ctas = 0;
}

extern "C" int avp_destroy(avp_ctx *ctx) {
  if (!ctx) return 0;
  cudaSetDevice(ctx->device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  free_scenarios(ctx); free_results(ctx); free_ws(ctx);
  free_dev(ctx->d_counter); free_dev(ctx->d_scratch); free_dev(ctx->d_cell_total); free_dev(ctx->d_pop_fgh);
  free_dev(ctx->d_order);
  free_dev(ctx->d_dq_gheap); free_dev(ctx->d_dq_state); free_dev(ctx->d_dq_out);
  if (ctx->ev0) cudaEventDestroy(ctx->ev0); if (ctx->ev1) cudaEventDestroy(ctx->ev1); if (ctx->evM) cudaEventDestroy(ctx->evM); if (ctx->evN) cudaEventDestroy(ctx->evN);
  if (ctx->evA) cudaEventDestroy(ctx->evA); if (ctx->evB) cudaEventDestroy(ctx->evB);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
  return 0;
}

extern "C" const char *avp_last_error(const avp_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }
extern "C" int64_t avp_launch_count(const avp_ctx *ctx) { return ctx ? ctx->launches : 0; }

extern "C" int avp_scenarios_upload(avp_ctx *ctx, int n, const double *poses, const int32_t *obs_off, const int32_t *nv,
                                    const int32_t *vert_off, const double *verts, const double *boundary_override) {
  if (!ctx) return -3;
  if (n <= 0 || !poses || !obs_off) FAIL("avp_scenarios_upload: bad arguments");
  CK(cudaSetDevice(ctx->device));
  ctx->n = 0; ctx->rasterised = false; ctx->dq_scen = -1;
  const double ds = ctx->cfg.map_discrete_size;
  ctx->h_scen.assign(n, ScenDev());
  int64_t cost_off = 0, col_off = 0, id_off = 0;
  for (int i = 0; i < n; ++i) {
    ScenDev &S = ctx->h_scen[i];
    memset(&S, 0, sizeof(S));
    memcpy(S.pose, poses + 6 * i, 6 * sizeof(double));
    if (boundary_override) memcpy(S.b, boundary_override + 4 * i, 4 * sizeof(double));
    else {                                                       // costmap.py:143-146, :169-172
      const double x0 = S.pose[0], y0 = S.pose[1], xf = S.pose[3], yf = S.pose[4];
      S.b[0] = floor((x0 < xf ? x0 : xf) - 12); S.b[1] = floor((x0 > xf ? x0 : xf) + 12);
      S.b[2] = floor((y0 < yf ? y0 : yf) - 12); S.b[3] = floor((y0 > yf ? y0 : yf) + 12);
    }
    S.nx = (int)((S.b[1] - S.b[0]) / ds); S.ny = (int)((S.b[3] - S.b[2]) / ds);           // costmap.py:182-185
    if (S.nx < 3 || S.ny < 3 || S.nx > 16384 || S.ny > 16384) FAIL("avp_scenarios_upload: map extent out of range");
    S.stepx = (S.b[1] - S.b[0]) / (S.nx - 1); S.stepy = (S.b[3] - S.b[2]) / (S.ny - 1);   // np.linspace step
    S.inv_stepx = 1.0 / S.stepx;
    volatile double x1 = 1.0 * S.stepx, y1 = 1.0 * S.stepy; x1 = x1 + S.b[0]; y1 = y1 + S.b[2];
    S.dx = x1 - S.b[0]; S.dy = y1 - S.b[2];                                                  // costmap.py:190-191
    S.stride = (int)((S.b[1] - S.b[0]) / S.dx); S.mx = S.stride; S.my = (int)((S.b[3] - S.b[2]) / S.dy);
    const long W = (long)floor((S.b[1] - S.b[0]) / S.dx) + 2, H = (long)floor((S.b[3] - S.b[2]) / S.dy) + 2;
    S.n_ids = (int32_t)(H * S.stride + W + 8);
    S.obs_begin = obs_off[i]; S.obs_end = obs_off[i + 1];
    S.cost_off = cost_off; cost_off += (int64_t)S.nx * S.ny; cost_off = (cost_off + 15) & ~15ll;
    S.col_off = col_off; col_off += (S.nx + 1 + 3) & ~3;        // 16-byte aligned: the column starts are bulk-copied to shared memory (k_plan)
    S.id_off = id_off; id_off += (S.n_ids + 3) & ~3;
  }
  {   // longest-first order for the persistent CTAs' work counter (the eager Dijkstra grows with the start-goal distance)
    std::vector<std::pair<double, int32_t>> key(n);
    for (int i = 0; i < n; ++i) { const ScenDev &S = ctx->h_scen[i]; const double ddx = S.pose[0] - S.pose[3], ddy = S.pose[1] - S.pose[4]; key[i] = {-(ddx * ddx + ddy * ddy), i}; }
    std::stable_sort(key.begin(), key.end());
    std::vector<int32_t> order(n);
    for (int i = 0; i < n; ++i) order[i] = key[i].second;
    CK(ensure_dev(&ctx->d_order, &ctx->cap_order, sizeof(int32_t) * n));
    CK(cudaMemcpy(ctx->d_order, order.data(), sizeof(int32_t) * n, cudaMemcpyHostToDevice));
  }
  ctx->cost_bytes = cost_off; ctx->col_count = col_off; ctx->id_count = id_off;      // ctx->n is set after the last copy succeeded
  const int n_poly = obs_off[n];
  const int n_vert = n_poly > 0 ? vert_off[n_poly] : 0;
  CK(ensure_dev(&ctx->d_scen, &ctx->cap_scen, sizeof(ScenDev) * n));
  CK(ensure_dev(&ctx->d_nv, &ctx->cap_nv, sizeof(int32_t) * (n_poly + 1)));
  CK(ensure_dev(&ctx->d_vert_off, &ctx->cap_vert_off, sizeof(int32_t) * (n_poly + 1)));
  CK(ensure_dev(&ctx->d_verts, &ctx->cap_verts, sizeof(double) * 2 * (n_vert + 1)));
  CK(ensure_dev(&ctx->d_cost, &ctx->cap_cost, (size_t)cost_off + 16));
  CK(ensure_dev(&ctx->d_col, &ctx->cap_col, sizeof(int32_t) * (col_off + 1)));
  if (!ctx->d_hval || (size_t)id_off > ctx->cap_ids || (size_t)id_off * 4 < ctx->cap_ids) {     // the four per-id arrays share one capacity (entries)
    free_dev(ctx->d_hval); free_dev(ctx->d_ost); free_dev(ctx->d_gx); free_dev(ctx->d_gy);
    ctx->d_hval = ctx->d_ost = nullptr; ctx->d_gx = ctx->d_gy = nullptr; ctx->cap_ids = 0;
    CK(cudaMalloc(&ctx->d_hval, sizeof(int32_t) * id_off));
    CK(cudaMalloc(&ctx->d_ost, sizeof(int32_t) * id_off));
    CK(cudaMalloc(&ctx->d_gx, sizeof(double) * id_off));
    CK(cudaMalloc(&ctx->d_gy, sizeof(double) * id_off));
    ctx->cap_ids = (size_t)id_off;
  }
  CK(cudaMemcpyAsync(ctx->d_scen, ctx->h_scen.data(), sizeof(ScenDev) * n, cudaMemcpyHostToDevice, ctx->stream));
  if (n_poly > 0) {
    CK(cudaMemcpyAsync(ctx->d_nv, nv, sizeof(int32_t) * n_poly, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_vert_off, vert_off, sizeof(int32_t) * (n_poly + 1), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_verts, verts, sizeof(double) * 2 * n_vert, cudaMemcpyHostToDevice, ctx->stream));
  }
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->n = n;
  return 0;
}

extern "C" int avp_rasterise(avp_ctx *ctx) {
  if (!ctx) return -3;
  if (ctx->n <= 0) FAIL("avp_rasterise: no scenarios uploaded");
  CK(cudaSetDevice(ctx->device));
  const int n = ctx->n;
  if (!ctx->d_cell_total) CK(cudaMalloc(&ctx->d_cell_total, 2 * sizeof(long long)));
  if (!ctx->d_cells) CK(ensure_dev(&ctx->d_cells, &ctx->cap_cells, sizeof(double2) * 1024));
  CK(cudaMemsetAsync(ctx->d_cost, 0, (size_t)ctx->cost_bytes, ctx->stream));
  k_raster<<<n, 64, 0, ctx->stream>>>(n, ctx->d_scen, ctx->d_nv, ctx->d_vert_off, ctx->d_verts, ctx->d_cost); ctx->launches++;
  k_count_cols<<<n, 128, 0, ctx->stream>>>(n, ctx->d_scen, ctx->d_cost, ctx->d_col); ctx->launches++;
  // cell offsets by a device-side scan; the fill runs against the current capacity of the (grow-only) cell list and only a batch
  // that needs more than any batch before it costs a second round
  long long tot[2] = {0, 0};
  for (int round = 0; round < 2; ++round) {
    k_scan_cells<<<1, 1024, 0, ctx->stream>>>(n, ctx->d_scen, (long long)(ctx->cap_cells / sizeof(double2)) - 1, ctx->d_cell_total); ctx->launches++;
    k_fill_cells<<<n, 128, 0, ctx->stream>>>(n, ctx->d_scen, ctx->d_cost, ctx->d_col, ctx->d_cells, ctx->d_cell_total); ctx->launches++;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(tot, ctx->d_cell_total, sizeof(tot), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(ctx->h_scen.data(), ctx->d_scen, sizeof(ScenDev) * n, cudaMemcpyDeviceToHost, ctx->stream));      // n_obs, raster_error, offsets for the host-side queries
    CK(cudaStreamSynchronize(ctx->stream));
    if (!tot[1]) break;
    if (round == 1) FAIL("avp_rasterise: the obstacle cell list could not be sized");
    free_dev(ctx->d_cells); ctx->d_cells = nullptr; ctx->cap_cells = 0;
    CK(ensure_dev(&ctx->d_cells, &ctx->cap_cells, sizeof(double2) * (size_t)(tot[0] + tot[0] / 8 + 1024)));
  }
  ctx->cell_count = tot[0];
  ctx->rasterised = true;
  return 0;
}

extern "C" int avp_fetch_map(avp_ctx *ctx, int s, int32_t *dims, double *geom, uint8_t *cost_map, int64_t cap) {
  if (!ctx) return -3;
  if (s < 0 || s >= ctx->n) FAIL("avp_fetch_map: scenario index out of range");
  if (!ctx->rasterised) FAIL("avp_fetch_map: call avp_rasterise first");
  CK(cudaSetDevice(ctx->device));
  const ScenDev &S = ctx->h_scen[s];
  if (dims) { dims[0] = S.nx; dims[1] = S.ny; dims[2] = S.n_obs; dims[3] = S.raster_error; }
  if (geom) { memcpy(geom, S.b, 4 * sizeof(double)); geom[4] = S.dx; geom[5] = S.dy; }
  if (cost_map) {
    if (cap < (int64_t)S.nx * S.ny) FAIL("avp_fetch_map: buffer too small");
    CK(cudaMemcpyAsync(cost_map, ctx->d_cost + S.cost_off, (size_t)S.nx * S.ny, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
  }
  return 0;
}

extern "C" int avp_collision_check(avp_ctx *ctx, int s, int m, const double *poses, uint8_t *out) {
  if (!ctx) return -3;
  if (s < 0 || s >= ctx->n || m < 0) FAIL("avp_collision_check: bad arguments");
  if (!ctx->rasterised) FAIL("avp_collision_check: call avp_rasterise first");
  if (m == 0) return 0;
  CK(cudaSetDevice(ctx->device));
  const size_t pb = sizeof(double) * 3 * m;
  if (ensure_scratch(ctx, pb + m + 64)) return -1;
  double *d_p = (double *)ctx->d_scratch; uint8_t *d_o = (uint8_t *)ctx->d_scratch + pb;
  CK(cudaMemcpyAsync(d_p, poses, pb, cudaMemcpyHostToDevice, ctx->stream));
  const int wpb = 4, blocks = (m + wpb - 1) / wpb;
  k_check_batch<<<blocks, wpb * 32, 0, ctx->stream>>>(ctx->cfg, ctx->d_scen, s, ctx->d_cells, ctx->d_col, m, d_p, d_o); ctx->launches++;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(out, d_o, m, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return 0;
}

extern "C" int avp_check_start_goal(avp_ctx *ctx, uint8_t *out2n) {
  if (!ctx) return -3;
  if (ctx->n <= 0 || !out2n) FAIL("avp_check_start_goal: bad arguments");
  if (!ctx->rasterised) FAIL("avp_check_start_goal: call avp_rasterise first");
  CK(cudaSetDevice(ctx->device));
  const int m = 2 * ctx->n;
  if (ensure_scratch(ctx, (size_t)m + 64)) return -1;
  uint8_t *d_o = (uint8_t *)ctx->d_scratch;
  k_check_start_goal<<<(m + 3) / 4, 128, 0, ctx->stream>>>(ctx->cfg, ctx->d_scen, ctx->n, ctx->d_cells, ctx->d_col, d_o); ctx->launches++;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(out2n, d_o, m, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return 0;
}

extern "C" int avp_corridor(avp_ctx *ctx, int s, int m, const double *poses, double expand_dis, double *out4, int32_t *status) {
  if (!ctx) return -3;
  if (s < 0 || s >= ctx->n || m < 0 || !poses || !out4 || !status) FAIL("avp_corridor: bad arguments");
  if (!ctx->rasterised) FAIL("avp_corridor: call avp_rasterise first");
  if (m == 0) return 0;
  CK(cudaSetDevice(ctx->device));
  const size_t pb = sizeof(double) * 3 * m, ob = sizeof(double) * 4 * m;
  if (ensure_scratch(ctx, pb + ob + sizeof(int32_t) * m + 64)) return -1;
  double *d_p = (double *)ctx->d_scratch, *d_o = d_p + 3 * (size_t)m; int32_t *d_s = (int32_t *)(d_o + 4 * (size_t)m);
  CK(cudaMemcpyAsync(d_p, poses, pb, cudaMemcpyHostToDevice, ctx->stream));
  const int wpb = 4, blocks = (m + wpb - 1) / wpb;
  k_corridor<<<blocks, wpb * 32, 0, ctx->stream>>>(ctx->cfg, ctx->d_scen, s, ctx->d_cells, ctx->d_col, expand_dis, m, d_p, d_o, d_s); ctx->launches++;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(out4, d_o, ob, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(status, d_s, sizeof(int32_t) * m, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return 0;
}

extern "C" int avp_expand_pure(avp_ctx *ctx, int s, const double parent_pose[3], double *out_pose, int32_t *out_flags, double *out_rsL) {
  if (!ctx) return -3;
  if (s < 0 || s >= ctx->n) FAIL("avp_expand_pure: scenario index out of range");
  if (!ctx->rasterised) FAIL("avp_expand_pure: call avp_rasterise first");
  CK(cudaSetDevice(ctx->device));
  const int nc = 2 * ctx->cfg.steering_angle_num;
  if (ensure_scratch(ctx, sizeof(double) * 4 * nc + sizeof(int32_t) * nc + 64)) return -1;
  double *d_pose = (double *)ctx->d_scratch, *d_L = d_pose + 3 * nc; int32_t *d_fl = (int32_t *)(d_L + nc);
  k_expand_pure<<<1, nc * 32, 0, ctx->stream>>>(ctx->cfg, ctx->d_scen, s, ctx->d_cells, ctx->d_col, parent_pose[0], parent_pose[1], parent_pose[2], d_pose, d_fl, d_L);
  ctx->launches++;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(out_pose, d_pose, sizeof(double) * 3 * nc, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(out_rsL, d_L, sizeof(double) * nc, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(out_flags, d_fl, sizeof(int32_t) * nc, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return 0;
}

extern "C" int avp_rs_optimal(avp_ctx *ctx, int m, const double *q, double maxc, double step_size, int xy_np, int phi_np,
                              double *lengths, char *ctypes, int32_t *nseg, double *L, int cap_pts,
                              double *x, double *y, double *yaw, int32_t *dir, int32_t *n_pts) {
  if (!ctx) return -3;
  if (m <= 0 || cap_pts <= 0) FAIL("avp_rs_optimal: bad arguments");
  CK(cudaSetDevice(ctx->device));
  size_t off = 0; auto take = [&](size_t b) { size_t o = off; off += (b + 15) & ~(size_t)15; return o; };
  const size_t o_q = take(sizeof(double) * 6 * m), o_len = take(sizeof(double) * 5 * m), o_ct = take(8 * (size_t)m), o_ns = take(4 * (size_t)m),
               o_L = take(8 * (size_t)m), o_x = take(8 * (size_t)m * cap_pts), o_y = take(8 * (size_t)m * cap_pts), o_yaw = take(8 * (size_t)m * cap_pts),
               o_dir = take(4 * (size_t)m * cap_pts), o_np = take(4 * (size_t)m);
  if (ensure_scratch(ctx, off + 64)) return -1;
  char *B = (char *)ctx->d_scratch;
  CK(cudaMemcpyAsync(B + o_q, q, sizeof(double) * 6 * m, cudaMemcpyHostToDevice, ctx->stream));
  k_rs_optimal<<<(m + 3) / 4, 128, 0, ctx->stream>>>(m, (double *)(B + o_q), maxc, step_size, xy_np, phi_np, (double *)(B + o_len), B + o_ct, (int32_t *)(B + o_ns),
                                                      (double *)(B + o_L), cap_pts, (double *)(B + o_x), (double *)(B + o_y), (double *)(B + o_yaw),
                                                      (int32_t *)(B + o_dir), (int32_t *)(B + o_np));
  ctx->launches++;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(lengths, B + o_len, sizeof(double) * 5 * m, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(ctypes, B + o_ct, 8 * (size_t)m, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(nseg, B + o_ns, 4 * (size_t)m, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(L, B + o_L, 8 * (size_t)m, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(x, B + o_x, 8 * (size_t)m * cap_pts, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(y, B + o_y, 8 * (size_t)m * cap_pts, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(yaw, B + o_yaw, 8 * (size_t)m * cap_pts, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(dir, B + o_dir, 4 * (size_t)m * cap_pts, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(n_pts, B + o_np, 4 * (size_t)m, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return 0;
}

/* Workspaces of a launch over the uploaded batch.  Slots: a search owns one from its first pop to its last (also while it
 * waits in the run queue), so as many as there are scenarios is the comfortable number; with less memory than that the
 * kernel sends fresh scenarios to the back of the queue while every slot is taken. */
static int ensure_ws(avp_ctx *ctx, int ctas) {
  const int nchild = 2 * ctx->cfg.steering_angle_num;
  const int max_pops = ctx->cfg.max_pops > 0 ? ctx->cfg.max_pops : 20000;
  const int node_cap = (nchild * (max_pops + 1) + 2 + 7) & ~7;       // multiple of 8: every slot's 16-byte heap records start on a 128-byte line
  const int n = ctx->n;
  if (ctx->ws_n >= n && ctx->ws_n <= 4 * n + 64 && ctx->ws_ctas >= ctas && ctx->node_cap == node_cap) return 0;      // grow-only with slack
  free_ws(ctx);
  int hb = 1; while (hb < 2 * node_cap) hb <<= 1;
  ctx->node_cap = node_cap; ctx->htab_size = hb; ctx->dheap_cap = 1 << 16;
  CK(cudaMalloc(&ctx->d_dheap, sizeof(unsigned long long) * (size_t)n * ctx->dheap_cap));
  CK(cudaMalloc(&ctx->d_state, sizeof(ScenState) * (size_t)n));
  CK(cudaMalloc(&ctx->d_ctl, sizeof(PlanCtl)));
  int ring = 2; while (ring < n + 1) ring <<= 1;
  CK(cudaMalloc(&ctx->d_queue, sizeof(int32_t) * ring)); ctx->q_mask = ring - 1;
  CK(cudaMalloc(&ctx->d_course, sizeof(double) * (size_t)ctas * 3 * AVP_COURSE_CAP));
  CK(cudaMalloc(&ctx->d_course_dir, sizeof(int32_t) * (size_t)ctas * AVP_COURSE_CAP));
  CK(cudaMalloc(&ctx->d_cand, (size_t)ctas * AVP_CAND_SMEM));
  const size_t slot_bytes = (sizeof(Node) + sizeof(NodeShot) + sizeof(OEnt)) * (size_t)node_cap + sizeof(int32_t) * (size_t)hb;
  size_t free_b = 0, total_b = 0;
  CK(cudaMemGetInfo(&free_b, &total_b));
  size_t fit = (size_t)((double)free_b * 0.8) / slot_bytes;
  { const char *se = getenv("AVP_SLOTS"); if (se && atoi(se) > 0) fit = (size_t)atoi(se); }      // development aid: force the slot pool to run dry
  int slots = n; if ((size_t)slots > fit) slots = (int)fit;
  if (slots < 1) FAIL("plan: not enough device memory for the search workspaces");
  CK(cudaMalloc(&ctx->d_nodes, sizeof(Node) * (size_t)slots * node_cap));
  CK(cudaMalloc(&ctx->d_nshot, sizeof(NodeShot) * (size_t)slots * node_cap));
  CK(cudaMalloc(&ctx->d_oheap, sizeof(OEnt) * (size_t)slots * node_cap));
  CK(cudaMalloc(&ctx->d_htab, sizeof(int32_t) * (size_t)slots * hb));
  int sring = 2; while (sring < slots + 1) sring <<= 1;
  CK(cudaMalloc(&ctx->d_slot_ring, sizeof(int32_t) * sring)); ctx->slot_mask = sring - 1;
  ctx->ws_n = n; ctx->ws_slots = slots; ctx->ws_ctas = ctas;
  return 0;
}

static int ensure_results(avp_ctx *ctx, int cap_path, int cap_pops) {
  if (ctx->res_n == ctx->n && ctx->cap_path == cap_path && ctx->cap_pops == cap_pops) return 0;
  free_results(ctx);
  const size_t n = ctx->n;
  CK(cudaMalloc(&ctx->d_sums, sizeof(avp_plan_summary) * n));
  CK(cudaMalloc(&ctx->d_paths, sizeof(double) * n * cap_path * 3));
  CK(cudaMalloc(&ctx->d_pops, sizeof(int32_t) * n * (size_t)(cap_pops > 0 ? cap_pops : 1)));
  CK(cudaMalloc(&ctx->d_hq, sizeof(int32_t) * n * AVP_HQ_CAP * 3));
  CK(cudaMalloc(&ctx->d_dbg, sizeof(int) * n * 8));
  CK(cudaMalloc(&ctx->d_prof, sizeof(long long) * n * 16));
  CK(cudaMemset(ctx->d_prof, 0, sizeof(long long) * n * 16));
  CK(cudaMalloc(&ctx->d_wprof, sizeof(long long) * n * 384));
  CK(cudaMemset(ctx->d_wprof, 0, sizeof(long long) * n * 384));
  CK(cudaMemset(ctx->d_dbg, 0, sizeof(int) * n * 8));
  ctx->res_n = ctx->n; ctx->cap_path = cap_path; ctx->cap_pops = cap_pops;
  return 0;
}

static int wait_search(avp_ctx *ctx, cudaEvent_t ev) {
  const char *lim = getenv("AVP_HOST_TIMEOUT_S");        // development aid: never wait forever on a kernel
  if (lim && atof(lim) > 0) {
    const double limit = atof(lim);
    timespec t0, t1; clock_gettime(CLOCK_MONOTONIC, &t0);
    for (;;) {
      cudaError_t q = cudaEventQuery(ev);
      if (q == cudaSuccess) break;
      if (q != cudaErrorNotReady) { ctx->err = std::string("k_search: ") + cudaGetErrorString(q); return -1; }
      clock_gettime(CLOCK_MONOTONIC, &t1);
      if ((t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec) > limit) {
        std::vector<int> dbg((size_t)ctx->n * 8, 0);
        cudaMemcpy(dbg.data(), ctx->d_dbg, sizeof(int) * dbg.size(), cudaMemcpyDeviceToHost);
        std::string msg = "k_search timed out; unfinished scenarios (id: phase,pops,h_closed,open):";
        int shown = 0;
        for (int i = 0; i < ctx->n && shown < 32; ++i) if (dbg[8 * i] != 9) { char b[96]; snprintf(b, sizeof b, " [%d: %d,%d,%d,%d]", i, dbg[8 * i], dbg[8 * i + 1], dbg[8 * i + 2], dbg[8 * i + 3]); msg += b; ++shown; }
        ctx->err = msg;
        return -11;
      }
      timespec ts = {0, 1000000}; nanosleep(&ts, nullptr);
    }
  }
  CK(cudaStreamSynchronize(ctx->stream));
  return 0;
}

/* One pass: queue / slot pool initialisation, the eager Dijkstra of every scenario (one warp each), then ONE persistent search
 * kernel (avp_plan.cuh).  No host round trip in between; results are a deterministic function of the scenarios whatever the
 * scheduling (quantum, CTA width, SM-pair placement). */
static int launch_search(avp_ctx *ctx, float *elapsed_ms) {
  if (ctx->n <= 0) FAIL("plan: no scenarios uploaded");
  ctx->dq_scen = -1;
  if (!ctx->rasterised) FAIL("plan: call avp_rasterise first");
  CK(cudaSetDevice(ctx->device));
  // CTA width: 512 threads, one CTA per SM (15 evaluator warps per pop: the shortest pop), or 256 threads, two CTAs per SM
  // (more searches in flight per SM: throughput for batches with far more long searches than SMs)
  int block = 512;
  { const char *be = getenv("AVP_PLAN_BLOCK"); if (be && (atoi(be) == 256 || atoi(be) == 640)) block = atoi(be); }
  const int per_sm = block == 256 ? ctx->ctas_per_sm[1] : 1;
  int grid = ctx->n_sm * per_sm; if (grid > ctx->n) grid = ctx->n;
  // Two launches for batches with more scenarios than a wave of wide CTAs: most searches are short (C2: median 19 pops) and their time
  // is the single-warp Dijkstra resumes, during which a 512-thread CTA idles.  A first launch of NARROW CTAs (128 threads, three per
  // SM) gives every fresh scenario a few pops -- the short ones finish there, twelve Dijkstra warps per SM instead of one --, every
  // unfinished search goes to the run queue, and the wide launch resumes them (stream order: no host round trip, nothing re-planned).
  bool two_phase = ctx->n > 2 * ctx->n_sm;
  { const char *te = getenv("AVP_TWO_PHASE"); if (te) two_phase = atoi(te) != 0; }
  // the narrow CTAs: 64 threads (commit warp + one evaluator warp, eight CTAs per SM: as many Dijkstra warps) or 128 threads (three per SM)
  int nblock = 64;
  { const char *ne = getenv("AVP_NARROW_BLOCK"); if (ne && (atoi(ne) == 64 || atoi(ne) == 128)) nblock = atoi(ne); }
  int grid_n = ctx->n_sm * (nblock == 64 ? ctx->narrow64_per_sm : ctx->narrow_per_sm); if (grid_n > ctx->n) grid_n = ctx->n;
  {
    int need = ctx->n_sm * ctx->ctas_per_sm[1] > grid ? ctx->n_sm * ctx->ctas_per_sm[1] : grid;
    if (grid_n > need) need = grid_n;
    if (ensure_ws(ctx, need)) return -1;
  }
  PlanParams PP; memset(&PP, 0, sizeof(PP));
  KParams &P = PP.K;
  P.cfg = ctx->cfg; if (P.cfg.max_pops <= 0) P.cfg.max_pops = 20000;
  P.n_scen = ctx->n; P.scen = ctx->d_scen; P.cost = ctx->d_cost; P.cells = ctx->d_cells; P.col_start = ctx->d_col;
  P.hval = ctx->d_hval; P.ost = ctx->d_ost; P.gx = ctx->d_gx; P.gy = ctx->d_gy;
  P.dheap = ctx->d_dheap; P.dheap_cap = ctx->dheap_cap; P.nodes = ctx->d_nodes; P.node_cap = ctx->node_cap; P.oheap = ctx->d_oheap; P.nshot = ctx->d_nshot;
  P.htab = ctx->d_htab; P.htab_size = ctx->htab_size; P.htab_stride = ctx->htab_size; P.course = ctx->d_course; P.course_dir = ctx->d_course_dir; P.cand_scratch = ctx->d_cand;
  P.sums = ctx->d_sums; P.paths = ctx->d_paths; P.cap_path = ctx->cap_path; P.pops = ctx->cap_pops > 0 ? ctx->d_pops : nullptr; P.cap_pops = ctx->cap_pops;
  P.pop_fgh = nullptr;
  if (ctx->trace_fgh && ctx->cap_pops > 0) {
    CK(ensure_dev(&ctx->d_pop_fgh, &ctx->cap_fgh, sizeof(double) * 3 * (size_t)ctx->n * ctx->cap_pops));
    P.pop_fgh = ctx->d_pop_fgh;
  }
  P.hq_log = ctx->d_hq; P.work_counter = ctx->d_counter; P.dbg = getenv("AVP_HOST_TIMEOUT_S") ? ctx->d_dbg : nullptr; P.prof = ctx->d_prof; P.wprof = ctx->d_wprof;
  { const char *tpp = getenv("AVP_TRACE_POP"); P.trace_pop = tpp ? atoi(tpp) : -1; } P.watchdog_cycles = ctx->watchdog_cycles;
  P.work_list = ctx->d_order; P.n_work = ctx->n;
  PP.state = ctx->d_state; PP.ctl = ctx->d_ctl; PP.queue = ctx->d_queue; PP.q_mask = ctx->q_mask;
  PP.slot_ring = ctx->d_slot_ring; PP.slot_mask = ctx->slot_mask; PP.n_slots = ctx->ws_slots;
  { const char *qe = getenv("AVP_QUANTUM"); PP.quantum = (qe && atoi(qe) > 0) ? atoi(qe) : 512; }
  { const char *oe = getenv("AVP_OVERFLOW_ODD"); PP.overflow_odd = (oe && atoi(oe) == 0) ? 0 : 1; }
  { const char *fe = getenv("AVP_FORCE_YIELD"); PP.force_yield = (fe && atoi(fe) != 0) ? 1 : 0; }
  PP.phase = two_phase ? 1 : 0; PP.cell_smem = AVP_CELL_SMEM;
  // SM pairs: only when the grid covers every SM (else the hardware's placement decides) and not switched off (AVP_SPREAD=0)
  { const char *se = getenv("AVP_SPREAD"); P.spread = (grid == ctx->n_sm * per_sm && ctx->n_sm >= 4 && !(se && atoi(se) == 0)) ? 1 : 0; }
  { const char *me = getenv("AVP_SPREAD_MAX"); PP.spread_max = me ? atoi(me) : (ctx->n_sm / 2 + ctx->n_sm / 4) * per_sm; }
#if defined(AVP_PROFILE) || defined(AVP_PROFILE_LIGHT)
  CK(cudaMemsetAsync(ctx->d_prof, 0, sizeof(long long) * (size_t)ctx->n * 16, ctx->stream));
  CK(cudaMemsetAsync(ctx->d_wprof, 0, sizeof(long long) * (size_t)ctx->n * 384, ctx->stream));
#endif
  const int ring_max = (ctx->q_mask > ctx->slot_mask ? ctx->q_mask : ctx->slot_mask) + 1;
  CK(cudaEventRecord(ctx->ev0, ctx->stream));
  k_plan_init<<<(ring_max + 255) / 256, 256, 0, ctx->stream>>>(PP); ctx->launches++;
  {
    int dgrid = (ctx->n + AVP_DIJ_WARPS - 1) / AVP_DIJ_WARPS; const int dmax = ctx->n_sm * 5; if (dgrid > dmax) dgrid = dmax;
    k_dij_eager<<<dgrid, AVP_DIJ_WARPS * 32, 0, ctx->stream>>>(PP); ctx->launches++;
  }
  CK(cudaEventRecord(ctx->evM, ctx->stream));
  if (two_phase) {
    PlanParams P1 = PP;
    P1.phase = 1; P1.cell_smem = 0; P1.K.spread = 0;
    { const char *be = getenv("AVP_NARROW_BUDGET"); P1.quantum = (be && atoi(be) > 0) ? atoi(be) : 128; }
    if (nblock == 64) k_plan<64><<<grid_n, 64, AVP_PLAN_DYN_SMEM(64, 0), ctx->stream>>>(P1);
    else k_plan<128><<<grid_n, 128, AVP_PLAN_DYN_SMEM(128, 0), ctx->stream>>>(P1);
    ctx->launches++;
    PP.phase = 2;
  }
  CK(cudaEventRecord(ctx->evN, ctx->stream));
  if (block == 512) k_plan<512><<<grid, 512, AVP_PLAN_DYN_SMEM(512, AVP_CELL_SMEM), ctx->stream>>>(PP);
  else if (block == 640) k_plan<640><<<grid, 640, AVP_PLAN_DYN_SMEM(640, AVP_CELL_SMEM), ctx->stream>>>(PP);
  else k_plan<256><<<grid, 256, AVP_PLAN_DYN_SMEM(256, AVP_CELL_SMEM), ctx->stream>>>(PP);
  ctx->launches++;
  CK(cudaEventRecord(ctx->ev1, ctx->stream));
  CK(cudaGetLastError());
  ctx->plan_block = block; ctx->plan_grid = grid;
  if (wait_search(ctx, ctx->ev1)) return -1;
  float ms1 = 0.f, ms2 = 0.f;
  CK(cudaEventElapsedTime(&ms1, ctx->ev0, ctx->evM));
  CK(cudaEventElapsedTime(&ms2, ctx->evM, ctx->ev1));
  { float msn = 0.f; CK(cudaEventElapsedTime(&msn, ctx->evM, ctx->evN)); ctx->narrow_ms = two_phase ? msn : 0.f; }
  ctx->pass_ms[0] = ms1; ctx->pass_ms[1] = ms2;
  if (elapsed_ms) *elapsed_ms = ms1 + ms2;
  PlanCtl c;
  CK(cudaMemcpyAsync(&c, ctx->d_ctl, sizeof(c), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->n_suspends = c.n_suspends;
  if (c.finalised != ctx->n || c.error) { char b[160]; snprintf(b, sizeof b, "plan: the search kernel ended with %d of %d scenarios finished (queue error %d)", c.finalised, ctx->n, c.error); ctx->err = b; return -12; }
  return 0;
}

extern "C" int avp_fetch_results(avp_ctx *ctx, avp_plan_summary *summaries, double *final_path, int cap_path, int32_t *pops, int cap_pops) {
  if (!ctx) return -3;
  if (ctx->res_n != ctx->n || !ctx->d_sums) FAIL("avp_fetch_results: no results");
  if (cap_path != ctx->cap_path || (pops && cap_pops != ctx->cap_pops)) FAIL("avp_fetch_results: capacities differ from the plan call");
  CK(cudaSetDevice(ctx->device));
  const size_t n = ctx->n;
  if (summaries) CK(cudaMemcpyAsync(summaries, ctx->d_sums, sizeof(avp_plan_summary) * n, cudaMemcpyDeviceToHost, ctx->stream));
  if (final_path) CK(cudaMemcpyAsync(final_path, ctx->d_paths, sizeof(double) * n * cap_path * 3, cudaMemcpyDeviceToHost, ctx->stream));
  if (pops && cap_pops > 0) CK(cudaMemcpyAsync(pops, ctx->d_pops, sizeof(int32_t) * n * cap_pops, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return 0;
}

/* f, g, h of every popped node (hybrid_a_star.py:206-216, :224-230 at the time of open_list.get()), recorded beside the pop
 * indices of the following plans when on != 0 and cap_pops > 0; read with avp_fetch_pop_fgh */
extern "C" int avp_trace_fgh(avp_ctx *ctx, int on) { if (!ctx) return -3; ctx->trace_fgh = on != 0; return 0; }
extern "C" int avp_fetch_pop_fgh(avp_ctx *ctx, double *out, int cap_pops) {
  if (!ctx) return -3;
  if (!ctx->d_pop_fgh || !ctx->trace_fgh || cap_pops != ctx->cap_pops || !out) FAIL("avp_fetch_pop_fgh: no f/g/h trace of that size (avp_trace_fgh, then plan with cap_pops > 0)");
  CK(cudaSetDevice(ctx->device));
  CK(cudaMemcpyAsync(out, ctx->d_pop_fgh, sizeof(double) * 3 * (size_t)ctx->n * cap_pops, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return 0;
}

extern "C" int avp_plan_configure(avp_ctx *ctx, int cap_path, int cap_pops) {
  if (!ctx) return -3;
  if (cap_path <= 0 || cap_pops < 0) FAIL("avp_plan_configure: bad capacities");
  CK(cudaSetDevice(ctx->device));
  return ensure_results(ctx, cap_path, cap_pops);
}

extern "C" int avp_plan_batch(avp_ctx *ctx, avp_plan_summary *summaries, double *final_path, int cap_path, int32_t *pops, int cap_pops) {
  if (!ctx) return -3;
  if (cap_path <= 0) FAIL("avp_plan_batch: cap_path must be positive");
  CK(cudaSetDevice(ctx->device));
  if (ensure_results(ctx, cap_path, pops ? cap_pops : 0)) return -1;
  int rc = launch_search(ctx, nullptr);
  if (rc) return rc;
  return avp_fetch_results(ctx, summaries, final_path, cap_path, pops, pops ? cap_pops : 0);
}

extern "C" int avp_plan_batch_resident(avp_ctx *ctx, float *elapsed_ms) {
  if (!ctx) return -3;
  if (ctx->res_n != ctx->n || !ctx->d_sums) { if (ensure_results(ctx, ctx->cap_path > 0 ? ctx->cap_path : 512, ctx->cap_pops)) return -1; }
  return launch_search(ctx, elapsed_ms);
}

extern "C" int avp_result_device_buffer(avp_ctx *ctx, void **sums, void **paths, int64_t *n, int64_t *cap_path) {
  if (!ctx) return -3;
  if (ctx->res_n != ctx->n || !ctx->d_sums) FAIL("avp_result_device_buffer: no results");
  if (sums) *sums = ctx->d_sums; if (paths) *paths = ctx->d_paths; if (n) *n = ctx->n; if (cap_path) *cap_path = ctx->cap_path;
  return 0;
}

extern "C" int avp_split_paths(avp_ctx *ctx, int cap_pts, int cap_seg, double *split_pts, int32_t *seg_len, int32_t *info) {
  if (!ctx) return -3;
  if (ctx->res_n != ctx->n || !ctx->d_sums) FAIL("avp_split_paths: no plan results on the device (call avp_plan_batch first)");
  if (cap_pts <= 0 || cap_seg <= 0 || !split_pts || !seg_len || !info) FAIL("avp_split_paths: bad arguments");
  CK(cudaSetDevice(ctx->device));
  const size_t n = ctx->n;
  const size_t b_pts = sizeof(double) * n * cap_pts * 3, b_seg = sizeof(int32_t) * n * cap_seg, b_inf = sizeof(int32_t) * n * 4;
  if (ensure_scratch(ctx, b_pts + b_seg + b_inf + 64)) return -1;
  double *d_pts = (double *)ctx->d_scratch; int32_t *d_seg = (int32_t *)((char *)ctx->d_scratch + b_pts), *d_inf = (int32_t *)((char *)ctx->d_scratch + b_pts + b_seg);
  CK(cudaMemsetAsync(d_seg, 0, b_seg, ctx->stream));
  k_split_batch<<<(unsigned)((n + 3) / 4), 128, 0, ctx->stream>>>(ctx->cfg, ctx->d_scen, (int)n, ctx->d_cells, ctx->d_col, ctx->d_sums, ctx->d_paths, ctx->cap_path,
                                                               d_pts, cap_pts, d_seg, cap_seg, d_inf); ctx->launches++;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(split_pts, d_pts, b_pts, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(seg_len, d_seg, b_seg, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(info, d_inf, b_inf, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return 0;
}

extern "C" int avp_split_path(avp_ctx *ctx, int s, int n_pts, const double *path, int cap_pts, int cap_seg, double *split_pts, int32_t *seg_len, int32_t *info4) {
  if (!ctx) return -3;
  if (s < 0 || s >= ctx->n || n_pts < 0 || cap_pts <= 0 || cap_seg <= 0 || !split_pts || !seg_len || !info4 || (n_pts > 0 && !path)) FAIL("avp_split_path: bad arguments");
  if (!ctx->rasterised) FAIL("avp_split_path: call avp_rasterise first");
  CK(cudaSetDevice(ctx->device));
  const size_t b_in = sizeof(double) * 3 * (size_t)(n_pts > 0 ? n_pts : 1), b_pts = sizeof(double) * 3 * (size_t)cap_pts, b_seg = sizeof(int32_t) * (size_t)cap_seg;
  if (ensure_scratch(ctx, b_in + b_pts + b_seg + 16 + 64)) return -1;
  double *d_in = (double *)ctx->d_scratch, *d_pts = (double *)((char *)ctx->d_scratch + b_in);
  int32_t *d_seg = (int32_t *)((char *)ctx->d_scratch + b_in + b_pts), *d_inf = (int32_t *)((char *)ctx->d_scratch + b_in + b_pts + b_seg);
  if (n_pts > 0) CK(cudaMemcpyAsync(d_in, path, sizeof(double) * 3 * (size_t)n_pts, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemsetAsync(d_seg, 0, b_seg, ctx->stream));
  k_split_one<<<1, 32, 0, ctx->stream>>>(ctx->cfg, ctx->d_scen, s, ctx->d_cells, ctx->d_col, d_in, n_pts, d_pts, cap_pts, d_seg, cap_seg, d_inf); ctx->launches++;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(split_pts, d_pts, b_pts, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(seg_len, d_seg, b_seg, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(info4, d_inf, sizeof(int32_t) * 4, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return 0;
}

extern "C" int avp_fetch_hvalues(avp_ctx *ctx, int s, int32_t *hval, int64_t cap, int64_t *n_ids) {
  if (!ctx) return -3;
  if (s < 0 || s >= ctx->n) FAIL("avp_fetch_hvalues: scenario index out of range");
  CK(cudaSetDevice(ctx->device));
  const ScenDev &S = ctx->h_scen[s];
  if (n_ids) *n_ids = S.n_ids;
  if (hval) {
    if (cap < S.n_ids) FAIL("avp_fetch_hvalues: buffer too small");
    CK(cudaMemcpyAsync(hval, ctx->d_hval + S.id_off, sizeof(int32_t) * S.n_ids, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
  }
  return 0;
}

extern "C" int avp_fetch_hq_log(avp_ctx *ctx, int s, int32_t *log3, int cap_entries) {
  if (!ctx) return -3;
  if (s < 0 || s >= ctx->n || !ctx->d_hq) FAIL("avp_fetch_hq_log: bad arguments");
  CK(cudaSetDevice(ctx->device));
  const int m = cap_entries < AVP_HQ_CAP ? cap_entries : AVP_HQ_CAP;
  CK(cudaMemcpyAsync(log3, ctx->d_hq + (size_t)s * AVP_HQ_CAP * 3, sizeof(int32_t) * 3 * m, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return 0;
}

extern "C" int avp_device_info(avp_ctx *ctx, int32_t *n_sm, int32_t *slots, int32_t *block) {
  if (!ctx) return -3;
  if (n_sm) *n_sm = ctx->n_sm; if (slots) *slots = ctx->slots; if (block) *block = ctx->plan_block;
  return 0;
}

/* development aids: per-scenario progress checkpoints and an in-kernel watchdog (SM clock cycles) */
extern "C" int avp_set_watchdog(avp_ctx *ctx, long long cycles) { if (!ctx) return -3; ctx->watchdog_cycles = cycles; return 0; }
extern "C" int avp_fetch_debug(avp_ctx *ctx, int32_t *out8n) {
  if (!ctx) return -3;
  if (!ctx->d_dbg) FAIL("avp_fetch_debug: no results");
  CK(cudaSetDevice(ctx->device));
  CK(cudaMemcpy(out8n, ctx->d_dbg, sizeof(int) * (size_t)ctx->n * 8, cudaMemcpyDeviceToHost));
  return 0;
}

/* CUDA-event bracket on the context's stream, so callers can time any sequence of entry points
 * (bench.py: resident step = rasterise + search; e2e step = upload + rasterise + search + fetch) */
extern "C" int avp_timer_start(avp_ctx *ctx) {
  if (!ctx) return -3;
  CK(cudaSetDevice(ctx->device));
  if (!ctx->evA) { CK(cudaEventCreate(&ctx->evA)); CK(cudaEventCreate(&ctx->evB)); }
  CK(cudaStreamSynchronize(ctx->stream));
  CK(cudaEventRecord(ctx->evA, ctx->stream));
  return 0;
}
extern "C" int avp_timer_stop(avp_ctx *ctx, float *elapsed_ms) {
  if (!ctx) return -3;
  if (!ctx->evA) FAIL("avp_timer_stop: timer not started");
  CK(cudaSetDevice(ctx->device));
  CK(cudaEventRecord(ctx->evB, ctx->stream));
  CK(cudaEventSynchronize(ctx->evB));
  if (elapsed_ms) CK(cudaEventElapsedTime(elapsed_ms, ctx->evA, ctx->evB));
  return 0;
}
/* CUDA-event duration of the most recent search kernel launch */
extern "C" int avp_last_search_ms(avp_ctx *ctx, float *elapsed_ms) {
  if (!ctx) return -3;
  if (elapsed_ms) *elapsed_ms = ctx->pass_ms[0] + ctx->pass_ms[1];
  return 0;
}
/* CUDA-event times of the last search: eager Dijkstra kernel, search kernel; n_info = number of times a search let go of its SM
 * (round-robin suspensions) + 100000 * CTA width */
extern "C" int avp_last_search_passes(avp_ctx *ctx, float *ms_dijkstra, float *ms_search, int32_t *n_info) {
  if (!ctx) return -3;
  if (ms_dijkstra) *ms_dijkstra = ctx->pass_ms[0]; if (ms_search) *ms_search = ctx->pass_ms[1];
  if (n_info) *n_info = (ctx->n_suspends % 100000) + 100000 * ctx->plan_block;
  return 0;
}

/* CUDA-event time of the narrow first search launch of the last plan (0: the batch was planned by one launch) */
extern "C" int avp_last_narrow_ms(avp_ctx *ctx, float *ms) { if (!ctx) return -3; if (ms) *ms = ctx->narrow_ms; return 0; }

/* per-scenario SM-cycle accumulators of the search kernel's phases (thread 0 of the CTA):
 * [0] init + eager Dijkstra, [1] loop top, [2] heappop + poses/queries, [3] lookups + collision checks +
 * rs instances, [4] selection + course plan, [5] course + shot check, [6] commit, [7] commit preparation,
 * [8] cycles in open-heap pushes, [9] pushes, [10] cycles in Dijkstra resumes during commits, [11] resumes,
 * [12] sum of final heap positions of pushed nodes, [13] cycles in heappop, [14] sum of heap sizes at pops */
extern "C" int avp_fetch_warp_profile(avp_ctx *ctx, int64_t *out128n) {
  if (!ctx || !out128n) return -3;
  if (!ctx->d_wprof) FAIL("avp_fetch_warp_profile: no results");
  CK(cudaSetDevice(ctx->device));
  CK(cudaMemcpy(out128n, ctx->d_wprof, sizeof(long long) * (size_t)ctx->n * 384, cudaMemcpyDeviceToHost));
  return 0;
}

extern "C" int avp_fetch_profile(avp_ctx *ctx, int64_t *out8n) {
  if (!ctx) return -3;
  if (!ctx->d_prof) FAIL("avp_fetch_profile: no results");
  CK(cudaSetDevice(ctx->device));
  CK(cudaMemcpy(out8n, ctx->d_prof, sizeof(long long) * (size_t)ctx->n * 16, cudaMemcpyDeviceToHost));
  return 0;
}

/* replaces compute_h.Dijkstra(map).compute_path(node_x, node_y) (compute_h.py:198-214) for scenario s:
 * stateful and resumable exactly like the reference object.  reset != 0 starts a fresh Dijkstra
 * object.  out: popped distance (-1: the reference would block forever), len(closedlist), target id.
 * The h table it fills is readable with avp_fetch_hvalues.  avp_plan_batch reuses the same per-id
 * arrays: a search run invalidates this state. */
extern "C" int avp_dijkstra_query(avp_ctx *ctx, int s, int reset, double node_x, double node_y, int32_t *dist, int32_t *closed_len, int32_t *target_id) {
  if (!ctx) return -3;
  if (s < 0 || s >= ctx->n) FAIL("avp_dijkstra_query: scenario index out of range");
  if (!ctx->rasterised) FAIL("avp_dijkstra_query: call avp_rasterise first");
  CK(cudaSetDevice(ctx->device));
  const int gcap = 1 << 20;
  if (!ctx->d_dq_gheap) {
    CK(cudaMalloc(&ctx->d_dq_gheap, sizeof(unsigned long long) * gcap));
    CK(cudaMalloc(&ctx->d_dq_state, sizeof(DijPersist)));
    CK(cudaMalloc(&ctx->d_dq_out, sizeof(int32_t) * 4));
    CK(cudaMemset(ctx->d_dq_state, 0, sizeof(DijPersist)));
  }
  if (ctx->dq_scen != s) { reset = 1; ctx->dq_scen = s; }
  k_dij_query<<<1, 32, 0, ctx->stream>>>(ctx->d_scen, s, ctx->d_cost, ctx->d_hval, ctx->d_ost, ctx->d_gx, ctx->d_gy, ctx->d_dq_gheap, gcap,
                                         ctx->d_dq_state, reset, node_x, node_y, ctx->d_dq_out);
  ctx->launches++;
  CK(cudaGetLastError());
  int32_t out[4];
  CK(cudaMemcpyAsync(out, ctx->d_dq_out, sizeof(out), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  if (dist) *dist = out[0]; if (closed_len) *closed_len = out[1]; if (target_id) *target_id = out[2];
  return 0;
}
