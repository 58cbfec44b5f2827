/* avp_b200.h -- C ABI of the B200-native hybrid-A* hot path (libavp_b200.so).
 *
 * The reference (wenqing-2021/AutomatedValetParking) is pure Python and has no FFI; its
 * boundary for this path is a set of Python classes.  Each entry point
 * below names the reference interface it stands in for.  All entry points
 *   - return int: 0 = OK, negative = API error (avp_last_error() has the text);
 *     per-scenario *planning* outcomes are in-band (avp_plan_summary.status);
 *   - take plain pointers + sizes; the caller allocates every output buffer; the library
 *     never retains a host pointer past the call;
 *   - run on one CUDA device through one avp_ctx (not thread-safe; one ctx per process/GPU).
 * There is NO CPU fallback: avp_create fails if no CUDA device is usable.
 */
#ifndef AVP_B200_H
#define AVP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AVP_MAX_STEER 16
#define AVP_MAX_RS_SEG 5
#define AVP_MAX_RS_POINTS 2048

/* ---- configuration: config/config.yaml keys + Vehicle constants (costmap.py:52-63) ---- */
typedef struct avp_config {
  /* config.yaml:2-8 */
  int32_t steering_angle_num;   /* 5 */
  int32_t n_substeps;           /* math.ceil(dt / trajectory_dt) evaluated by the host in Python (=3) */
  double dt;                    /* 0.6 */
  double ddt;                   /* trajectory_dt 0.2 */
  double map_discrete_size;     /* 0.1 */
  double flag_radius;           /* 18 */
  int32_t extended_num;         /* 1 */
  int32_t collision_mode;       /* 0 = 'distance' (default), 1 = 'circle' (config.yaml:18) */
  /* config.yaml:11-13 */
  double cost_gear, cost_heading_change, cost_scale;
  /* config.yaml:16-17 */
  double safe_side_dis, safe_fr_dis;
  /* Vehicle (costmap.py:52-63) */
  double lw, lf, lr, lb, max_steering_angle, max_v, min_radius_turn;
  /* np.linspace(-max_steer, max_steer, n) and np.tan of it, evaluated by the host with
   * numpy so that theta' = theta + (max_v*tan)/lw*dt is pure IEEE arithmetic on the
   * device (hybrid_a_star.py:81-83,146-148) */
  double steer[AVP_MAX_STEER];
  double tan_steer[AVP_MAX_STEER];
  /* capacity limits of the search (reference: unbounded Python lists) */
  int32_t max_pops;             /* nodes taken from the open list before status CAPACITY */
  int32_t reserved;
} avp_config;

/* ---- per-scenario planning outcome (in-band status) ---- */
enum {
  AVP_OK = 0,                 /* path found (reference: normal return of a_star_plan) */
  AVP_OPEN_EXHAUSTED = 1,     /* open list empty, last pop outside flag_radius: reference raises
                                 AttributeError at path_planner.py:104 (e.g. Case20) */
  AVP_OPEN_EXHAUSTED_RS = 2,  /* open list empty, last pop had a COLLIDING rs shot: reference
                                 silently returns astar_path + colliding rs path */
  AVP_H_UNREACHABLE = 3,      /* heuristic target unreachable: reference blocks forever in
                                 queue.get() (compute_h.py:77) */
  AVP_RS_DEGENERATE = 4,      /* rs word shorter than 0.01: reference AssertionError (rs_curve.py:153) */
  AVP_CAPACITY = 5,           /* max_pops / node / heap capacity reached (no reference analogue) */
  AVP_RASTER_AMBIGUOUS = 6    /* an edge sample matched two grid lines: reference TypeError
                                 (costmap.py:260) */
};

typedef struct avp_plan_summary {
  int32_t status;
  int32_t n_pops;          /* nodes taken from the open list (path_planner.py:70) */
  int32_t global_index;    /* hybrid_a_star.global_index = 10 * expansions (hybrid_a_star.py:239) */
  int32_t n_closed;        /* len(closed_list) */
  int32_t n_open;          /* len(open_list.queue) */
  int32_t n_astar;         /* len(a_star_path) = 1 + 3*depth (hybrid_a_star.py:351-389) */
  int32_t n_rs;            /* len(rs_path.x) */
  int32_t n_final;         /* len(final_path) = n_astar + n_rs - 1 */
  int32_t rs_nseg;
  int32_t last_index;      /* index of the last popped node */
  int32_t n_hq;            /* number of Dijkstra.compute_path calls (incl. the eager one) */
  int32_t h_closed;        /* len(Dijkstra.closedlist) at the end */
  int32_t nx, ny;          /* cost_map.shape */
  int32_t n_obs;           /* count(cost_map == 255) */
  int32_t n_hcalls;        /* calc_node_heuristic calls */
  double rs_L;
  double rs_lengths[AVP_MAX_RS_SEG];
  char rs_ctypes[8];       /* e.g. "LRLR", NUL padded */
  double origin[2];        /* (boundary[0], boundary[2]): origin of the fp32 relative path copy */
  double pitch[2];         /* _discrete_x, _discrete_y */
  double boundary[4];
  double last_pose[3];     /* x, y, theta of the last popped node: point 0 of rs_path (rs_curve.py:118-132) */
} avp_plan_summary;

typedef struct avp_ctx avp_ctx;

/* Library / device ----------------------------------------------------------------- */
/* replaces: nothing in the reference (it has no device); creates the per-GPU context. */
int avp_create(int device_id, const avp_config *cfg, avp_ctx **out);
int avp_destroy(avp_ctx *ctx);
const char *avp_last_error(const avp_ctx *ctx);
/* number of kernel launches issued by this context so far (bench.py "gpu_launches") */
int64_t avp_launch_count(const avp_ctx *ctx);

/* Scenarios ------------------------------------------------------------------------
 * replaces costmap.Case.read + Map.__init__ inputs (costmap.py:134-176): n scenarios, each
 * 6 pose doubles (x0,y0,theta0,xf,yf,thetaf); polygons flattened:
 *   obs_off[n+1]   : first obstacle of scenario i in nv[]
 *   nv[]           : vertex count per obstacle
 *   vert_off[]     : first vertex (in (x,y) pairs) of each obstacle in verts[]; length obs_off[n]+1
 *   verts[]        : x,y interleaved
 * boundary_override: NULL, or n*4 doubles replacing floor(min/max -/+ 12).
 * All pointers are HOST pointers; data is copied to the device (counted in e2e H2D bytes). */
int avp_scenarios_upload(avp_ctx *ctx, int n, const double *poses, const int32_t *obs_off,
                         const int32_t *nv, const int32_t *vert_off, const double *verts,
                         const double *boundary_override);

/* replaces Map.discrete_map + Map.detect_obstacle_edge (costmap.py:178-261) for every
 * uploaded scenario (kernel K1). */
int avp_rasterise(avp_ctx *ctx);

/* replaces reading Map.cost_map / map_position / boundary / _discrete_x/_y of scenario s.
 * dims[4]=(nx,ny,n_obs,raster_error); geom[6]=(b0,b1,b2,b3,dx,dy); cost_map: nx*ny bytes (0/255) indexed [ix*ny+iy],
 * may be NULL to query the sizes only. */
int avp_fetch_map(avp_ctx *ctx, int s, int32_t *dims, double *geom, uint8_t *cost_map, int64_t cap);

/* replaces distance_checker.check / two_circle_checker.check (collision_check.py:88-240)
 * for m poses (x,y,theta triples) against scenario s's raster; out[m] = 0/1. */
int avp_collision_check(avp_ctx *ctx, int s, int m, const double *poses, uint8_t *out);

/* the same check at every loaded scenario's own start and goal pose (headings wrapped with pi_2_pi as in
 * hybrid_a_star.py:105,109), one launch for the batch: out2n[2*s] = start collides, out2n[2*s+1] = goal collides.
 * What the scenario recipes of BASELINE configs 2-4 need ("reject a draw if the checker reports a collision at start or goal"). */
int avp_check_start_goal(avp_ctx *ctx, uint8_t *out2n);

/* replaces the obstacle-raster scan of path_opti.compute_collision_H (optimization/path_optimazition.py:221-658)
 * and ocp_optimization.compute_collision_H (optimization/ocp_optimization.py:36-480) for m path points
 * (x,y,theta triples, theta in [-pi, pi]) of scenario s: out4[4*i..] = x_max, y_max, x_min, y_min, the free
 * distances (<= expand_dis) the reference turns into H_max = (x_max + x, y_max + y), H_min = (x - x_min, y - y_min)
 * (:651-654).  status[i] = 1 (and NaNs) for a heading outside [-pi, pi]: the reference raises UnboundLocalError. */
int avp_corridor(avp_ctx *ctx, int s, int m, const double *poses, double expand_dis, double *out4, int32_t *status);

/* replaces hybrid_a_star.expand_node's pure part for one parent (hybrid_a_star.py:133-151,
 * 185-204) + rs length (calc_node_heuristic's h_value_2): for each of the 2*n primitives
 * out_pose[3] and out_flags (bit0 = collision on a sub-step, bit1 = outside boundary),
 * out_rsL.  Used by the single-step drop-in method and by parity tests. */
int avp_expand_pure(avp_ctx *ctx, int s, const double parent_pose[3], double *out_pose,
                    int32_t *out_flags, double *out_rsL);

/* replaces rs_curve.calc_optimal_path (rs_curve.py:99-134) for m (start, goal) pairs
 * (6 doubles each): selected word only.  lengths[m*5], ctypes[m*8], nseg[m], L[m];
 * course arrays (x, y, yaw: m*cap_pts doubles; dir: m*cap_pts int32) and n_pts[m].
 * nseg[i] = -1: no word; -2: degenerate (reference AssertionError, rs_curve.py:153).
 * xy_np / phi_np: whether the reference call would see numpy scalars for the normalised
 * (x, y) and for phi = gyaw - syaw.  CPython's sum() adds exact Python floats with Neumaier
 * compensation but numpy scalars naively, so set_path's length sums (rs_curve.py:145,148)
 * depend on it; on the planner path xy_np = 1 and phi_np = (node is not the root). */
int avp_rs_optimal(avp_ctx *ctx, int m, const double *q, double maxc, double step_size, int xy_np, int phi_np,
                   double *lengths, char *ctypes, int32_t *nseg, double *L, int cap_pts,
                   double *x, double *y, double *yaw, int32_t *dir, int32_t *n_pts);

/* replaces PathPlanner.a_star_plan for EVERY uploaded scenario (path_planner.py:58-110),
 * including Dijkstra (compute_h.py) and the rs shot: the whole search runs on the device.
 * summaries[n]; final_path: n*cap_path*3 doubles (x,y,theta rows, astar + rs[1:]);
 * pops: n*cap_pops int32 (node indices in pop order; may be NULL); all HOST pointers
 * (D2H copies inside the call).  Returns 0 even when individual scenarios fail (status). */
int avp_plan_batch(avp_ctx *ctx, avp_plan_summary *summaries, double *final_path, int cap_path,
                   int32_t *pops, int cap_pops);

/* same work with results left on the device: timing entry used by bench.py's resident
 * (`value`) leg. elapsed_ms = CUDA-event time of the search kernel alone. */
int avp_plan_batch_resident(avp_ctx *ctx, float *elapsed_ms);
/* copy results of the last avp_plan_batch_resident to host buffers (same layout as above) */
int avp_fetch_results(avp_ctx *ctx, avp_plan_summary *summaries, double *final_path, int cap_path,
                      int32_t *pops, int cap_pops);
/* allocate result buffers for the resident leg: cap_path rows per scenario, cap_pops pop slots */
int avp_plan_configure(avp_ctx *ctx, int cap_path, int cap_pops);
/* device pointers of the per-scenario result records (summaries: n * sizeof(avp_plan_summary);
 * paths: n * cap_path * 3 doubles) for the NCCL all-gather of finished trajectories */
int avp_result_device_buffer(avp_ctx *ctx, void **sums, void **paths, int64_t *n, int64_t *cap_path);
/* the first <= 256 Dijkstra.compute_path calls of scenario s: (terminate grid id, returned
 * distance, len(closedlist) afterwards) triples -- the observable trace of compute_h.py:198-214 */
int avp_fetch_hq_log(avp_ctx *ctx, int s, int32_t *log3, int cap_entries);
/* SM count, persistent-grid size (CTAs) and CTA width of the search kernel */
int avp_device_info(avp_ctx *ctx, int32_t *n_sm, int32_t *slots, int32_t *block);

/* per-scenario h-table readback: replaces iterating Dijkstra.closedlist (compute_h.py:80):
 * hval[id] = distance of the first closedlist entry with grid_id == id, -1 if none. */
int avp_fetch_hvalues(avp_ctx *ctx, int s, int32_t *hval, int64_t cap, int64_t *n_ids);

/* replaces compute_h.Dijkstra(map).compute_path(node_x, node_y) (compute_h.py:198-214) for scenario s:
 * stateful and resumable like the reference object; reset != 0 starts a fresh Dijkstra object.
 * dist = popped distance of the target cell (-1: the reference would block forever in queue.get()),
 * closed_len = len(closedlist), target_id = terminate_grid_id.  The h table is read with
 * avp_fetch_hvalues.  A plan run reuses the per-id arrays and invalidates this state. */
int avp_dijkstra_query(avp_ctx *ctx, int s, int reset, double node_x, double node_y, int32_t *dist, int32_t *closed_len, int32_t *target_id);

/* parity aid for calc_node_cost / calc_node_heuristic (hybrid_a_star.py:243-298): with on != 0 the following plans record f, g, h
 * of every popped node as they are at open_list.get() (path_planner.py:70), beside the pop indices (needs cap_pops > 0);
 * avp_fetch_pop_fgh copies them out: out[n][cap_pops][3]. */
int avp_trace_fgh(avp_ctx *ctx, int on);
int avp_fetch_pop_fgh(avp_ctx *ctx, double *out, int cap_pops);

/* replaces PathPlanner.split_path (path_planner.py:112-192): gear-change detection (scipy.spatial.distance.cosine semantics,
 * NaN for a zero displacement) and the collision-checked extension points, for every finished plan of the batch (results of
 * the last avp_plan_batch, still on the device), one launch.  Per scenario: split_pts[cap_pts][3] = the segments back to back
 * (= out_final_path of path_planning, :52), seg_len[cap_seg] = their lengths (split_path_list), info[4] = {status, number of
 * segments, change_gear, number of points}.  status: 0 ok; 1 the path has no gear change (the reference raises IndexError at
 * :181); 2 the scenario has no path (plan status != AVP_OK); 3 cap_pts / cap_seg / the plan's cap_path too small. */
int avp_split_paths(avp_ctx *ctx, int cap_pts, int cap_seg, double *split_pts, int32_t *seg_len, int32_t *info);
/* the same for ONE caller-supplied path (n_pts rows x, y, theta) against scenario s's raster: PathPlanner.split_path(final_path) */
int avp_split_path(avp_ctx *ctx, int s, int n_pts, const double *path, int cap_pts, int cap_seg, double *split_pts, int32_t *seg_len, int32_t *info4);

/* CUDA-event stopwatch on the context's stream around any sequence of entry points, and the
 * CUDA-event duration of the most recent search-kernel launch (bench.py timing legs) */
int avp_timer_start(avp_ctx *ctx);
int avp_timer_stop(avp_ctx *ctx, float *elapsed_ms);
int avp_last_search_ms(avp_ctx *ctx, float *elapsed_ms);
/* the search is one eager-Dijkstra launch (one warp per scenario) and ONE persistent search launch with a round-robin run
 * queue: their CUDA-event times; n_info = number of times a search let go of its SM at the end of a quantum
 * (mod 100000) + 100000 * (CTA width of the search kernel) */
int avp_last_search_passes(avp_ctx *ctx, float *ms_dijkstra, float *ms_search, int32_t *n_info);
/* batches larger than two waves of wide CTAs are searched by two launches (narrow CTAs give every scenario a few pops and finish the
 * short searches; the wide launch resumes the rest from the run queue): CUDA-event time of the narrow launch, part of ms_search */
int avp_last_narrow_ms(avp_ctx *ctx, float *ms);

/* development aids: an in-kernel watchdog (SM clock cycles per scenario, 0 = off; a scenario
 * that exceeds it ends with AVP_CAPACITY) and the per-scenario progress checkpoints
 * (8 int32 each: phase, pops, len(closedlist), open size, ...).  The environment variable
 * AVP_HOST_TIMEOUT_S bounds the host's wait for the search kernel. */
int avp_set_watchdog(avp_ctx *ctx, long long cycles);
int avp_fetch_debug(avp_ctx *ctx, int32_t *out8n);
/* per-scenario SM-cycle accumulators / counters of the search kernel (16 int64 each, see avp_api.cu) */
int avp_fetch_profile(avp_ctx *ctx, int64_t *out8n);
/* pipelined pass-2 kernel only: per scenario and warp (16 x 24 int64): [0..7] the cycles lane 0 of the warp spent
 * WORKING in each evaluator phase (time at the phase barriers excluded), [8..23] absolute clocks of the warp at
 * the phase boundaries of the pop selected with the environment variable AVP_TRACE_POP (development aid) */
int avp_fetch_warp_profile(avp_ctx *ctx, int64_t *out384n);

#ifdef __cplusplus
}
#endif
#endif /* AVP_B200_H */
