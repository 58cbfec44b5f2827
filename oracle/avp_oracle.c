/* avp_oracle.c -- CPU restatement of the reference's hybrid-A* hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in automatedvaletparking_b200/ may import, link or call
 * this file; it is the checker used by tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs.
 *
 * Parity pin: the reference has no tests or golden vectors for this path (SURVEY.md §4,
 * §8c).  This restatement is pinned against the UNMODIFIED reference run in the build
 * container: tests/golden/cases/Case*.npz (full planner traces of the 20 BenchmarkCases,
 * made by tests/golden/gen_case_golden.py) and tests/golden/leaf_*.npz (per-function
 * vectors made by tests/golden/gen_leaf_golden.py).  tests/test_oracle_*.py check it
 * against those fixtures.
 *
 * Every function cites the reference file:line it follows (paths relative to the
 * reference root).  Arithmetic is IEEE fp64 with the host libm, compiled with
 * -ffp-contract=off; the few fused multiply-adds that numpy's BLAS performs are explicit.
 */
#define _GNU_SOURCE
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>
#include "../include/avp_b200.h"

/* libm calls go through volatile pointers so the compiler cannot fold pow(x,2) -> x*x */
static double (*volatile p_pow)(double, double) = pow;
static double (*volatile p_fmod)(double, double) = fmod;

#define PI 3.141592653589793 /* math.pi (rs_curve.py:25) */

/* ------------------------------------------------------------------ numpy / CPython helpers */

/* CPython 3.12 builtin sum() over floats: first element via 0 + x, the rest with Neumaier
 * compensation, compensation folded in at the end (Python/bltinmodule.c builtin_sum_impl). */
/* `npmask` bit i set = element i is an np.float64 (not an exact Python float): the float
 * fast path (Neumaier) only covers the leading run of exact floats; at the first np.float64
 * the compensation is folded in and the rest is added naively (PyNumber_Add).  Which
 * elements are np.float64 follows from the reference's own type propagation, see
 * generate_path(). */
static double py_sum(const double *v, int n, unsigned npmask) {
  if (n == 0) return 0.0;
  double f = 0.0 + v[0], c = 0.0;
  int i = 1;
  if (!(npmask & 1u)) {
    for (; i < n && !(npmask & (1u << i)); ++i) {
      double x = v[i], t = f + x;
      if (fabs(f) >= fabs(x)) c += (f - t) + x; else c += (x - t) + f;
      f = t;
    }
    if (c != 0.0 && isfinite(c)) f += c;
  }
  for (; i < n; ++i) f = f + v[i];
  return f;
}

/* CPython 3.12 math.hypot(x, y) (Modules/mathmodule.c vector_norm, n = 2): scaled,
 * Dekker-split squares, compensated sum, one differential correction step. */
typedef struct { double hi, lo; } dl_t;
static dl_t dl_fast_sum(double a, double b) { dl_t r; r.hi = a + b; r.lo = (a - r.hi) + b; return r; }
static dl_t dl_split(double x) { double t = x * 134217729.0; dl_t r; r.hi = t - (t - x); r.lo = x - r.hi; return r; }
static dl_t dl_mul(double x, double y) {
  dl_t xx = dl_split(x), yy = dl_split(y);
  double p = xx.hi * yy.hi, q = xx.hi * yy.lo + xx.lo * yy.hi;
  dl_t r; r.hi = p + q; r.lo = p - r.hi + q + xx.lo * yy.lo; return r;
}
static double py_hypot(double a, double b) {
  double vec[2] = {fabs(a), fabs(b)};
  double max = vec[0] > vec[1] ? vec[0] : vec[1];
  if (isinf(vec[0]) || isinf(vec[1])) return INFINITY;
  if (isnan(vec[0]) || isnan(vec[1])) return NAN;
  if (max == 0.0) return max;
  int max_e; frexp(max, &max_e);
  if (max_e < -1023) return DBL_MIN * py_hypot(vec[0] / DBL_MIN, vec[1] / DBL_MIN);
  double scale = ldexp(1.0, -max_e), csum = 1.0, frac1 = 0.0, frac2 = 0.0, x, h;
  dl_t pr, sm;
  for (int i = 0; i < 2; ++i) {
    x = vec[i] * scale; pr = dl_mul(x, x); sm = dl_fast_sum(csum, pr.hi);
    csum = sm.hi; frac1 += pr.lo; frac2 += sm.lo;
  }
  h = sqrt(csum - 1.0 + (frac1 + frac2));
  pr = dl_mul(-h, h); sm = dl_fast_sum(csum, pr.hi); csum = sm.hi; frac1 += pr.lo; frac2 += sm.lo;
  x = csum - 1.0 + (frac1 + frac2);
  h += x / (2.0 * h);
  return h / scale;
}

/* Python float %: fmod then sign fix (Objects/floatobject.c float_rem) */
static double py_mod(double v, double w) {
  double m = p_fmod(v, w);
  if (m != 0.0) { if ((w < 0) != (m < 0)) m += w; } else m = copysign(0.0, w);
  return m;
}

/* np.linspace(start, stop, num)[k] (numpy/_core/function_base.py): k*step + start, last = stop */
static double np_linspace_at(double start, double stop, int num, int k) {
  if (num == 1) return 0.0 * (stop - start) + start; /* div = 0 branch: y = y*delta */
  int div = num - 1;
  double delta = stop - start, step = delta / div, y;
  if (k == num - 1) return stop;
  if (step == 0.0) y = ((double)k / div) * delta; else y = (double)k * step;
  return y + start;
}

/* np.add.reduce over a strided 1-D float64 (pairwise_sum in loops_utils.h), n < 128 here */
static double np_sum(const double *a, int n, int stride) {
  if (n < 8) { double r = -0.0; for (int i = 0; i < n; ++i) r += a[i * stride]; return r; }
  double r[8]; int i;
  for (i = 0; i < 8; ++i) r[i] = a[i * stride];
  for (i = 8; i < n - (n % 8); i += 8) for (int j = 0; j < 8; ++j) r[j] += a[(i + j) * stride];
  double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
  for (; i < n; ++i) res += a[i * stride];
  return res;
}

/* ------------------------------------------------------------------ map (map/costmap.py) */

typedef struct orc_map {
  double pose[6];          /* x0,y0,theta0,xf,yf,thetaf (costmap.py:141-142) */
  double boundary[4];      /* costmap.py:169-172 */
  double ds;               /* discrete_size */
  int nx, ny;              /* cost_map.shape (costmap.py:182-186) */
  double dx, dy;           /* _discrete_x/_y (costmap.py:190-191) */
  uint8_t *cost;           /* [ix*ny+iy] in {0,255} */
  double *xs, *ys;         /* map_position */
  int n_obs; int *obs_ix, *obs_iy; double *obs_x, *obs_y; /* np.where(cost_map==255) order */
  int raster_error;        /* 1 if an edge sample matched >1 grid line (reference TypeError) */
} orc_map;

static int cmp_rows(const void *a, const void *b) {
  const double *p = a, *q = b;
  if (p[0] < q[0]) return -1; if (p[0] > q[0]) return 1;
  if (p[1] < q[1]) return -1; if (p[1] > q[1]) return 1; return 0;
}

static void map_collect_obstacles(orc_map *m) {
  int n = 0;
  for (int i = 0; i < m->nx * m->ny; ++i) n += (m->cost[i] == 255);
  free(m->obs_ix); free(m->obs_iy); free(m->obs_x); free(m->obs_y);
  m->n_obs = n;
  m->obs_ix = malloc(sizeof(int) * (n + 1)); m->obs_iy = malloc(sizeof(int) * (n + 1));
  m->obs_x = malloc(sizeof(double) * (n + 1)); m->obs_y = malloc(sizeof(double) * (n + 1));
  n = 0;
  for (int ix = 0; ix < m->nx; ++ix) for (int iy = 0; iy < m->ny; ++iy)
    if (m->cost[ix * m->ny + iy] == 255) { /* collision_check.py:55-57 */
      m->obs_ix[n] = ix; m->obs_iy[n] = iy; m->obs_x[n] = m->xs[ix]; m->obs_y[n] = m->ys[iy]; ++n;
    }
}

/* costmap.py:134-156 (extents), :160-176, :178-195, :197-261 */
orc_map *orc_map_build(const double pose[6], int obs_num, const int32_t *nv, const double *verts,
                       double discrete_size, const double *boundary_override) {
  orc_map *m = calloc(1, sizeof(*m));
  memcpy(m->pose, pose, sizeof(double) * 6);
  m->ds = discrete_size;
  if (boundary_override) memcpy(m->boundary, boundary_override, sizeof(double) * 4);
  else {
    double x0 = pose[0], y0 = pose[1], xf = pose[3], yf = pose[4];
    double xmin = (x0 < xf ? x0 : xf) - 12, xmax = (x0 > xf ? x0 : xf) + 12;   /* python min/max */
    double ymin = (y0 < yf ? y0 : yf) - 12, ymax = (y0 > yf ? y0 : yf) + 12;
    m->boundary[0] = floor(xmin); m->boundary[1] = floor(xmax);
    m->boundary[2] = floor(ymin); m->boundary[3] = floor(ymax);
  }
  const double *b = m->boundary;
  m->nx = (int)((b[1] - b[0]) / discrete_size);
  m->ny = (int)((b[3] - b[2]) / discrete_size);
  m->cost = calloc((size_t)m->nx * m->ny, 1);
  m->xs = malloc(sizeof(double) * m->nx); m->ys = malloc(sizeof(double) * m->ny);
  for (int k = 0; k < m->nx; ++k) m->xs[k] = np_linspace_at(b[0], b[1], m->nx, k);
  for (int k = 0; k < m->ny; ++k) m->ys[k] = np_linspace_at(b[2], b[3], m->ny, k);
  m->dx = m->xs[1] - m->xs[0]; m->dy = m->ys[1] - m->ys[0];

  int voff = 0;
  for (int o = 0; o < obs_num; ++o) {
    int n0 = nv[o];
    double *P = malloc(sizeof(double) * 2 * n0);
    memcpy(P, verts + 2 * voff, sizeof(double) * 2 * n0); voff += n0;
    /* np.unique(axis=0): lexicographic row sort + drop duplicates (costmap.py:206) */
    qsort(P, n0, 2 * sizeof(double), cmp_rows);
    int n = 0;
    for (int i = 0; i < n0; ++i)
      if (n == 0 || P[2 * i] != P[2 * (n - 1)] || P[2 * i + 1] != P[2 * (n - 1) + 1]) { P[2 * n] = P[2 * i]; P[2 * n + 1] = P[2 * i + 1]; ++n; }
    double cx = np_sum(P, n, 2) / n, cy = np_sum(P + 1, n, 2) / n;            /* :210-211 */
    double *ang = malloc(sizeof(double) * n); int *ord = malloc(sizeof(int) * n);
    for (int i = 0; i < n; ++i) { ang[i] = atan2(P[2 * i + 1] - cy, P[2 * i] - cx) + PI; ord[i] = i; } /* :215 */
    for (int i = 1; i < n; ++i) { int t = ord[i], j = i; while (j > 0 && ang[ord[j - 1]] > ang[t]) { ord[j] = ord[j - 1]; --j; } ord[j] = t; } /* argsort */
    for (int j = 0; j < n; ++j) {
      const double *p1 = P + 2 * ord[j], *p2 = P + 2 * ord[(j + 1 == n) ? 0 : j + 1];   /* :218-223 */
      double vx = p2[0] - p1[0], vy = p2[1] - p1[1];
      double ra = atan2(vy, vx), c = cos(ra), s = sin(ra);                               /* :229-232 */
      double len = fma(c, vx, s * vy);      /* np.dot(rotation_matrix, v)[0]: BLAS gemv = fma(M00,v0, M01*v1) */
      int points_num = (int)floor(len / m->dx);                                          /* :240-241 */
      for (int k = 0; k < points_num; ++k) {
        double px = np_linspace_at(0.0, len, points_num, k);
        double ox = c * px + p1[0], oy = s * px + p1[1];   /* np.dot(R^T, [px;0]) then + p1 (:246-251) */
        int ixm = -1, iym = -1, nxm = 0, nym = 0;
        for (int i = 0; i < m->nx; ++i) if (m->xs[i] < ox && m->xs[i] > ox - m->dx) { if (!nxm) ixm = i; ++nxm; } /* :253-254 */
        for (int i = 0; i < m->ny; ++i) if (m->ys[i] < oy && m->ys[i] > oy - m->dy) { if (!nym) iym = i; ++nym; } /* :256-257 */
        if (nxm > 0 && nym > 0) { if (nxm > 1 || nym > 1) m->raster_error = 1; m->cost[ixm * m->ny + iym] = 255; }     /* :259-261 */
      }
    }
    free(P); free(ang); free(ord);
  }
  map_collect_obstacles(m);
  return m;
}

void orc_map_free(orc_map *m) {
  if (!m) return;
  free(m->cost); free(m->xs); free(m->ys); free(m->obs_ix); free(m->obs_iy); free(m->obs_x); free(m->obs_y); free(m);
}
void orc_map_info(const orc_map *m, int32_t *dims, double *geom, int32_t *n_obs, int32_t *err) {
  dims[0] = m->nx; dims[1] = m->ny;
  memcpy(geom, m->boundary, 4 * sizeof(double)); geom[4] = m->dx; geom[5] = m->dy;
  *n_obs = m->n_obs; *err = m->raster_error;
}
void orc_map_cost(const orc_map *m, uint8_t *out) { memcpy(out, m->cost, (size_t)m->nx * m->ny); }
void orc_map_positions(const orc_map *m, double *xs, double *ys) { memcpy(xs, m->xs, sizeof(double) * m->nx); memcpy(ys, m->ys, sizeof(double) * m->ny); }

/* costmap.py:319-329 */
static long map_index(const orc_map *m, double gx, double gy) {
  long i0 = (long)floor((gx - m->boundary[0]) / m->dx);
  long i1 = (long)floor((m->boundary[3] - gy) / m->dy) * (long)(int)((m->boundary[1] - m->boundary[0]) / m->dx);
  return i0 + i1;
}
long orc_convert_position_to_index(const orc_map *m, double gx, double gy) { return map_index(m, gx, gy); }

/* ------------------------------------------------------------------ collision (collision_check.py) */

/* Vehicle.create_anticlockpoint (costmap.py:85-121): R^T.dot(local) + [x, y].  The BLAS
 * gemv on the transposed view evaluates row r as fma(A[r][1], v1, A[r][0]*v0). */
static void vehicle_corners(const avp_config *cfg, double x, double y, double th, double vb[5][2]) {
  double c = cos(th), s = sin(th);
  double T[2][2] = {{c, -s}, {s, c}};     /* trans_matrix.transpose() */
  double fr = cfg->safe_fr_dis, sd = cfg->safe_side_dis;
  double loc[4][2] = {{-cfg->lr - fr, -cfg->lb / 2 - sd}, {cfg->lw + cfg->lf + fr, -cfg->lb / 2 - sd},
                      {cfg->lw + cfg->lf + fr, cfg->lb / 2 + sd}, {-cfg->lr - fr, cfg->lb / 2 + sd}};
  for (int i = 0; i < 4; ++i) {
    vb[i][0] = fma(T[0][1], loc[i][1], T[0][0] * loc[i][0]) + x;
    vb[i][1] = fma(T[1][1], loc[i][1], T[1][0] * loc[i][0]) + y;
  }
  vb[4][0] = vb[0][0]; vb[4][1] = vb[0][1];
}
void orc_vehicle_corners(const avp_config *cfg, double x, double y, double th, double *out10) {
  double vb[5][2]; vehicle_corners(cfg, x, y, th, vb); memcpy(out10, vb, sizeof(vb));
}

/* distance_checker.check (collision_check.py:144-240) with get_near_obstacles (:29-73) */
static int check_distance(const orc_map *m, const avp_config *cfg, double x, double y, double th) {
  double vb[5][2];
  vehicle_corners(cfg, x, y, th, vb);
  double x_max = vb[0][0], x_min = vb[0][0], y_max = vb[0][1], y_min = vb[0][1];
  for (int i = 1; i < 5; ++i) {                       /* python max()/min(): strict comparisons */
    if (vb[i][0] > x_max) x_max = vb[i][0]; if (vb[i][0] < x_min) x_min = vb[i][0];
    if (vb[i][1] > y_max) y_max = vb[i][1]; if (vb[i][1] < y_min) y_min = vb[i][1];
  }
  double d0 = vb[0][0] - vb[3][0], d1 = vb[0][1] - vb[3][1];
  double v_lb = sqrt(d0 * d0 + d1 * d1);                                             /* :165-166 */
  d0 = vb[3][0] - vb[2][0]; d1 = vb[3][1] - vb[2][1];
  double v_length = sqrt(d0 * d0 + d1 * d1);                                         /* :168-169 */
  double lk[4], lb_[4];
  for (int i = 0; i < 4; ++i) {                                                      /* :149-155,:180-190 */
    const double *p1 = vb[i], *p2 = vb[(i < 3) ? i + 1 : 0];
    lk[i] = (p2[1] - p1[1]) / (p2[0] - p1[0]);
    lb_[i] = p1[1] - lk[i] * p1[0];
  }
  int collision = 0;
  for (int n = 0; n < m->n_obs; ++n) {
    double ox = m->obs_x[n], oy = m->obs_y[n];
    if (!(ox >= x_min && ox <= x_max)) continue;                                      /* :60-63 */
    if (!(oy >= y_min && oy <= y_max)) continue;                                      /* :66-69 */
    double dis[4];
    for (int i = 0; i < 4; ++i) dis[i] = fabs(lk[i] * ox + lb_[i] - oy) / sqrt(1 + lk[i] * lk[i]); /* :158-160 */
    int check_1 = fabs(dis[0] - dis[2]) < v_lb - 0.01;                                /* :202 */
    int check_2 = fabs(dis[1] - dis[3]) < v_length - 0.01;                            /* :203 */
    if (check_1 && check_2) { collision = 1; break; }
    if (!collision) {                                                                 /* :210-230 */
      int on_x = 0, on_y = 0;
      for (int i = 0; i < 5; ++i) if (ox == vb[i][0]) { on_x = 1; break; }
      if (on_x) for (int i = 0; i < 5; ++i) if (oy == vb[i][1]) { on_y = 1; break; }
      if (on_x && on_y) { collision = 1; break; }
    }
    if (!collision) {                                                                 /* :233-238 */
      for (int i = 0; i < 4; ++i) {
        double k1 = (vb[i][1] - oy) / (vb[i][0] - ox);
        if (k1 == lk[i]) { collision = 1; break; }
      }
    }
  }
  return collision;
}

/* two_circle_checker.check (collision_check.py:88-137) */
static int check_circle(const orc_map *m, const avp_config *cfg, double x, double y, double th) {
  double lr = cfg->lr, lw = cfg->lw, lf = cfg->lf, lb = cfg->lb;
  double Rd = 0.5 * sqrt(p_pow((lr + lw + lf) / 2, 2.0) + p_pow(lb, 2.0));
  double c = cos(th), s = sin(th);
  double kf = 1.0 / 4 * (3 * lw + 3 * lf - lr), kr = 1.0 / 4 * (lw + lf - 3 * lr);
  double fx = x + kf * c, fy = y + kf * s, rx = x + kr * c, ry = y + kr * s;
  double right, left, upper, down;
  if (fx >= rx) { right = fx + Rd; left = rx - Rd; } else { right = rx + Rd; left = fx - Rd; }
  if (fy >= ry) { upper = fy + Rd; down = ry - Rd; } else { upper = ry + Rd; down = fy - Rd; }
  int collision = 0;
  for (int n = 0; n < m->n_obs; ++n) {
    double ox = m->obs_x[n], oy = m->obs_y[n];
    if (!(ox > left && ox < right)) continue;
    if (!(oy > down && oy < upper)) continue;
    if (sqrt(p_pow(ox - fx, 2.0) + p_pow(oy - fy, 2.0)) <= Rd) collision = 1;
    else if (sqrt(p_pow(ox - rx, 2.0) + p_pow(oy - ry, 2.0)) <= Rd) collision = 1;
  }
  return collision;
}

int orc_check(const orc_map *m, const avp_config *cfg, double x, double y, double th) {
  return cfg->collision_mode == 1 ? check_circle(m, cfg, x, y, th) : check_distance(m, cfg, x, y, th);
}

/* ------------------------------------------------------------------ corridor extraction */

/* path_opti.compute_collision_H (optimization/path_optimazition.py:221-658) and its copy
 * ocp_optimization.compute_collision_H (optimization/ocp_optimization.py:36-480), per path point:
 * the free distances x_max, y_max, x_min, y_min (each <= expand_dis) from the inflated vehicle rectangle
 * to the obstacle raster cells around it.  out4[i] = {x_max, y_max, x_min, y_min}; the reference then forms
 * H_max = (x_max + x, y_max + y), H_min = (x - x_min, y - y_min) (:651-654).  A heading outside [-pi, pi]
 * leaves `case` unbound in the reference (UnboundLocalError): status[i] = 1 and NaNs.
 * Quirks kept: vehicle_boundary has shape (5,2,1), so k, b are 1-element arrays and pow(_k, 2) is
 * numpy's array power (a multiplication, not libm pow); the areas are tested in the order right, front,
 * left, rear and the first hit wins (break); inf/nan slopes of axis-aligned headings flow through IEEE
 * comparisons; a quotient by |sin|,|cos| = 0 gives inf/nan and never updates a minimum. */
void orc_corridor(const orc_map *m, const avp_config *cfg, double expand_dis, int n, const double *poses, double *out4, int32_t *status) {
  static const int base_x[4] = {+1, +1, -1, -1}, base_y[4] = {-1, +1, +1, -1};   /* case 1: right, front, left, rear */
  const double e = expand_dis;
  for (int p = 0; p < n; ++p) {
    const double x = poses[3 * p], y = poses[3 * p + 1], th = poses[3 * p + 2];
    double *o = out4 + 4 * p;
    int shift;
    if (th >= -PI && th < -PI / 2) shift = 2;            /* case 3 */
    else if (th >= -PI / 2 && th < 0) shift = 3;         /* case 4 */
    else if (th >= 0 && th < PI / 2) shift = 0;          /* case 1 */
    else if (th >= PI / 2 && th <= PI) shift = 1;        /* case 2 */
    else { o[0] = o[1] = o[2] = o[3] = NAN; if (status) status[p] = 1; continue; }
    if (status) status[p] = 0;
    double vb[5][2];
    vehicle_corners(cfg, x, y, th, vb);
    double bx_max = vb[0][0], bx_min = vb[0][0], by_max = vb[0][1], by_min = vb[0][1];
    for (int i = 1; i < 5; ++i) {
      if (vb[i][0] > bx_max) bx_max = vb[i][0]; if (vb[i][0] < bx_min) bx_min = vb[i][0];
      if (vb[i][1] > by_max) by_max = vb[i][1]; if (vb[i][1] < by_min) by_min = vb[i][1];
    }
    bx_max += e; bx_min -= e; by_max += e; by_min -= e;                          /* :255-258 */
    double lk[4], lb_[4], area[4][4];
    for (int i = 0; i < 4; ++i) {
      const double *p1 = vb[i], *p2 = vb[(i < 3) ? i + 1 : 0];
      lk[i] = (p2[1] - p1[1]) / (p2[0] - p1[0]);                                  /* compute_k_b :282-287 */
      lb_[i] = p1[1] - lk[i] * p1[0];
      area[i][0] = (p2[0] < p1[0]) ? p2[0] : p1[0]; area[i][1] = (p2[0] > p1[0]) ? p2[0] : p1[0];   /* get_area_boundary :289-294 (python min/max) */
      area[i][2] = (p2[1] < p1[1]) ? p2[1] : p1[1]; area[i][3] = (p2[1] > p1[1]) ? p2[1] : p1[1];
    }
    const double as = fabs(sin(th)), ac = fabs(cos(th));
    double x_max = e, x_min = e, y_max = e, y_min = e;
    for (int c = 0; c < m->n_obs; ++c) {
      const double ox = m->obs_x[c], oy = m->obs_y[c];
      if (!(ox >= bx_min && ox <= bx_max)) continue;                              /* :266-269 */
      if (!(oy >= by_min && oy <= by_max)) continue;                              /* :272-275 */
      for (int k = 0; k < 4; ++k) {
        const int sx = base_x[(k + shift) & 3], sy = base_y[(k + shift) & 3];
        const double ax0 = (sx < 0) ? area[k][0] - e : area[k][0], ax1 = (sx > 0) ? area[k][1] + e : area[k][1];
        const double ay0 = (sy < 0) ? area[k][2] - e : area[k][2], ay1 = (sy > 0) ? area[k][3] + e : area[k][3];
        if (ox > ax0 && ox < ax1 && oy > ay0 && oy < ay1) {
          const double sd = fabs(lk[k] * ox + lb_[k] - oy) / sqrt(1 + lk[k] * lk[k]);     /* compute_distance :296-298 */
          const double ver = sd / ac, hor = sd / as;                                      /* :303-305 */
          if (sx > 0) { if (hor < x_max) x_max = hor; } else { if (hor < x_min) x_min = hor; }
          if (sy > 0) { if (ver < y_max) y_max = ver; } else { if (ver < y_min) y_min = ver; }
          break;
        }
      }
    }
    o[0] = x_max; o[1] = y_max; o[2] = x_min; o[3] = y_min;
  }
}

/* ------------------------------------------------------------------ Reeds-Shepp (path_plan/rs_curve.py) */

typedef struct { int n; double len[5]; char ct[6]; double L; } rs_word;
typedef struct { int n; rs_word w[64]; int degenerate; } rs_set;

/* rs_curve.py:649-656 */
static double pi_2_pi(double th) { while (th > PI) th -= 2.0 * PI; while (th < -PI) th += 2.0 * PI; return th; }
double orc_pi_2_pi(double th) { return pi_2_pi(th); }
/* rs_curve.py:669-680 */
static double M(double th) { double phi = py_mod(th, 2.0 * PI); if (phi < -PI) phi += 2.0 * PI; if (phi > PI) phi -= 2.0 * PI; return phi; }
/* rs_curve.py:659-666 */
static void R(double x, double y, double *r, double *th) { *r = py_hypot(x, y); *th = atan2(y, x); }

/* rs_curve.py:137-156 */
static void set_path(rs_set *S, const double *len, const char *ct, unsigned npmask) {
  int n = (int)strlen(ct);
  for (int e = 0; e < S->n; ++e) {
    if (strcmp(S->w[e].ct, ct) == 0) {
      double d[5]; for (int i = 0; i < n; ++i) d[i] = S->w[e].len[i] - len[i];
      if (py_sum(d, n, npmask) <= 0.01) return;
    }
  }
  double a[5]; for (int i = 0; i < n; ++i) a[i] = fabs(len[i]);
  double L = py_sum(a, n, npmask);
  if (L >= 1000.0) return;            /* MAX_LENGTH (rs_curve.py:24) */
  if (!(L >= 0.01)) { S->degenerate = 1; return; }   /* assert (rs_curve.py:153) */
  rs_word *w = &S->w[S->n++]; w->n = n; memcpy(w->len, len, sizeof(double) * n); strcpy(w->ct, ct); w->L = L;
}

/* rs_curve.py:159-167 */
static int LSL(double x, double y, double phi, double *t, double *u, double *v) {
  double uu, tt; R(x - sin(phi), y - 1.0 + cos(phi), &uu, &tt);
  if (tt >= 0.0) { double vv = M(phi - tt); if (vv >= 0.0) { *t = tt; *u = uu; *v = vv; return 1; } }
  return 0;
}
/* rs_curve.py:170-183 */
static int LSR(double x, double y, double phi, double *t, double *u, double *v) {
  double u1, t1; R(x + sin(phi), y - 1.0 - cos(phi), &u1, &t1);
  u1 = p_pow(u1, 2.0);
  if (u1 >= 4.0) {
    double uu = sqrt(u1 - 4.0), theta = atan2(2.0, uu), tt = M(t1 + theta), vv = M(tt - phi);
    if (tt >= 0.0 && vv >= 0.0) { *t = tt; *u = uu; *v = vv; return 1; }
  }
  return 0;
}
/* rs_curve.py:186-197 */
static int LRL(double x, double y, double phi, double *t, double *u, double *v) {
  double u1, t1; R(x - sin(phi), y - 1.0 + cos(phi), &u1, &t1);
  if (u1 <= 4.0) {
    double uu = -2.0 * asin(0.25 * u1), tt = M(t1 + 0.5 * uu + PI), vv = M(phi - tt + uu);
    if (tt >= 0.0 && uu <= 0.0) { *t = tt; *u = uu; *v = vv; return 1; }
  }
  return 0;
}
/* rs_curve.py:213-229 */
static int SLS(double x, double y, double phi, double *t, double *u, double *v) {
  phi = M(phi);
  if (y > 0.0 && 0.0 < phi && phi < PI * 0.99) {
    double xd = -y / tan(phi) + x;
    *t = xd - tan(phi / 2.0); *u = phi;
    *v = sqrt(p_pow(x - xd, 2.0) + p_pow(y, 2.0)) - tan(phi / 2.0);
    return 1;
  } else if (y < 0.0 && 0.0 < phi && phi < PI * 0.99) {
    double xd = -y / tan(phi) + x;
    *t = xd - tan(phi / 2.0); *u = phi;
    *v = -sqrt(p_pow(x - xd, 2.0) + p_pow(y, 2.0)) - tan(phi / 2.0);
    return 1;
  }
  return 0;
}
/* rs_curve.py:308-323 */
static void calc_tauOmega(double u, double v, double xi, double eta, double phi, double *tau, double *omega) {
  double delta = M(u - v), A = sin(u) - sin(delta), B = cos(u) - cos(delta) - 1.0;
  double t1 = atan2(eta * A - xi * B, xi * A + eta * B);
  double t2 = 2.0 * (cos(delta) - cos(v) - cos(u)) + 3.0;
  *tau = (t2 < 0) ? M(t1 + PI) : M(t1);
  *omega = M(*tau - u + v - phi);
}
/* rs_curve.py:326-337 */
static int LRLRn(double x, double y, double phi, double *t, double *u, double *v) {
  double xi = x + sin(phi), eta = y - 1.0 - cos(phi), rho = 0.25 * (2.0 + sqrt(xi * xi + eta * eta));
  if (rho <= 1.0) {
    double uu = acos(rho), tt, vv; calc_tauOmega(uu, -uu, xi, eta, phi, &tt, &vv);
    if (tt >= 0.0 && vv <= 0.0) { *t = tt; *u = uu; *v = vv; return 1; }
  }
  return 0;
}
/* rs_curve.py:340-352 */
static int LRLRp(double x, double y, double phi, double *t, double *u, double *v) {
  double xi = x + sin(phi), eta = y - 1.0 - cos(phi), rho = (20.0 - xi * xi - eta * eta) / 16.0;
  if (0.0 <= rho && rho <= 1.0) {
    double uu = -acos(rho);
    if (uu >= -0.5 * PI) {
      double tt, vv; calc_tauOmega(uu, uu, xi, eta, phi, &tt, &vv);
      if (tt >= 0.0 && vv >= 0.0) { *t = tt; *u = uu; *v = vv; return 1; }
    }
  }
  return 0;
}
/* rs_curve.py:391-403 */
static int LRSR(double x, double y, double phi, double *t, double *u, double *v) {
  double xi = x + sin(phi), eta = y - 1.0 - cos(phi), rho, theta; R(-eta, xi, &rho, &theta);
  if (rho >= 2.0) {
    double tt = theta, uu = 2.0 - rho, vv = M(tt + 0.5 * PI - phi);
    if (tt >= 0.0 && uu <= 0.0 && vv <= 0.0) { *t = tt; *u = uu; *v = vv; return 1; }
  }
  return 0;
}
/* rs_curve.py:406-419 */
static int LRSL(double x, double y, double phi, double *t, double *u, double *v) {
  double xi = x - sin(phi), eta = y - 1.0 + cos(phi), rho, theta; R(xi, eta, &rho, &theta);
  if (rho >= 2.0) {
    double r = sqrt(rho * rho - 4.0), uu = 2.0 - r, tt = M(theta + atan2(r, -2.0)), vv = M(phi - 0.5 * PI - tt);
    if (tt >= 0.0 && uu <= 0.0 && vv <= 0.0) { *t = tt; *u = uu; *v = vv; return 1; }
  }
  return 0;
}
/* rs_curve.py:494-510 */
static int LRSLR(double x, double y, double phi, double *t, double *u, double *v) {
  double xi = x + sin(phi), eta = y - 1.0 - cos(phi), rho, theta; R(xi, eta, &rho, &theta);
  if (rho >= 2.0) {
    double uu = 4.0 - sqrt(rho * rho - 4.0);
    if (uu <= 0.0) {
      double tt = M(atan2((4.0 - uu) * xi - 2.0 * eta, -2.0 * xi + (uu - 4.0) * eta)), vv = M(tt - phi);
      if (tt >= 0.0 && vv >= 0.0) { *t = tt; *u = uu; *v = vv; return 1; }
    }
  }
  return 0;
}

typedef int (*word_fn)(double, double, double, double *, double *, double *);
#define HP (0.5 * PI)

/* rs_curve.py:627-644 with SCS/CSC/CCC/CCCC/CCSC/CCSCC (:200-534), in the reference's order */
/* Type propagation (needed by py_sum): x, y are np.float64 whenever maxc is (always on the
 * planner path: 1/min_radius_turn is a numpy scalar) -> `xy_np`; phi = q1[2]-q0[2] is an
 * np.float64 for every node but the root, whose theta is a Python float -> `phi_np`.
 * math.* results are Python floats; M(theta) keeps theta's type.  Hence per word the
 * lengths are Python floats except: the segment derived from phi (last forward / first in
 * the "backwards" variants) and SLS's t (type of x) and u (type of phi). */
static void generate_path(const double q0[3], const double q1[3], double maxc, int xy_np, int phi_np, rs_set *S) {
  double dx = q1[0] - q0[0], dy = q1[1] - q0[1], dth = q1[2] - q0[2];
  double c = cos(q0[2]), s = sin(q0[2]);
  double x = (c * dx + s * dy) * maxc, y = (-s * dx + c * dy) * maxc, phi = dth;
  double t, u, v, l[5];
  S->n = 0; S->degenerate = 0;
  /* SCS (:200-210) */
  const unsigned m_sls = (xy_np ? 1u : 0u) | (phi_np ? 2u : 0u), P = phi_np ? 1u : 0u;
  if (SLS(x, y, phi, &t, &u, &v)) { l[0] = t; l[1] = u; l[2] = v; set_path(S, l, "SLS", m_sls); }
  if (SLS(x, -y, -phi, &t, &u, &v)) { l[0] = t; l[1] = u; l[2] = v; set_path(S, l, "SRS", m_sls); }
  /* the 4 reflections used everywhere: (x,y,phi,+) (-x,y,-phi,-) (x,-y,-phi,+) (-x,-y,phi,-) */
  const double sx[4] = {1, -1, 1, -1}, sy[4] = {1, 1, -1, -1}, sp[4] = {1, -1, -1, 1}, sg[4] = {1, -1, 1, -1};
  /* CSC (:232-265) */
  { word_fn f[2] = {LSL, LSR}; const char *nm[2][2] = {{"LSL", "RSR"}, {"LSR", "RSL"}};
    for (int k = 0; k < 2; ++k) for (int r = 0; r < 4; ++r)
      if (f[k](sx[r] * x, sy[r] * y, sp[r] * phi, &t, &u, &v)) { l[0] = sg[r] * t; l[1] = sg[r] * u; l[2] = sg[r] * v; set_path(S, l, nm[k][r >> 1], P << 2); } }
  /* CCC (:268-305) */
  for (int r = 0; r < 4; ++r)
    if (LRL(sx[r] * x, sy[r] * y, sp[r] * phi, &t, &u, &v)) { l[0] = sg[r] * t; l[1] = sg[r] * u; l[2] = sg[r] * v; set_path(S, l, (r >> 1) ? "RLR" : "LRL", P << 2); }
  double xb = x * cos(phi) + y * sin(phi), yb = x * sin(phi) - y * cos(phi);
  for (int r = 0; r < 4; ++r)
    if (LRL(sx[r] * xb, sy[r] * yb, sp[r] * phi, &t, &u, &v)) { l[0] = sg[r] * v; l[1] = sg[r] * u; l[2] = sg[r] * t; set_path(S, l, (r >> 1) ? "RLR" : "LRL", P); }
  /* CCCC (:355-388) */
  for (int r = 0; r < 4; ++r)
    if (LRLRn(sx[r] * x, sy[r] * y, sp[r] * phi, &t, &u, &v)) { l[0] = sg[r] * t; l[1] = sg[r] * u; l[2] = sg[r] * -u; l[3] = sg[r] * v; set_path(S, l, (r >> 1) ? "RLRL" : "LRLR", P << 3); }
  for (int r = 0; r < 4; ++r)
    if (LRLRp(sx[r] * x, sy[r] * y, sp[r] * phi, &t, &u, &v)) { l[0] = sg[r] * t; l[1] = sg[r] * u; l[2] = sg[r] * u; l[3] = sg[r] * v; set_path(S, l, (r >> 1) ? "RLRL" : "LRLR", P << 3); }
  /* CCSC (:422-491) */
  { const char *n1[2] = {"LRSL", "RLSR"}, *n2[2] = {"LRSR", "RLSL"};
    for (int r = 0; r < 4; ++r)
      if (LRSL(sx[r] * x, sy[r] * y, sp[r] * phi, &t, &u, &v)) { l[0] = sg[r] * t; l[1] = sg[r] * -HP; l[2] = sg[r] * u; l[3] = sg[r] * v; set_path(S, l, n1[r >> 1], P << 3); }
    for (int r = 0; r < 4; ++r)
      if (LRSR(sx[r] * x, sy[r] * y, sp[r] * phi, &t, &u, &v)) { l[0] = sg[r] * t; l[1] = sg[r] * -HP; l[2] = sg[r] * u; l[3] = sg[r] * v; set_path(S, l, n2[r >> 1], P << 3); }
    const char *n3[2] = {"LSRL", "RSLR"}, *n4[2] = {"RSRL", "LSLR"};
    xb = x * cos(phi) + y * sin(phi); yb = x * sin(phi) - y * cos(phi);
    for (int r = 0; r < 4; ++r)
      if (LRSL(sx[r] * xb, sy[r] * yb, sp[r] * phi, &t, &u, &v)) { l[0] = sg[r] * v; l[1] = sg[r] * u; l[2] = sg[r] * -HP; l[3] = sg[r] * t; set_path(S, l, n3[r >> 1], P); }
    for (int r = 0; r < 4; ++r)
      if (LRSR(sx[r] * xb, sy[r] * yb, sp[r] * phi, &t, &u, &v)) { l[0] = sg[r] * v; l[1] = sg[r] * u; l[2] = sg[r] * -HP; l[3] = sg[r] * t; set_path(S, l, n4[r >> 1], P); } }
  /* CCSCC (:513-534) */
  for (int r = 0; r < 4; ++r)
    if (LRSLR(sx[r] * x, sy[r] * y, sp[r] * phi, &t, &u, &v)) { l[0] = sg[r] * t; l[1] = sg[r] * -HP; l[2] = sg[r] * u; l[3] = sg[r] * -HP; l[4] = sg[r] * v; set_path(S, l, (r >> 1) ? "RLSRL" : "LRSLR", P << 4); }
}

typedef struct { int n; double x[AVP_MAX_RS_POINTS], y[AVP_MAX_RS_POINTS], yaw[AVP_MAX_RS_POINTS]; int dir[AVP_MAX_RS_POINTS]; } rs_course;

/* rs_curve.py:597-624 */
static void interpolate(int ind, double l, char m, double maxc, double ox, double oy, double oyaw, rs_course *C) {
  if (m == 'S') {
    C->x[ind] = ox + l / maxc * cos(oyaw); C->y[ind] = oy + l / maxc * sin(oyaw); C->yaw[ind] = oyaw;
  } else {
    double ldx = sin(l) / maxc, ldy = 0.0;
    if (m == 'L') ldy = (1.0 - cos(l)) / maxc; else if (m == 'R') ldy = (1.0 - cos(l)) / (-maxc);
    double gdx = cos(-oyaw) * ldx + sin(-oyaw) * ldy, gdy = -sin(-oyaw) * ldx + cos(-oyaw) * ldy;
    C->x[ind] = ox + gdx; C->y[ind] = oy + gdy;
  }
  if (m == 'L') C->yaw[ind] = oyaw + l; else if (m == 'R') C->yaw[ind] = oyaw - l;
  C->dir[ind] = (l > 0.0) ? 1 : -1;
}

/* rs_curve.py:537-594; returns -1 if the point buffer would overflow */
static int generate_local_course(double L, const double *lengths, const char *mode, int nseg, double maxc, double step, rs_course *C) {
  int point_num = (int)(L / step) + nseg + 3;
  if (point_num > AVP_MAX_RS_POINTS) return -1;
  for (int i = 0; i < point_num; ++i) { C->x[i] = 0.0; C->y[i] = 0.0; C->yaw[i] = 0.0; C->dir[i] = 0; }
  int ind = 1;
  C->dir[0] = (lengths[0] > 0.0) ? 1 : -1;
  double d = (lengths[0] > 0.0) ? step : -step, pd = d, ll = 0.0;
  for (int i = 0; i < nseg; ++i) {
    char m = mode[i]; double l = lengths[i];
    d = (l > 0.0) ? step : -step;
    double ox = C->x[ind], oy = C->y[ind], oyaw = C->yaw[ind];
    ind -= 1;
    if (i >= 1 && (lengths[i - 1] * lengths[i]) > 0) pd = -d - ll; else pd = d - ll;
    while (fabs(pd) <= fabs(l)) { ind += 1; if (ind >= point_num) return -1; interpolate(ind, pd, m, maxc, ox, oy, oyaw, C); pd += d; }
    ll = l - pd - d;
    ind += 1; if (ind >= point_num) return -1;
    interpolate(ind, l, m, maxc, ox, oy, oyaw, C);
  }
  int n = point_num;
  while (n > 0 && C->x[n - 1] == 0.0) --n;     /* :588-592 */
  C->n = n;
  return 0;
}

typedef struct { int ok; int degenerate; rs_word w; rs_course c; } rs_result;

/* rs_curve.py:99-134: selected word only (selection uses L / maxc, last <= wins) */
static void calc_optimal_path(const double q0[3], const double q1[3], double maxc, double step_size, int want_course, int xy_np, int phi_np, rs_result *out) {
  rs_set S; generate_path(q0, q1, maxc, xy_np, phi_np, &S);
  out->ok = 0; out->degenerate = S.degenerate;
  if (S.n == 0) return;
  int mini = 0; double minL = S.w[0].L / maxc;
  for (int i = 0; i < S.n; ++i) { double Li = S.w[i].L / maxc; if (Li <= minL) { minL = Li; mini = i; } }
  out->w = S.w[mini];
  if (want_course) {
    rs_course *C = &out->c;
    if (generate_local_course(out->w.L, out->w.len, out->w.ct, out->w.n, maxc, step_size * maxc, C)) { out->ok = 0; return; }
    double cm = cos(-q0[2]), sm = sin(-q0[2]);
    for (int i = 0; i < C->n; ++i) {
      double ix = C->x[i], iy = C->y[i];
      C->x[i] = cm * ix + sm * iy + q0[0]; C->y[i] = -sm * ix + cm * iy + q0[1];
      C->yaw[i] = pi_2_pi(C->yaw[i] + q0[2]);
    }
  }
  for (int i = 0; i < out->w.n; ++i) out->w.len[i] = out->w.len[i] / maxc;
  out->w.L = out->w.L / maxc;
  out->ok = 1;
}

/* all retained words (normalised lengths), for leaf parity tests */
int orc_rs_words(const double q0[3], const double q1[3], double maxc, int xy_np, int phi_np, int32_t *nseg, double *lengths, char *ctypes, double *L) {
  rs_set S; generate_path(q0, q1, maxc, xy_np, phi_np, &S);
  for (int i = 0; i < S.n; ++i) { nseg[i] = S.w[i].n; memcpy(lengths + 5 * i, S.w[i].len, sizeof(double) * 5); memset(ctypes + 8 * i, 0, 8); strcpy(ctypes + 8 * i, S.w[i].ct); L[i] = S.w[i].L; }
  return S.degenerate ? -1 - S.n : S.n;
}
int orc_rs_optimal(const double q0[3], const double q1[3], double maxc, double step_size, int xy_np, int phi_np, int32_t *nseg, double *lengths, char *ctypes, double *L,
                   int cap, double *x, double *y, double *yaw, int32_t *dir, int32_t *n_pts) {
  rs_result *rp = malloc(sizeof(rs_result)); calc_optimal_path(q0, q1, maxc, step_size, 1, xy_np, phi_np, rp);
#define r (*rp)
  if (!r.ok) { int rc_ = r.degenerate ? -2 : -1; free(rp); return rc_; }
  *nseg = r.w.n; memcpy(lengths, r.w.len, sizeof(double) * r.w.n); memset(ctypes, 0, 8); strcpy(ctypes, r.w.ct); *L = r.w.L;
  *n_pts = r.c.n;
  for (int i = 0; i < r.c.n && i < cap; ++i) { x[i] = r.c.x[i]; y[i] = r.c.y[i]; yaw[i] = r.c.yaw[i]; dir[i] = r.c.dir[i]; }
  { int rc_ = r.degenerate ? 1 : 0; free(rp); return rc_; }
#undef r
}

/* ------------------------------------------------------------------ Dijkstra (path_plan/compute_h.py) */

typedef struct { int32_t dist; int32_t id; double x, y; } hgrid;
typedef struct orc_dij {
  const orc_map *m;
  hgrid *heap; int hn, hcap;       /* queue.PriorityQueue == heapq on a list */
  long n_ids;
  uint8_t *seen;                   /* openlist_index membership (compute_h.py:220,235) */
  uint8_t *inheap;
  int32_t *hval;                   /* distance of the first closedlist entry per id, -1 none */
  long closed_len;                 /* len(closedlist) */
  long terminate_id; int find_terminate;
  int unreachable;
} orc_dij;

static int grid_lt(const hgrid *a, const hgrid *b) { return (a->dist == b->dist) ? (a->id < b->id) : (a->dist < b->dist); } /* compute_h.py:33-38 */

/* CPython heapq._siftdown / _siftup / heappush / heappop (Lib/heapq.py:207-278) */
static void hq_siftdown(hgrid *h, int start, int pos) {
  hgrid item = h[pos];
  while (pos > start) { int parent = (pos - 1) >> 1; if (grid_lt(&item, &h[parent])) { h[pos] = h[parent]; pos = parent; continue; } break; }
  h[pos] = item;
}
static void hq_siftup(hgrid *h, int n, int pos) {
  int start = pos; hgrid item = h[pos]; int child = 2 * pos + 1;
  while (child < n) { int right = child + 1; if (right < n && !grid_lt(&h[child], &h[right])) child = right; h[pos] = h[child]; pos = child; child = 2 * pos + 1; }
  h[pos] = item; hq_siftdown(h, start, pos);
}

orc_dij *orc_dij_new(const orc_map *m) {
  orc_dij *d = calloc(1, sizeof(*d)); d->m = m;
  long stride = (long)(int)((m->boundary[1] - m->boundary[0]) / m->dx);
  long W = (long)floor((m->boundary[1] - m->boundary[0]) / m->dx) + 2, H = (long)floor((m->boundary[3] - m->boundary[2]) / m->dy) + 2;
  d->n_ids = H * stride + W + 8;
  d->seen = calloc(d->n_ids, 1); d->inheap = calloc(d->n_ids, 1);
  d->hval = malloc(sizeof(int32_t) * d->n_ids); for (long i = 0; i < d->n_ids; ++i) d->hval[i] = -1;
  d->hcap = 1 << 16; d->heap = malloc(sizeof(hgrid) * d->hcap);
  return d;
}
void orc_dij_free(orc_dij *d) { if (!d) return; free(d->seen); free(d->inheap); free(d->hval); free(d->heap); free(d); }

/* compute_h.py:237-255 */
static int is_obstacle(const orc_map *m, double gx, double gy) {
  long xi = (long)floor((gx - m->boundary[0]) / m->dx) - 1, yi = (long)floor((gy - m->boundary[2]) / m->dy) - 1;
  long mx = (long)(int)((m->boundary[1] - m->boundary[0]) / m->dx), my = (long)(int)((m->boundary[3] - m->boundary[2]) / m->dy);
  if (xi >= mx) xi = mx - 1; if (yi >= my) yi = my - 1;
  if (xi < 0) xi += m->nx; if (yi < 0) yi += m->ny;      /* python negative indexing */
  if (xi < 0 || xi >= m->nx || yi < 0 || yi >= m->ny) return 0; /* reference: IndexError; unreachable inside the boundary */
  return m->cost[xi * m->ny + yi] == 255;
}

/* compute_h.py:216-235 */
static void add_grid_to_openlist(orc_dij *d, double gx, double gy, int priority) {
  long id = map_index(d->m, gx, gy);
  if (id < 0 || id >= d->n_ids) return;
  if (d->seen[id]) {
    if (d->inheap[id]) for (int i = 0; i < d->hn; ++i) if (d->heap[i].id == id) { if (d->heap[i].dist > priority) d->heap[i].dist = priority; break; } /* in place, no re-sift */
  } else {
    if (d->hn == d->hcap) { d->hcap *= 2; d->heap = realloc(d->heap, sizeof(hgrid) * d->hcap); }
    hgrid g = {priority, (int32_t)id, gx, gy};
    d->heap[d->hn++] = g; hq_siftdown(d->heap, 0, d->hn - 1);
    d->seen[id] = 1; d->inheap[id] = 1;
  }
}

/* compute_h.py:84-195 */
static void update_openlist(orc_dij *d, const hgrid *cur) {
  const orc_map *m = d->m; const double *b = m->boundary;
  static const int ddx[8] = {-1, 0, 1, -1, 1, -1, 0, 1}, ddy[8] = {1, 1, 1, 0, 0, -1, -1, -1};
  for (int i = 0; i < 8; ++i) {
    double gx = ddx[i] < 0 ? cur->x - m->dx : (ddx[i] > 0 ? cur->x + m->dx : cur->x);
    double gy = ddy[i] < 0 ? cur->y - m->dy : (ddy[i] > 0 ? cur->y + m->dy : cur->y);
    if (is_obstacle(m, gx, gy)) continue;
    int ok = 1;
    if (ddx[i] < 0 && !(gx >= b[0])) ok = 0; if (ddx[i] > 0 && !(gx <= b[1])) ok = 0;
    if (ddy[i] > 0 && !(gy <= b[3])) ok = 0; if (ddy[i] < 0 && !(gy >= b[2])) ok = 0;
    if (ok) add_grid_to_openlist(d, gx, gy, cur->dist + ((ddx[i] && ddy[i]) ? 14 : 10));
  }
}

/* compute_h.py:198-214 (+ initial_map :50-72, update_closedlist :74-82).  Returns the
 * distance, or -1 when the queue runs dry (the reference would block forever). */
int orc_dij_compute_path(orc_dij *d, double node_x, double node_y) {
  const orc_map *m = d->m;
  d->find_terminate = 0;
  hgrid cur = {0, 0, m->pose[3], m->pose[4]};
  long gid = map_index(m, cur.x, cur.y); cur.id = (int32_t)gid;
  d->closed_len++; if (gid >= 0 && gid < d->n_ids && d->hval[gid] < 0) d->hval[gid] = 0;
  d->terminate_id = map_index(m, node_x, node_y);
  while (!d->find_terminate) {
    update_openlist(d, &cur);
    if (d->hn == 0) { d->unreachable = 1; return -1; }
    hgrid last = d->heap[--d->hn];                      /* heappop */
    if (d->hn) { cur = d->heap[0]; d->heap[0] = last; hq_siftup(d->heap, d->hn, 0); } else cur = last;
    d->inheap[cur.id] = 0;
    if (cur.id == d->terminate_id) d->find_terminate = 1;
    d->closed_len++; if (d->hval[cur.id] < 0) d->hval[cur.id] = cur.dist;
  }
  return cur.dist;
}
long orc_dij_closed_len(const orc_dij *d) { return d->closed_len; }
long orc_dij_n_ids(const orc_dij *d) { return d->n_ids; }
void orc_dij_hvalues(const orc_dij *d, int32_t *out) { memcpy(out, d->hval, sizeof(int32_t) * d->n_ids); }

/* ------------------------------------------------------------------ hybrid A* (path_plan/hybrid_a_star.py, path_planner.py) */

typedef struct {
  double x, y, theta, f, g, h, steer;
  int32_t index, parent; uint8_t forward, in_open, in_closed, used;
} onode;

typedef struct {
  const orc_map *m; const avp_config *cfg; orc_dij *dij;
  onode *nodes; int ncap;           /* indexed by node index */
  int32_t *heap; int hn, hcap;      /* open_list.queue: node indices, ordered by f (hybrid_a_star.py:61-68) */
  int32_t *htab; int hmask;         /* exact pose -> node index (replaces the O(n) == scans :155-172) */
  int n_closed, global_index;
  double goal[3];
  int n_hq, n_hcalls;
  void *rsbuf;                      /* scratch rs_result */
  int64_t *hq_log; int hq_cap;      /* (terminate id, dist, closed_len) per compute_path call */
  int status;
} astar;

static uint64_t pose_hash(double x, double y, double t) {
  uint64_t a, b, c; x += 0.0; y += 0.0; t += 0.0; memcpy(&a, &x, 8); memcpy(&b, &y, 8); memcpy(&c, &t, 8);
  uint64_t h = a * 0x9E3779B97F4A7C15ULL; h ^= (h >> 29); h += b * 0xBF58476D1CE4E5B9ULL; h ^= (h >> 31); h += c * 0x94D049BB133111EBULL; h ^= (h >> 30);
  return h * 0xD6E8FEB86659FD93ULL;
}
static int htab_find(const astar *A, double x, double y, double t) {
  uint64_t p = pose_hash(x, y, t) >> 20;
  for (;; ++p) { int32_t e = A->htab[p & A->hmask]; if (e < 0) return -1; const onode *n = &A->nodes[e]; if (n->x == x && n->y == y && n->theta == t) return e; }
}
static void htab_insert(astar *A, int idx) {
  const onode *n = &A->nodes[idx]; uint64_t p = pose_hash(n->x, n->y, n->theta) >> 20;
  for (;; ++p) if (A->htab[p & A->hmask] < 0) { A->htab[p & A->hmask] = idx; return; }
}
static int node_lt(const astar *A, int a, int b) { return A->nodes[a].f < A->nodes[b].f; }
static void open_siftdown(astar *A, int start, int pos) {
  int32_t item = A->heap[pos];
  while (pos > start) { int parent = (pos - 1) >> 1; if (node_lt(A, item, A->heap[parent])) { A->heap[pos] = A->heap[parent]; pos = parent; continue; } break; }
  A->heap[pos] = item;
}
static void open_siftup(astar *A, int pos) {
  int n = A->hn, start = pos; int32_t item = A->heap[pos]; int child = 2 * pos + 1;
  while (child < n) { int right = child + 1; if (right < n && !node_lt(A, A->heap[child], A->heap[right])) child = right; A->heap[pos] = A->heap[child]; pos = child; child = 2 * pos + 1; }
  A->heap[pos] = item; open_siftdown(A, start, pos);
}
static void open_put(astar *A, int idx) { A->heap[A->hn++] = idx; open_siftdown(A, 0, A->hn - 1); }
static int open_get(astar *A) { int32_t last = A->heap[--A->hn], ret = last; if (A->hn) { ret = A->heap[0]; A->heap[0] = last; open_siftup(A, 0); } return ret; }

static int dij_query(astar *A, double x, double y) {
  int d = orc_dij_compute_path(A->dij, x, y);
  if (A->n_hq < A->hq_cap) { A->hq_log[3 * A->n_hq] = A->dij->terminate_id; A->hq_log[3 * A->n_hq + 1] = d; A->hq_log[3 * A->n_hq + 2] = A->dij->closed_len; }
  A->n_hq++;
  if (d < 0) A->status = AVP_H_UNREACHABLE;
  return d;
}

/* hybrid_a_star.py:243-259 */
static double calc_node_cost(const astar *A, const onode *n, double father_theta, int father_gear) {
  double cost_gear = 0; if ((int)n->forward != father_gear) cost_gear = A->cfg->cost_gear;
  double cost_heading = fabs(n->theta - father_theta);
  double cost = cost_gear + A->cfg->cost_heading_change * cost_heading;
  return A->cfg->cost_scale * cost;
}
/* hybrid_a_star.py:261-298 */
static double calc_node_heuristic(astar *A, const onode *n) {
  A->n_hcalls++;
  long id = map_index(A->m, n->x, n->y);
  int h1;
  if (id >= 0 && id < A->dij->n_ids && A->dij->hval[id] >= 0) h1 = A->dij->hval[id];
  else { h1 = dij_query(A, n->x, n->y); if (h1 < 0) return 0.0; }
  double q0[3] = {n->x, n->y, n->theta};
  rs_result *rp = A->rsbuf; calc_optimal_path(q0, A->goal, 1 / A->cfg->min_radius_turn, 0.5, 0, 1, n->index != 0, rp);
  if (rp->degenerate || !rp->ok) { A->status = AVP_RS_DEGENERATE; return 0.0; }
  double h2 = rp->w.L, hv1 = h1 / 100.0;
  return (h2 > hv1) ? h2 : hv1;   /* max(h_value_1, h_value_2) */
}

/* hybrid_a_star.py:126-241 */
static void expand_node(astar *A, int cur_idx) {
  const avp_config *c = A->cfg; const double *b = A->m->boundary;
  onode cur = A->nodes[cur_idx];
  int ns = c->steering_angle_num, next_index = 2 * ns;
  for (int i = 0; i < next_index; ++i) {
    double steer = c->steer[i % ns], tn = c->tan_steer[i % ns];
    int fwd = (i < next_index / 2.0);
    double speed = fwd ? c->max_v : -c->max_v;
    double td = speed * c->dt;
    double th = cur.theta + (c->max_v * tn) / c->lw * c->dt;
    th = pi_2_pi(th);
    double x_ = cur.x + td * cos(th), y_ = cur.y + td * sin(th);
    int found = htab_find(A, x_, y_, th);
    int in_closed = (found >= 0 && A->nodes[found].in_closed);
    int oob = (A->n_closed > 0) && (x_ > b[1] || x_ < b[0] || y_ > b[3] || y_ < b[2]);   /* :155-163 */
    if (in_closed || oob) continue;
    int child;
    if (found < 0) {                                                                   /* :175-216 */
      child = A->global_index + i + 1;
      if (child >= A->ncap) { A->status = AVP_CAPACITY; return; }
      onode *n = &A->nodes[child]; memset(n, 0, sizeof(*n));
      n->x = x_; n->y = y_; n->theta = th; n->index = child; n->parent = cur.index; n->forward = (uint8_t)fwd; n->steer = steer; n->used = 1;
      int collision = 0;
      for (int s = 0; s < c->n_substeps; ++s) {
        double td_i = speed * c->ddt * (s + 1);
        double th_i = cur.theta + (c->max_v * tn) / c->lw * c->ddt * (s + 1);
        th_i = pi_2_pi(th_i);
        double x_i = cur.x + td_i * cos(th_i), y_i = cur.y + td_i * sin(th_i);
        collision = orc_check(A->m, c, x_i, y_i, th_i);
        if (collision) { n->in_closed = 1; A->n_closed++; break; }
      }
      htab_insert(A, child);
      if (!collision) {
        n->g = calc_node_cost(A, n, cur.theta, cur.forward);
        n->h = calc_node_heuristic(A, n); if (A->status) return;
        n->f = n->g + n->h;
        open_put(A, child); n->in_open = 1;
      }
    } else {                                                                           /* :219-230 */
      child = found; onode *n = &A->nodes[child];
      double new_h = calc_node_heuristic(A, n); if (A->status) return;
      double new_g = calc_node_cost(A, n, cur.theta, cur.forward);
      double new_f = new_h + new_g;
      if (new_f < n->f) { n->f = new_f; n->g = new_g; n->h = new_h; n->parent = cur.index; n->forward = (uint8_t)fwd; n->steer = steer; }
    }
  }
  onode *cn = &A->nodes[cur_idx]; cn->in_closed = 1; cn->in_open = 0; A->n_closed++;   /* :235-237 */
  A->global_index += next_index;
}

/* caller-visible result of one plan */
typedef struct orc_plan_out {
  avp_plan_summary sum;
  int32_t *pops; int cap_pops;
  double *pop_state;     /* n_pops*3 or NULL */
  double *pop_fgh;       /* n_pops*3 or NULL */
  double *final_path; int cap_path;    /* rows x,y,theta */
  double *rs_x, *rs_y, *rs_yaw; int32_t *rs_dir; int cap_rs;
  int64_t *hq_log; int cap_hq;
  int32_t *hval_out; long hval_cap;    /* final h table (first closedlist distance per grid id), may be NULL */
} orc_plan_out;

/* PathPlanner.a_star_plan (path_planner.py:58-110) incl. hybrid_a_star.__init__ (:72-124),
 * try_reach_goal/try_rs_curve (:300-349), finish_path (:351-389). */
int orc_plan(const orc_map *m, const avp_config *cfg, orc_plan_out *out) {
  astar A; memset(&A, 0, sizeof(A));
  A.m = m; A.cfg = cfg; A.dij = orc_dij_new(m);
  int max_pops = cfg->max_pops > 0 ? cfg->max_pops : 100000;
  A.ncap = 2 * cfg->steering_angle_num * (max_pops + 1) + 2;
  A.nodes = calloc(A.ncap, sizeof(onode));
  A.hcap = A.ncap; A.heap = malloc(sizeof(int32_t) * A.hcap);
  int hb = 1; while (hb < 2 * A.ncap) hb <<= 1; A.hmask = hb - 1; A.htab = malloc(sizeof(int32_t) * hb); memset(A.htab, 0xff, sizeof(int32_t) * hb);
  A.hq_log = out->hq_log; A.hq_cap = out->hq_log ? out->cap_hq : 0;
  A.rsbuf = malloc(sizeof(rs_result));
  rs_result *rsp = calloc(1, sizeof(rs_result));
#define rs (*rsp)
  avp_plan_summary *S = &out->sum; memset(S, 0, sizeof(*S));
  S->nx = m->nx; S->ny = m->ny; S->n_obs = m->n_obs; S->pitch[0] = m->dx; S->pitch[1] = m->dy;
  memcpy(S->boundary, m->boundary, sizeof(double) * 4); S->origin[0] = m->boundary[0]; S->origin[1] = m->boundary[2];
  if (m->raster_error) { A.status = AVP_RASTER_AMBIGUOUS; goto done; }

  dij_query(&A, m->pose[0], m->pose[1]);                       /* hybrid_a_star.py:89-91 */
  if (A.status) goto done;
  onode *n0 = &A.nodes[0]; n0->x = m->pose[0]; n0->y = m->pose[1]; n0->theta = pi_2_pi(m->pose[2]); n0->index = 0; n0->parent = -1; n0->forward = 1; n0->used = 1; n0->in_open = 1;
  A.goal[0] = m->pose[3]; A.goal[1] = m->pose[4]; A.goal[2] = pi_2_pi(m->pose[5]);
  htab_insert(&A, 0); open_put(&A, 0);

  int reach_goal = 0, cur = -1, in_radius = 0, collision = 0, n_pops = 0;
  double maxc = 1 / cfg->min_radius_turn;
  while (A.hn > 0 && !reach_goal) {                              /* path_planner.py:68 */
    if (n_pops >= max_pops) { A.status = AVP_CAPACITY; break; }
    cur = open_get(&A);
    const onode *cn = &A.nodes[cur];
    if (out->pops && n_pops < out->cap_pops) {
      out->pops[n_pops] = cn->index;
      if (out->pop_state) { out->pop_state[3 * n_pops] = cn->x; out->pop_state[3 * n_pops + 1] = cn->y; out->pop_state[3 * n_pops + 2] = cn->theta; }
      if (out->pop_fgh) { out->pop_fgh[3 * n_pops] = cn->f; out->pop_fgh[3 * n_pops + 1] = cn->g; out->pop_fgh[3 * n_pops + 2] = cn->h; }
    }
    n_pops++;
    /* try_reach_goal (hybrid_a_star.py:300-316) */
    collision = 0; in_radius = 0; rs.ok = 0;
    double distance = sqrt(p_pow(cn->x - A.goal[0], 2.0) + p_pow(cn->y - A.goal[1], 2.0));
    if (distance < cfg->flag_radius) {
      in_radius = 1;
      double q0[3] = {cn->x, cn->y, cn->theta};
      calc_optimal_path(q0, A.goal, maxc, 0.5, 1, 1, cn->index != 0, &rs);            /* try_rs_curve (:318-349) */
      if (rs.degenerate || !rs.ok) { A.status = AVP_RS_DEGENERATE; break; }
      for (int i = 0; i < rs.c.n; ++i) {
        collision = orc_check(m, cfg, rs.c.x[i], rs.c.y[i], pi_2_pi(rs.c.yaw[i]));
        if (collision) break;
      }
    }
    if (!collision && in_radius) { reach_goal = 1; break; }
    expand_node(&A, cur);
    if (A.status) break;
  }
  S->n_pops = n_pops; S->last_index = cur >= 0 ? A.nodes[cur].index : -1;
  if (!A.status && !reach_goal) A.status = (in_radius && rs.ok) ? AVP_OPEN_EXHAUSTED_RS : AVP_OPEN_EXHAUSTED;
  if (A.status == AVP_OK || A.status == AVP_OPEN_EXHAUSTED_RS) {
    /* finish_path (hybrid_a_star.py:351-389) */
    int chain[4096], nc = 0, k = cur;
    while (A.nodes[k].index != 0 && nc < 4095) { chain[nc++] = k; k = A.nodes[k].parent; }
    chain[nc++] = k;
    int np_ = 0; double *fp = out->final_path; int cap = out->cap_path;
#define PUSH(px, py, pt) do { if (fp && np_ < cap) { fp[3 * np_] = (px); fp[3 * np_ + 1] = (py); fp[3 * np_ + 2] = (pt); } np_++; } while (0)
    PUSH(A.nodes[k].x, A.nodes[k].y, A.nodes[k].theta);
    for (int i = 0; i < nc; ++i) {
      int kk = nc - 1 - i; if (kk == 0) break;
      const onode *par = &A.nodes[chain[kk]], *ch = &A.nodes[chain[kk - 1]];
      for (int j = 0; j < cfg->n_substeps; ++j) {
        double speed = ch->forward ? cfg->max_v : -cfg->max_v;
        double td_j = speed * cfg->ddt * (j + 1);
        double tanv = 0; for (int q = 0; q < cfg->steering_angle_num; ++q) if (cfg->steer[q] == ch->steer) tanv = cfg->tan_steer[q];
        double th_j = par->theta + (cfg->max_v * tanv) / cfg->lw * cfg->ddt * (j + 1);
        th_j = pi_2_pi(th_j);
        PUSH(par->x + td_j * cos(th_j), par->y + td_j * sin(th_j), th_j);
      }
    }
    S->n_astar = np_;
    for (int i = 1; i < rs.c.n; ++i) PUSH(rs.c.x[i], rs.c.y[i], rs.c.yaw[i]);     /* path_planner.py:104-108 */
    S->n_final = np_; S->n_rs = rs.c.n; S->rs_nseg = rs.w.n; S->rs_L = rs.w.L;
    memcpy(S->rs_lengths, rs.w.len, sizeof(double) * rs.w.n); strcpy(S->rs_ctypes, rs.w.ct);
    for (int i = 0; i < rs.c.n && i < out->cap_rs; ++i) {
      if (out->rs_x) { out->rs_x[i] = rs.c.x[i]; out->rs_y[i] = rs.c.y[i]; out->rs_yaw[i] = rs.c.yaw[i]; }
      if (out->rs_dir) out->rs_dir[i] = rs.c.dir[i];
    }
  }
done:
  S->status = A.status; S->global_index = A.global_index; S->n_closed = A.n_closed; S->n_open = A.hn;
  S->n_hq = A.n_hq; S->h_closed = (int32_t)A.dij->closed_len; S->n_hcalls = A.n_hcalls;
  if (S->last_index >= 0 && S->n_pops > 0) { const onode *ln = &A.nodes[S->last_index]; S->last_pose[0] = ln->x; S->last_pose[1] = ln->y; S->last_pose[2] = ln->theta; }
  if (out->hval_out) for (long i = 0; i < A.dij->n_ids && i < out->hval_cap; ++i) out->hval_out[i] = A.dij->hval[i];
  orc_dij_free(A.dij); free(A.nodes); free(A.heap); free(A.htab); free(A.rsbuf); free(rsp);
  return 0;
#undef rs
}

/* rollout + collision flags + rs length of the 2n successors of one pose: the pure part of
 * expand_node, for kernel-level parity tests (mirrors avp_expand_pure) */
void orc_expand_pure(const orc_map *m, const avp_config *c, const double parent[3], double *out_pose, int32_t *out_flags, double *out_rsL) {
  const double *b = m->boundary; int ns = c->steering_angle_num;
  double goal[3] = {m->pose[3], m->pose[4], pi_2_pi(m->pose[5])};
  rs_result *rp = malloc(sizeof(rs_result));
  for (int i = 0; i < 2 * ns; ++i) {
    double tn = c->tan_steer[i % ns]; int fwd = i < ns; double speed = fwd ? c->max_v : -c->max_v;
    double th = pi_2_pi(parent[2] + (c->max_v * tn) / c->lw * c->dt);
    double x_ = parent[0] + speed * c->dt * cos(th), y_ = parent[1] + speed * c->dt * sin(th);
    out_pose[3 * i] = x_; out_pose[3 * i + 1] = y_; out_pose[3 * i + 2] = th;
    int fl = 0;
    for (int s = 0; s < c->n_substeps; ++s) {
      double th_i = pi_2_pi(parent[2] + (c->max_v * tn) / c->lw * c->ddt * (s + 1));
      double td_i = speed * c->ddt * (s + 1);
      if (orc_check(m, c, parent[0] + td_i * cos(th_i), parent[1] + td_i * sin(th_i), th_i)) { fl |= 1; break; }
    }
    if (x_ > b[1] || x_ < b[0] || y_ > b[3] || y_ < b[2]) fl |= 2;
    out_flags[i] = fl;
    double q0[3] = {x_, y_, th}; calc_optimal_path(q0, goal, 1 / c->min_radius_turn, 0.5, 0, 1, 1, rp);
    out_rsL[i] = rp->ok ? rp->w.L : NAN;
  }
  free(rp);
}

double orc_py_hypot(double a, double b) { return py_hypot(a, b); }
double orc_py_sum(const double *v, int n, unsigned npmask) { return py_sum(v, n, npmask); }

/* ------------------------------------------------------------------ split_path (path_plan/path_planner.py:112-192)
 * Sequential restatement.  compute_cosin = 1 - scipy.spatial.distance.cosine(v1, v2) (:126-133): scipy 1.18
 * correlation(centered=False) is  dist = 1.0 - uv / math.sqrt(uu * vv)  with uv, uu, vv = np.dot (2-element ddot of the
 * installed BLAS: fma(a1, b1, a0 * b0), pinned by tests/golden/leaf_split.npz), then np.clip(dist, 0, 2); a zero
 * displacement gives 0/0 = NaN and NaN < 0 is False.  Segments are written back to back into out (rows x, y, theta),
 * their lengths into seg_len; info = {status, n segments, change_gear, n points}; status 0 ok, 1 = IndexError at :181
 * (no gear change), 3 = capacity. */
void orc_split_path(const orc_map *m, const avp_config *cfg, int nf, const double *p, double *out, int cap_pts,
                    int32_t *seg_len, int cap_seg, int32_t *info) {
  int start = 0, change = 0, have = 0, nseg = 0, npts = 0, over = 0;
  double ext[64][3];
  int next = cfg->extended_num < 64 ? cfg->extended_num : 64;
#define SP_PUT(at, x, y, t) do { if ((at) < cap_pts) { out[3 * (at)] = (x); out[3 * (at) + 1] = (y); out[3 * (at) + 2] = (t); } else over = 1; } while (0)
  for (int i = 0; i < nf - 2; ++i) {
    double u0 = p[3 * (i + 1)] - p[3 * i], u1 = p[3 * (i + 1) + 1] - p[3 * i + 1];                   /* :127-128 */
    double v0 = p[3 * (i + 2)] - p[3 * (i + 1)], v1 = p[3 * (i + 2) + 1] - p[3 * (i + 1) + 1];       /* :130-131 */
    double uv = fma(u1, v1, u0 * v0), uu = fma(u1, u1, u0 * u0), vv = fma(v1, v1, v0 * v0);
    double dist = 1.0 - uv / sqrt(uu * vv);
    if (dist < 0.0) dist = 0.0; else if (dist > 2.0) dist = 2.0;
    double compute_cosin = 1 - dist;
    if (!(compute_cosin < 0)) continue;                                                              /* :136 */
    change++;
    int end = i + 2, len = 0;
    if (change > 1 && have > 0) {                                                                    /* :141-149 */
      for (int j = 0; j < have; ++j) SP_PUT(npts + j, ext[have - 1 - j][0], ext[have - 1 - j][1], ext[have - 1 - j][2]);
      len = have; have = 0;
    }
    for (int r = start; r < end; ++r, ++len) SP_PUT(npts + len, p[3 * r], p[3 * r + 1], p[3 * r + 2]);   /* :139 */
    for (int j = 0; j < next; ++j) {                                                                 /* :152-174 */
      double xi = p[3 * i], thi = p[3 * i + 2], xn = p[3 * (i + 1)], yn = p[3 * (i + 1) + 1], thn = p[3 * (i + 1) + 2];
      int f1 = (xn > xi) && (thi > -PI / 2 && thi < PI / 2);
      int f2 = (xn < xi) && ((thi > PI / 2 && thi < PI) || (thi > -PI && thi < -PI / 2));
      double speed = (f1 || f2) ? cfg->max_v : -cfg->max_v;
      double td = speed * cfg->ddt * (j + 1);
      double xj = xn + td * cos(thn), yj = yn + td * sin(thn);
      if (!orc_check(m, cfg, xj, yj, thn)) {
        SP_PUT(npts + len, xj, yj, thn);
        ext[have][0] = xj; ext[have][1] = yj; ext[have][2] = thn; ++have; ++len;
      }
    }
    if (nseg < cap_seg) seg_len[nseg] = len; else over = 1;
    ++nseg; npts += len; start = i + 1;                                                              /* :176-177 */
  }
  int status = 0;
  if (nseg == 0) status = 1;                                                                         /* :181 split_path[-1] on [] */
  else {
    int len = 0;
    if (have > 0) { for (int j = 0; j < have; ++j) SP_PUT(npts + j, ext[have - 1 - j][0], ext[have - 1 - j][1], ext[have - 1 - j][2]); len = have; }   /* :183-188 */
    for (int r = start; r < nf; ++r, ++len) SP_PUT(npts + len, p[3 * r], p[3 * r + 1], p[3 * r + 2]);   /* :180 */
    if (nseg < cap_seg) seg_len[nseg] = len; else over = 1;
    ++nseg; npts += len;
  }
#undef SP_PUT
  if (over) status = 3;
  info[0] = status; info[1] = nseg; info[2] = change; info[3] = npts;
}
