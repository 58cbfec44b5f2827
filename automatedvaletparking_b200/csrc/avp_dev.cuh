// avp_dev.cuh -- device-side building blocks of the hybrid-A* hot path (sm_100a).
//
// Everything here is fp64 IEEE arithmetic evaluated in the reference's operation order;
// the file is compiled with -fmad=false and every fused multiply-add is explicit.
// Citations are to the reference (wenqing-2021/AutomatedValetParking) file:line.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include "../../include/avp_b200.h"
#include "avp_libm.h"

#define AVP_PI 3.141592653589793  /* math.pi (rs_curve.py:25) */
#define AVP_HALF_PI (0.5 * AVP_PI)
#define AVP_FULL_MASK 0xffffffffu
#define AVP_MAX_VERT 64           /* vertices per obstacle polygon handled by the rasteriser */

// ------------------------------------------------------------------------------------------
// per-scenario device record
struct ScenDev {
  double pose[6];        // x0,y0,theta0,xf,yf,thetaf
  double b[4];           // boundary (costmap.py:169-172)
  double dx, dy;         // _discrete_x/_y (costmap.py:190-191)
  double stepx, stepy;   // np.linspace step (b1-b0)/(nx-1)
  double inv_stepx;      // 1 / stepx: first guess of a column index (col_range), never decides anything
  int32_t nx, ny;        // cost_map.shape
  int32_t stride;        // int((b1-b0)/dx): row stride of convert_position_to_index (costmap.py:327-328)
  int32_t mx, my;        // is_obstacle clamp (compute_h.py:243-246)
  int32_t n_ids;         // size of the per-id arrays
  int32_t n_obs;         // count(cost_map == 255)
  int32_t raster_error;
  int64_t cost_off;      // into cost maps (bytes)
  int64_t cell_off;      // into obstacle cell list (double2 entries)
  int64_t col_off;       // into column starts (int32 entries, nx+1 per scenario)
  int64_t id_off;        // into per-id arrays
  int32_t cell_cap;      // capacity of the cell list
  int32_t obs_begin, obs_end;  // polygons of this scenario
};

// ------------------------------------------------------------------------------------------
// small math helpers

// Out-of-line device entry points of the bit-exact libm restatements (avp_sincos.h, avp_libm.h):
// one copy of each body per kernel keeps code size and register pressure down.
__device__ __noinline__ double d_sin(double x) { return avp_sin(x); }
__device__ __noinline__ double d_cos(double x) { return avp_cos(x); }
// sine AND cosine of one argument in one function: the same two restatements, inlined side by side so that the compiler interleaves
// their (independent) dependency chains -- the bits are those of d_sin / d_cos, the latency is that of one of them plus a little
// (the evaluation is bound by fixed-latency fp64 dependencies, profiles/ncu_kplan_r02.csv: "wait" is its largest stall)
#ifdef AVP_NO_SINCOS     // A/B build: two calls
__device__ __forceinline__ void d_sincos(double x, double &s, double &c) { s = d_sin(x); c = d_cos(x); }
#else
// (the pair comes back in registers: reference parameters of an out-of-line function live in local memory)
__device__ __noinline__ double2 d_sincos_v(double x) { double2 r; r.x = avp_sin(x); r.y = avp_cos(x); return r; }
__device__ __forceinline__ void d_sincos(double x, double &s, double &c) { const double2 r = d_sincos_v(x); s = r.x; c = r.y; }
#endif
__device__ __noinline__ double d_atan2(double y, double x) { return avp_atan2(y, x); }
__device__ __noinline__ double d_asin(double x) { return avp_asin(x); }
__device__ __noinline__ double d_acos(double x) { return avp_acos(x); }
__device__ __noinline__ double d_tan(double x) { return avp_tan(x); }
__device__ __noinline__ double d_pow2(double x) { return avp_pow2(x); }   // libm pow(x, 2.0)

__device__ __forceinline__ double d_add(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double d_mul(double a, double b) { return __dmul_rn(a, b); }

// math.hypot of CPython 3.12 (Modules/mathmodule.c vector_norm, n = 2) -- used by rs_curve.R
// (rs_curve.py:659-666).  Scaling by a power of two is exact; Dekker products; compensated sum.
struct dl_t { double hi, lo; };
__device__ __forceinline__ dl_t dl_fast_sum(double a, double b) { dl_t r; r.hi = a + b; r.lo = (a - r.hi) + b; return r; }
__device__ __forceinline__ dl_t dl_split(double x) { double t = x * 134217729.0; dl_t r; r.hi = t - (t - x); r.lo = x - r.hi; return r; }
__device__ __forceinline__ dl_t dl_mul(double x, double y) {
  dl_t xx = dl_split(x), yy = dl_split(y);
  double p = xx.hi * yy.hi, q = xx.hi * yy.lo + xx.lo * yy.hi;
  dl_t r; r.hi = p + q; r.lo = p - r.hi + q + xx.lo * yy.lo; return r;
}
__device__ __noinline__ double py_hypot(double a, double b) {
  double v0 = fabs(a), v1 = fabs(b);
  double mx = v0 > v1 ? v0 : v1;
  if (isinf(v0) || isinf(v1)) return INFINITY;
  if (isnan(v0) || isnan(v1)) return NAN;
  if (mx == 0.0) return mx;
  int max_e; frexp(mx, &max_e);
  const bool tiny = max_e < -1023;           // ldexp(1.0, -max_e) would overflow: vector_norm rescales by DBL_MIN and calls itself once
  if (tiny) { v0 = v0 / 2.2250738585072014e-308; v1 = v1 / 2.2250738585072014e-308; mx = v0 > v1 ? v0 : v1; frexp(mx, &max_e); }
  double scale = ldexp(1.0, -max_e), csum = 1.0, frac1 = 0.0, frac2 = 0.0, x, h;
  dl_t pr, sm;
  x = v0 * scale; pr = dl_mul(x, x); sm = dl_fast_sum(csum, pr.hi); csum = sm.hi; frac1 += pr.lo; frac2 += sm.lo;
  x = v1 * scale; pr = dl_mul(x, x); sm = dl_fast_sum(csum, pr.hi); csum = sm.hi; frac1 += pr.lo; frac2 += sm.lo;
  h = sqrt(csum - 1.0 + (frac1 + frac2));
  pr = dl_mul(-h, h); sm = dl_fast_sum(csum, pr.hi); csum = sm.hi; frac1 += pr.lo; frac2 += sm.lo;
  x = csum - 1.0 + (frac1 + frac2);
  h += x / (2.0 * h);
  return tiny ? 2.2250738585072014e-308 * (h / scale) : h / scale;
}

// Python float % (Objects/floatobject.c float_rem); fmod is exact
// fmod(v, w) for w > 0, exact: each subtraction below is exact by Sterbenz' lemma
// (r in [k*w, 2*k*w] minus k*w, k a power of two), so the result equals the exact remainder.
__device__ __noinline__ double fmod_big(double v, double w) { return fmod(v, w); }      // out of line: every rs_M would carry a copy
__device__ __forceinline__ double fmod_small(double v, double w) {
  double r = fabs(v);
  if (r >= 8.0 * w) return fmod_big(v, w);
  if (r >= 4.0 * w) r -= 4.0 * w;
  if (r >= 2.0 * w) r -= 2.0 * w;
  if (r >= w) r -= w;
  return copysign(r, v);
}
__device__ __forceinline__ double py_mod(double v, double w) {
  double m = fmod_small(v, w);
  if (m != 0.0) { if ((w < 0) != (m < 0)) m += w; } else m = copysign(0.0, w);
  return m;
}

// CPython 3.12 sum() over a list whose element i is an np.float64 iff bit i of npmask is set
// (see oracle/avp_oracle.c py_sum for the derivation)
// Written as ONE fully unrolled loop over the (at most five) elements with constant indices, so that an array the caller fills with
// constant indices stays in registers: the compensated phase runs until the first np.float64 element (or the end), its correction is
// applied there, the rest is added naively -- the same operations in the same order as the two loops of the derivation.
__device__ __forceinline__ double py_sum(const double *v, int n, unsigned npmask) {
  double f = 0.0 + v[0], c = 0.0;
  bool comp = !(npmask & 1u);
#pragma unroll
  for (int i = 1; i < 5; ++i) {
    if (i < n) {
      const double x = v[i];
      if (comp && (npmask & (1u << i))) { if (c != 0.0 && isfinite(c)) f += c; comp = false; }
      if (comp) {
        const double t = f + x;
        if (fabs(f) >= fabs(x)) c += (f - t) + x; else c += (x - t) + f;
        f = t;
      } else f = f + x;
    }
  }
  if (comp && c != 0.0 && isfinite(c)) f += c;
  return f;
}

// rs_curve.py:649-656
__device__ __forceinline__ double pi_2_pi(double th) {
  while (th > AVP_PI) th -= 2.0 * AVP_PI;
  while (th < -AVP_PI) th += 2.0 * AVP_PI;
  return th;
}
// rs_curve.py:669-680
__device__ __noinline__ double rs_M(double th) {
  double phi = py_mod(th, 2.0 * AVP_PI);
  if (phi < -AVP_PI) phi += 2.0 * AVP_PI;
  if (phi > AVP_PI) phi -= 2.0 * AVP_PI;
  return phi;
}

// np.linspace(b0, b1, n)[k] (numpy/_core/function_base.py): k*step + start, last element = stop
__device__ __forceinline__ double lin_at(double start, double stop, double step, int n, int k) {
  return (k == n - 1) ? stop : d_add(d_mul((double)k, step), start);
}

// Map.convert_position_to_index (costmap.py:319-329)
__device__ __forceinline__ long long map_index(const ScenDev &S, double gx, double gy) {
  long long i0 = (long long)floor((gx - S.b[0]) / S.dx);
  long long i1 = (long long)floor((S.b[3] - gy) / S.dy) * (long long)S.stride;
  return i0 + i1;
}

// ------------------------------------------------------------------------------------------
// collision check (collision_check/collision_check.py)

struct VehGeom {
  double vb[5][2];       // create_anticlockpoint corners (costmap.py:85-121)
  double x_min, x_max, y_min, y_max;
  double v_lb, v_len;
  double lk[4], lb[4], ls[4];   // slope, intercept, sqrt(1+k^2) of the 4 boundary lines
  double ils[4];                // 1/ls: estimates only (see cell_hits)
};

__device__ __forceinline__ double shfl_dbl(double v, int src) {
  int lo = __double2loint(v), hi = __double2hiint(v);
  lo = __shfl_sync(AVP_FULL_MASK, lo, src); hi = __shfl_sync(AVP_FULL_MASK, hi, src);
  return __hiloint2double(hi, lo);
}

// Warp-collective: lane l < 4 computes corner l and boundary line l, the results are exchanged by
// shuffles (the same IEEE operations as the scalar form, evaluated once instead of 32 times and
// with a 4x shorter dependent chain).
__device__ __forceinline__ void veh_geom(const avp_config &c, double x, double y, double cs, double sn, VehGeom &g) {
  const int lane = threadIdx.x & 31, l = lane & 3;
  const double fr = c.safe_fr_dis, sd = c.safe_side_dis;
  const double lx0 = -c.lr - fr, lx1 = c.lw + c.lf + fr, ly0 = -c.lb / 2 - sd, ly1 = c.lb / 2 + sd;
  const double locx = (l == 0 || l == 3) ? lx0 : lx1, locy = (l < 2) ? ly0 : ly1;
  // trans_matrix.transpose().dot(local) + [x, y]; BLAS gemv row = fma(A[r][1], v1, A[r][0]*v0)
  const double cx = __fma_rn(-sn, locy, cs * locx) + x;
  const double cy = __fma_rn(cs, locy, sn * locx) + y;
#pragma unroll
  for (int i = 0; i < 4; ++i) { g.vb[i][0] = shfl_dbl(cx, i); g.vb[i][1] = shfl_dbl(cy, i); }
  g.vb[4][0] = g.vb[0][0]; g.vb[4][1] = g.vb[0][1];
  g.x_max = g.x_min = g.vb[0][0]; g.y_max = g.y_min = g.vb[0][1];
#pragma unroll
  for (int i = 1; i < 5; ++i) {
    if (g.vb[i][0] > g.x_max) g.x_max = g.vb[i][0]; if (g.vb[i][0] < g.x_min) g.x_min = g.vb[i][0];
    if (g.vb[i][1] > g.y_max) g.y_max = g.vb[i][1]; if (g.vb[i][1] < g.y_min) g.y_min = g.vb[i][1];
  }
  // lane l: line l between corner l and corner l+1 (collision_check.py:149-155, :180-190);
  // lanes 0/1 additionally the two side lengths (:165-169)
  const int nxt = (lane & ~3) | ((l + 1) & 3);
  const double p1x = cx, p1y = cy, p2x = shfl_dbl(cx, nxt), p2y = shfl_dbl(cy, nxt);
  const double lk = (p2y - p1y) / (p2x - p1x);
  const double lb = p1y - lk * p1x;
  const double ls = sqrt(1 + lk * lk);
  double d0 = (l == 0) ? g.vb[0][0] - g.vb[3][0] : g.vb[3][0] - g.vb[2][0];
  double d1 = (l == 0) ? g.vb[0][1] - g.vb[3][1] : g.vb[3][1] - g.vb[2][1];
  const double side = sqrt(d0 * d0 + d1 * d1);
  g.v_lb = shfl_dbl(side, 0); g.v_len = shfl_dbl(side, 1);
  const double ils = 1.0 / ls;
#pragma unroll
  for (int i = 0; i < 4; ++i) { g.lk[i] = shfl_dbl(lk, i); g.lb[i] = shfl_dbl(lb, i); g.ls[i] = shfl_dbl(ls, i); g.ils[i] = shfl_dbl(ils, i); }
}

// the per-cell predicate of distance_checker.check (collision_check.py:197-238).
// Exact, but division-free in the common case: every quotient the reference forms is first
// estimated with a reciprocal / a product; the true IEEE division is only evaluated when the
// estimate is too close to the decision threshold to be trusted (|margin| < 1e-9).
__device__ __forceinline__ bool cell_hits(const VehGeom &g, double ox, double oy) {
  // dis_i = |k_i*x + b_i - y| / sqrt(1 + k_i^2)   (collision_check.py:158-160)
  double num[4], est[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) { num[i] = fabs(g.lk[i] * ox + g.lb[i] - oy); est[i] = num[i] * g.ils[i]; }
  const double t1 = g.v_lb - 0.01, t2 = g.v_len - 0.01;
  const double m1 = fabs(est[0] - est[2]) - t1, m2 = fabs(est[1] - est[3]) - t2;
  bool c1, c2;
  const bool sure = (fabs(m1) > 1e-9) && (fabs(m2) > 1e-9) && (est[0] < 1e4) && (est[1] < 1e4) && (est[2] < 1e4) && (est[3] < 1e4);
  if (sure) { c1 = m1 < 0.0; c2 = m2 < 0.0; }
  else {                                   // near a threshold, or inf/nan slopes: the reference's own arithmetic
    double dis[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) dis[i] = num[i] / g.ls[i];
    c1 = fabs(dis[0] - dis[2]) < t1;                                    // :202
    c2 = fabs(dis[1] - dis[3]) < t2;                                    // :203
  }
  if (c1 && c2) return true;
  bool on_x = false, on_y = false;                                      // :210-230
#pragma unroll
  for (int i = 0; i < 5; ++i) on_x |= (ox == g.vb[i][0]);
  if (on_x) {
#pragma unroll
    for (int i = 0; i < 5; ++i) on_y |= (oy == g.vb[i][1]);
  }
  if (on_x && on_y) return true;
#pragma unroll
  for (int i = 0; i < 4; ++i) {                                         // :233-238: k1 == line_k[i]
    const double a = g.vb[i][1] - oy, d = g.vb[i][0] - ox, t = g.lk[i] * d;
    // fl(a/d) == lk needs a ~ lk*d to ~1e-15 relative; skip the division when they differ by far more
    if (fabs(a - t) <= 1e-9 * (fabs(a) + fabs(t)) || !(fabs(a) < 1e300) || !(fabs(t) < 1e300) || d == 0.0) {
      const double k1 = a / d;
      if (k1 == g.lk[i]) return true;
    }
  }
  return false;
}

// Column range [lo, hi] of raster columns whose x position lies in [x_min, x_max]
// (the inclusive AABB filter of get_near_obstacles, collision_check.py:60-63, applied to
// the np.where order, which is sorted by column).
__device__ __forceinline__ void col_range(const ScenDev &S, double x_min, double x_max, int &lo, int &hi) {
  const int nx = S.nx;
  const double inv = S.inv_stepx;                   // initial guesses only: the loops below settle the exact bounds
  int a = (int)floor((x_min - S.b[0]) * inv) - 1; if (a < 0) a = 0; if (a > nx - 1) a = nx - 1;
  while (a > 0 && lin_at(S.b[0], S.b[1], S.stepx, nx, a - 1) >= x_min) --a;
  while (a < nx && !(lin_at(S.b[0], S.b[1], S.stepx, nx, a) >= x_min)) ++a;
  int b = (int)floor((x_max - S.b[0]) * inv) + 1; if (b > nx - 1) b = nx - 1; if (b < 0) b = 0;
  while (b < nx - 1 && lin_at(S.b[0], S.b[1], S.stepx, nx, b + 1) <= x_max) ++b;
  while (b >= 0 && !(lin_at(S.b[0], S.b[1], S.stepx, nx, b) <= x_max)) --b;
  lo = a; hi = b;
}

// distance_checker.check (collision_check.py:144-240), warp-collective: all 32 lanes call it
// with the same pose; returns the same bool on every lane.
__device__ __forceinline__ bool check_distance_warp(const avp_config &c, const ScenDev &S, const double2 *cells,
                                                    const int32_t *col_start, double x, double y, double cs, double sn) {
  const int lane = threadIdx.x & 31;
  VehGeom g;
  veh_geom(c, x, y, cs, sn, g);
  int lo, hi;
  col_range(S, g.x_min, g.x_max, lo, hi);
  if (lo > hi) return false;
  const int beg = col_start[lo], end = col_start[hi + 1];
  bool hit = false;
  for (int base = beg; base < end; base += 32) {
    const int i = base + lane;
    bool h = false;
    if (i < end) {
      const double2 p = cells[i];
      if (p.x >= g.x_min && p.x <= g.x_max && p.y >= g.y_min && p.y <= g.y_max) h = cell_hits(g, p.x, p.y);
    }
    if (__any_sync(AVP_FULL_MASK, h)) { hit = true; break; }
  }
  return hit;
}

// The same check with the vehicle geometry in SHARED memory (one VehGeom per warp, `sg`): lanes 0..3 compute
// corner l / boundary line l / one bound of the AABB each and store them; every lane reads what it needs.
// Same IEEE operations as veh_geom, but no shuffles (veh_geom: 40 per pose) and no 32 doubles of geometry
// live in registers across the cell loop (under a 128-register cap the compiler re-derived the AABB from the
// corners in every iteration: 12 % of the pipelined kernel's instructions).
// the inflated rectangle in the vehicle frame (create_anticlockpoint, costmap.py:85-121): passed by value to the out-of-line check
struct VehDims { double lx0, lx1, ly0, ly1; };
__device__ __forceinline__ VehDims veh_dims(const avp_config &c) {
  const double fr = c.safe_fr_dis, sd = c.safe_side_dis;
  VehDims d; d.lx0 = -c.lr - fr; d.lx1 = c.lw + c.lf + fr; d.ly0 = -c.lb / 2 - sd; d.ly1 = c.lb / 2 + sd; return d;
}
__device__ __forceinline__ void veh_geom_sm(const VehDims vd, double x, double y, double cs, double sn, VehGeom *sg) {
  const int lane = threadIdx.x & 31, l = lane & 3;
  __syncwarp();                                   // the previous pose's readers are done
  if (lane < 4) {
    const double lx0 = vd.lx0, lx1 = vd.lx1, ly0 = vd.ly0, ly1 = vd.ly1;
    const double locx = (l == 0 || l == 3) ? lx0 : lx1, locy = (l < 2) ? ly0 : ly1;
    const double cx = __fma_rn(-sn, locy, cs * locx) + x;      // as veh_geom
    const double cy = __fma_rn(cs, locy, sn * locx) + y;
    sg->vb[l][0] = cx; sg->vb[l][1] = cy;
    if (l == 0) { sg->vb[4][0] = cx; sg->vb[4][1] = cy; }
  }
  __syncwarp();
  if (lane < 4) {
    const double p1x = sg->vb[l][0], p1y = sg->vb[l][1], p2x = sg->vb[l + 1][0], p2y = sg->vb[l + 1][1];   // vb[4] = vb[0]
    const double lk = (p2y - p1y) / (p2x - p1x);
    const double lb = p1y - lk * p1x;
    const double ls = sqrt(1 + lk * lk);
    sg->lk[l] = lk; sg->lb[l] = lb; sg->ls[l] = ls; sg->ils[l] = 1.0 / ls;
    if (l < 2) {                                  // the two side lengths (collision_check.py:165-169)
      const double d0 = (l == 0) ? sg->vb[0][0] - sg->vb[3][0] : sg->vb[3][0] - sg->vb[2][0];
      const double d1 = (l == 0) ? sg->vb[0][1] - sg->vb[3][1] : sg->vb[3][1] - sg->vb[2][1];
      const double side = sqrt(d0 * d0 + d1 * d1);
      if (l == 0) sg->v_lb = side; else sg->v_len = side;
    }
    // lane 0: x_min, 1: x_max, 2: y_min, 3: y_max -- the strict compare chain of veh_geom over corners 1..4
    const int dim = l >> 1; const bool want_max = (l & 1) != 0;
    double v = sg->vb[0][dim];
#pragma unroll
    for (int i = 1; i < 5; ++i) { const double w = sg->vb[i][dim]; if (want_max ? (w > v) : (w < v)) v = w; }
    if (l == 0) sg->x_min = v; else if (l == 1) sg->x_max = v; else if (l == 2) sg->y_min = v; else sg->y_max = v;
  }
  __syncwarp();
}
__device__ __noinline__ bool check_distance_warp_sm(const VehDims vd, const ScenDev &S, const double2 *cells,
                                                    const int32_t *col_start, double x, double y, double cs, double sn, VehGeom *sg) {
  const int lane = threadIdx.x & 31;
  veh_geom_sm(vd, x, y, cs, sn, sg);
  const double x_min = sg->x_min, x_max = sg->x_max, y_min = sg->y_min, y_max = sg->y_max;
  int lo, hi;
#ifdef AVP_EXACT_COLS
  col_range(S, x_min, x_max, lo, hi);
#else
  {   // a superset of the columns the rectangle touches (see check_distance_multi_sm: the per-cell AABB test below is the exact filter)
    const int nx = S.nx;
    double fa = floor((x_min - S.b[0]) * S.inv_stepx) - 1.0, fb = floor((x_max - S.b[0]) * S.inv_stepx) + 1.0;
    if (!(fa >= 0.0)) fa = 0.0; if (!(fb <= (double)(nx - 1))) fb = (double)(nx - 1);
    if (fa > fb) return false;
    lo = (int)fa; hi = (int)fb;
  }
#endif
  if (lo > hi) return false;
  const int beg = col_start[lo], end = col_start[hi + 1];
  bool hit = false;
  for (int base = beg; base < end; base += 32) {
    const int i = base + lane;
    bool h = false;
    if (i < end) {
      const double2 p = cells[i];
      if (p.x >= x_min && p.x <= x_max && p.y >= y_min && p.y <= y_max) h = cell_hits(*sg, p.x, p.y);
    }
    if (__any_sync(AVP_FULL_MASK, h)) { hit = true; break; }
  }
  return hit;
}

// Up to AVP_MULTI_POSE poses in ONE pass over the cell list: distance_checker.check(pose_0) or ... or check(pose_{np-1}) -- the
// sub-steps of one successor (hybrid_a_star.py:185-204: any hit decides), a few neighbouring points of a goal-shot course
// (:334-347, likewise).  Lanes 4p .. 4p+3 hold pose p (x, y, cos, sin) and compute its rectangle side by side (the latency of one
// pose, same IEEE operations as veh_geom); the cell loop runs once over a SUPERSET of the columns any rectangle touches and tests a
// cell against the union box first.  A cell takes part in check(pose) iff it passes that pose's inclusive AABB filter
// (get_near_obstacles, collision_check.py:55-69), which is tested per cell here exactly as there -- so the column range only has
// to cover the rectangles, it does not have to be the exact one of col_range.
#define AVP_MULTI_POSE 4
__device__ __noinline__ bool check_distance_multi_sm(const VehDims vd, const ScenDev &S, const double2 *cells, const int32_t *col_start,
                                                     int np, double x, double y, double cs, double sn, VehGeom *sgs) {
  const int lane = threadIdx.x & 31, l = lane & 3, p = lane >> 2;
  const bool act = p < np;
  VehGeom *sg = sgs + (act ? p : 0);
  __syncwarp();                                   // the previous call's readers are done
  if (act) {
    const double locx = (l == 0 || l == 3) ? vd.lx0 : vd.lx1, locy = (l < 2) ? vd.ly0 : vd.ly1;
    const double cx = __fma_rn(-sn, locy, cs * locx) + x;      // as veh_geom
    const double cy = __fma_rn(cs, locy, sn * locx) + y;
    sg->vb[l][0] = cx; sg->vb[l][1] = cy;
    if (l == 0) { sg->vb[4][0] = cx; sg->vb[4][1] = cy; }
  }
  __syncwarp();
  if (act) {
    const double p1x = sg->vb[l][0], p1y = sg->vb[l][1], p2x = sg->vb[l + 1][0], p2y = sg->vb[l + 1][1];   // vb[4] = vb[0]
    const double lk = (p2y - p1y) / (p2x - p1x);
    const double lb = p1y - lk * p1x;
    const double ls = sqrt(1 + lk * lk);
    sg->lk[l] = lk; sg->lb[l] = lb; sg->ls[l] = ls; sg->ils[l] = 1.0 / ls;
    if (l < 2) {                                  // the two side lengths (collision_check.py:165-169)
      const double d0 = (l == 0) ? sg->vb[0][0] - sg->vb[3][0] : sg->vb[3][0] - sg->vb[2][0];
      const double d1 = (l == 0) ? sg->vb[0][1] - sg->vb[3][1] : sg->vb[3][1] - sg->vb[2][1];
      const double side = sqrt(d0 * d0 + d1 * d1);
      if (l == 0) sg->v_lb = side; else sg->v_len = side;
    }
    const int dim = l >> 1; const bool want_max = (l & 1) != 0;    // lane 0: x_min, 1: x_max, 2: y_min, 3: y_max (the strict compare chain of veh_geom)
    double v = sg->vb[0][dim];
#pragma unroll
    for (int i = 1; i < 5; ++i) { const double w = sg->vb[i][dim]; if (want_max ? (w > v) : (w < v)) v = w; }
    if (l == 0) sg->x_min = v; else if (l == 1) sg->x_max = v; else if (l == 2) sg->y_min = v; else sg->y_max = v;
  }
  __syncwarp();
  double ux0 = sgs[0].x_min, ux1 = sgs[0].x_max, uy0 = sgs[0].y_min, uy1 = sgs[0].y_max;       // union box: a filter only
  for (int q = 1; q < np; ++q) {
    const double a = sgs[q].x_min, b = sgs[q].x_max, c = sgs[q].y_min, d = sgs[q].y_max;
    if (a < ux0) ux0 = a; if (b > ux1) ux1 = b; if (c < uy0) uy0 = c; if (d > uy1) uy1 = d;
  }
  // columns: one more than the estimate on either side (a column is S.stepx wide, the estimate is good to 1e-9 columns)
  const int nx = S.nx;
  double fa = floor((ux0 - S.b[0]) * S.inv_stepx) - 1.0, fb = floor((ux1 - S.b[0]) * S.inv_stepx) + 1.0;
  if (!(fa >= 0.0)) fa = 0.0; if (!(fb <= (double)(nx - 1))) fb = (double)(nx - 1);
  if (fa > fb) return false;                      // every rectangle lies beside the raster
  const int lo = (int)fa, hi = (int)fb;
  const int beg = col_start[lo], end = col_start[hi + 1];
  bool hit = false;
  for (int base = beg; base < end; base += 32) {
    const int i = base + lane;
    bool h = false;
    if (i < end) {
      const double2 c = cells[i];
      if (c.x >= ux0 && c.x <= ux1 && c.y >= uy0 && c.y <= uy1) {
        for (int q = 0; q < np && !h; ++q) {
          const VehGeom &g = sgs[q];
          if (c.x >= g.x_min && c.x <= g.x_max && c.y >= g.y_min && c.y <= g.y_max) h = cell_hits(g, c.x, c.y);
        }
      }
    }
    if (__any_sync(AVP_FULL_MASK, h)) { hit = true; break; }
  }
  return hit;
}

// two_circle_checker.check (collision_check.py:88-137), warp-collective
__device__ __forceinline__ bool check_circle_warp(const avp_config &c, const ScenDev &S, const double2 *cells,
                                                  double x, double y, double cs, double sn) {
  const int lane = threadIdx.x & 31;
  const double h2 = (c.lr + c.lw + c.lf) / 2;
  const double Rd = 0.5 * sqrt(d_pow2(h2) + d_pow2(c.lb));
  const double kf = 1.0 / 4 * (3 * c.lw + 3 * c.lf - c.lr), kr = 1.0 / 4 * (c.lw + c.lf - 3 * c.lr);
  const double fx = x + kf * cs, fy = y + kf * sn, rx = x + kr * cs, ry = y + kr * sn;
  double right, left, upper, down;
  if (fx >= rx) { right = fx + Rd; left = rx - Rd; } else { right = rx + Rd; left = fx - Rd; }
  if (fy >= ry) { upper = fy + Rd; down = ry - Rd; } else { upper = ry + Rd; down = fy - Rd; }
  bool hit = false;
  for (int base = 0; base < S.n_obs; base += 32) {
    const int i = base + lane;
    bool h = false;
    if (i < S.n_obs) {
      const double2 p = cells[i];
      if (p.x > left && p.x < right && p.y > down && p.y < upper) {
        h = (sqrt(d_pow2(p.x - fx) + d_pow2(p.y - fy)) <= Rd) || (sqrt(d_pow2(p.x - rx) + d_pow2(p.y - ry)) <= Rd);
      }
    }
    if (__any_sync(AVP_FULL_MASK, h)) { hit = true; break; }
  }
  return hit;
}

// cs, sn = cos(theta), sin(theta) of the pose (np.cos/np.sin in create_anticlockpoint, costmap.py:90-91)
__device__ __forceinline__ bool check_pose_cs_warp(const avp_config &c, const ScenDev &S, const double2 *cells,
                                                   const int32_t *col_start, double x, double y, double cs, double sn) {
  return c.collision_mode == 1 ? check_circle_warp(c, S, cells, x, y, cs, sn)
                               : check_distance_warp(c, S, cells, col_start, x, y, cs, sn);
}
__device__ __forceinline__ bool check_pose_cs_warp_sm(const avp_config &c, const ScenDev &S, const double2 *cells,
                                                      const int32_t *col_start, double x, double y, double cs, double sn, VehGeom *sg) {
  return c.collision_mode == 1 ? check_circle_warp(c, S, cells, x, y, cs, sn)
                               : check_distance_warp_sm(veh_dims(c), S, cells, col_start, x, y, cs, sn, sg);
}
__device__ __forceinline__ bool check_pose_warp(const avp_config &c, const ScenDev &S, const double2 *cells,
                                                const int32_t *col_start, double x, double y, double th) {
  double sn, cs; d_sincos(th, sn, cs);
  return check_pose_cs_warp(c, S, cells, col_start, x, y, cs, sn);
}

// ------------------------------------------------------------------------------------------
// Reeds-Shepp (path_plan/rs_curve.py).  46 word instances, in the reference's order:
//   0-1 SLS | 2-5 LSL | 6-9 LSR | 10-13 LRL | 14-17 LRL backwards | 18-21 LRLRn | 22-25 LRLRp |
//   26-29 LRSL | 30-33 LRSR | 34-37 LRSL backwards | 38-41 LRSR backwards | 42-45 LRSLR
#define RS_NINST 46

// ctype ids (set_path compares ctypes lists for equality, rs_curve.py:144)
enum { CT_SLS, CT_SRS, CT_LSL, CT_RSR, CT_LSR, CT_RSL, CT_LRL, CT_RLR, CT_LRLR, CT_RLRL, CT_LRSL, CT_RLSR,
       CT_LRSR, CT_RLSL, CT_LSRL, CT_RSLR, CT_RSRL, CT_LSLR, CT_LRSLR, CT_RLSRL, CT_COUNT };

__device__ __constant__ char rs_ct_names[CT_COUNT][8] = {
  "SLS", "SRS", "LSL", "RSR", "LSR", "RSL", "LRL", "RLR", "LRLR", "RLRL", "LRSL", "RLSR",
  "LRSR", "RLSL", "LSRL", "RSLR", "RSRL", "LSLR", "LRSLR", "RLSRL"};

__device__ __forceinline__ void rs_R(double x, double y, double &r, double &th) { r = py_hypot(x, y); th = d_atan2(y, x); }

// rs_curve.py:213-229
__device__ __forceinline__ bool rs_SLS(double x, double y, double phi, double &t, double &u, double &v) {
  phi = rs_M(phi);
  if (y > 0.0 && 0.0 < phi && phi < AVP_PI * 0.99) {
    const double tp = d_tan(phi), xd = -y / tp + x, th2 = d_tan(phi / 2.0);
    t = xd - th2; u = phi; v = sqrt(d_pow2(x - xd) + d_pow2(y)) - th2;     // ** 2 == libm pow(., 2.0)
    return true;
  } else if (y < 0.0 && 0.0 < phi && phi < AVP_PI * 0.99) {
    const double tp = d_tan(phi), xd = -y / tp + x, th2 = d_tan(phi / 2.0);
    t = xd - th2; u = phi; v = -sqrt(d_pow2(x - xd) + d_pow2(y)) - th2;
    return true;
  }
  return false;
}
// The word formulas below evaluate R(x, y) = (hypot, atan2) lazily: a result the reference computes but whose value cannot
// reach the output (the branch is not taken, or the value is never read) is not evaluated.  hypot and atan2 are pure, so this
// is exact.  atan2(Y, X) < 0 for every Y < 0 that is not so small that the quotient underflows to -0.0 (-0.0 >= 0.0 is true in
// the reference): the sign tests below only trust Y < -1e-100.
// rs_curve.py:159-167
__device__ __forceinline__ bool rs_LSL(double x, double y, double phi, double sphi, double cphi, double &t, double &u, double &v) {
  const double X = x - sphi, Y = y - 1.0 + cphi;
  if (Y < -1e-100) return false;                                   // t = atan2(Y, X) < 0
  const double tt = d_atan2(Y, X);
  if (tt >= 0.0) { const double vv = rs_M(phi - tt); if (vv >= 0.0) { t = tt; u = py_hypot(X, Y); v = vv; return true; } }
  return false;
}
// rs_curve.py:170-183
__device__ __forceinline__ bool rs_LSR(double x, double y, double phi, double sphi, double cphi, double &t, double &u, double &v) {
  const double X = x + sphi, Y = y - 1.0 - cphi;
  double u1 = py_hypot(X, Y);
  u1 = d_pow2(u1);                                                 // u1 ** 2 == libm pow(u1, 2.0)
  if (u1 >= 4.0) {
    const double t1 = d_atan2(Y, X);
    const double uu = sqrt(u1 - 4.0), theta = d_atan2(2.0, uu), tt = rs_M(t1 + theta), vv = rs_M(tt - phi);
    if (tt >= 0.0 && vv >= 0.0) { t = tt; u = uu; v = vv; return true; }
  }
  return false;
}
// rs_curve.py:186-197
__device__ __forceinline__ bool rs_LRL(double x, double y, double phi, double sphi, double cphi, double &t, double &u, double &v) {
  const double X = x - sphi, Y = y - 1.0 + cphi;
  if (fabs(X) > 4.000001 || fabs(Y) > 4.000001) return false;       // hypot(X, Y) >= max(|X|, |Y|) (1 - 2^-52) > 4: the branch is not taken
  const double u1 = py_hypot(X, Y);
  if (u1 <= 4.0) {
    const double t1 = d_atan2(Y, X);
    const double uu = -2.0 * d_asin(0.25 * u1), tt = rs_M(t1 + 0.5 * uu + AVP_PI), vv = rs_M(phi - tt + uu);
    if (tt >= 0.0 && uu <= 0.0) { t = tt; u = uu; v = vv; return true; }
  }
  return false;
}
// rs_curve.py:308-323
__device__ __forceinline__ void rs_tauOmega(double u, double v, double xi, double eta, double phi, double &tau, double &omega) {
  const double delta = rs_M(u - v);
  double su, cu, sd, cd; d_sincos(u, su, cu); d_sincos(delta, sd, cd);
  const double A = su - sd, B = cu - cd - 1.0;
  const double t1 = d_atan2(eta * A - xi * B, xi * A + eta * B);
  // both callers pass v = -u or v = u with |u| <= pi/2, and the cos restatement depends on |x| only below 2.426 (avp_cos): cos(v) is
  // the cos(u) already there, bit for bit
  const double cv = (fabs(v) == fabs(u) && fabs(u) < 2.0) ? cu : d_cos(v);
  const double t2 = 2.0 * (cd - cv - cu) + 3.0;
  tau = (t2 < 0) ? rs_M(t1 + AVP_PI) : rs_M(t1);
  omega = rs_M(tau - u + v - phi);
}
// rs_curve.py:326-337
__device__ __forceinline__ bool rs_LRLRn(double x, double y, double phi, double sphi, double cphi, double &t, double &u, double &v) {
  const double xi = x + sphi, eta = y - 1.0 - cphi, rho = 0.25 * (2.0 + sqrt(xi * xi + eta * eta));
  if (rho <= 1.0) {
    const double uu = d_acos(rho); double tt, vv; rs_tauOmega(uu, -uu, xi, eta, phi, tt, vv);
    if (tt >= 0.0 && vv <= 0.0) { t = tt; u = uu; v = vv; return true; }
  }
  return false;
}
// rs_curve.py:340-352
__device__ __forceinline__ bool rs_LRLRp(double x, double y, double phi, double sphi, double cphi, double &t, double &u, double &v) {
  const double xi = x + sphi, eta = y - 1.0 - cphi, rho = (20.0 - xi * xi - eta * eta) / 16.0;
  if (0.0 <= rho && rho <= 1.0) {
    const double uu = -d_acos(rho);
    if (uu >= -0.5 * AVP_PI) {
      double tt, vv; rs_tauOmega(uu, uu, xi, eta, phi, tt, vv);
      if (tt >= 0.0 && vv >= 0.0) { t = tt; u = uu; v = vv; return true; }
    }
  }
  return false;
}
// rs_curve.py:391-403
__device__ __forceinline__ bool rs_LRSR(double x, double y, double phi, double sphi, double cphi, double &t, double &u, double &v) {
  const double xi = x + sphi, eta = y - 1.0 - cphi;
  if (xi < -1e-100) return false;                                  // t = theta = atan2(xi, -eta) < 0
  const double rho = py_hypot(-eta, xi);
  if (rho >= 2.0) {
    const double theta = d_atan2(xi, -eta);
    const double tt = theta, uu = 2.0 - rho, vv = rs_M(tt + 0.5 * AVP_PI - phi);
    if (tt >= 0.0 && uu <= 0.0 && vv <= 0.0) { t = tt; u = uu; v = vv; return true; }
  }
  return false;
}
// rs_curve.py:406-419
__device__ __forceinline__ bool rs_LRSL(double x, double y, double phi, double sphi, double cphi, double &t, double &u, double &v) {
  const double xi = x - sphi, eta = y - 1.0 + cphi;
  const double rho = py_hypot(xi, eta);
  if (rho >= 2.0) {
    const double theta = d_atan2(eta, xi);
    const double r = sqrt(rho * rho - 4.0), uu = 2.0 - r, tt = rs_M(theta + d_atan2(r, -2.0)), vv = rs_M(phi - 0.5 * AVP_PI - tt);
    if (tt >= 0.0 && uu <= 0.0 && vv <= 0.0) { t = tt; u = uu; v = vv; return true; }
  }
  return false;
}
// rs_curve.py:494-510 (theta of R(xi, eta) is computed by the reference and never read)
__device__ __forceinline__ bool rs_LRSLR(double x, double y, double phi, double sphi, double cphi, double &t, double &u, double &v) {
  const double xi = x + sphi, eta = y - 1.0 - cphi;
  const double rho = py_hypot(xi, eta);
  if (rho >= 2.0) {
    const double uu = 4.0 - sqrt(rho * rho - 4.0);
    if (uu <= 0.0) {
      const double tt = rs_M(d_atan2((4.0 - uu) * xi - 2.0 * eta, -2.0 * xi + (uu - 4.0) * eta)), vv = rs_M(tt - phi);
      if (tt >= 0.0 && vv >= 0.0) { t = tt; u = uu; v = vv; return true; }
    }
  }
  return false;
}

// normalised query of generate_path (rs_curve.py:627-634)
struct RsQuery { double x, y, phi, xb, yb, sp, cp; };   // sp, cp = sin(phi), cos(phi): sin(-phi) = -sp, cos(-phi) = cp bit for bit
// c, s = cos(q0[2]), sin(q0[2]) (callers that already hold them pass them in: same function, same argument, same bits)
__device__ __forceinline__ void rs_query_cs(const double q0[3], double c, double s, const double q1[3], double maxc, RsQuery &Q) {
  const double dx = q1[0] - q0[0], dy = q1[1] - q0[1], dth = q1[2] - q0[2];
  Q.x = (c * dx + s * dy) * maxc; Q.y = (-s * dx + c * dy) * maxc; Q.phi = dth;
  double cp, sp; d_sincos(dth, sp, cp);
  Q.sp = sp; Q.cp = cp;
  Q.xb = Q.x * cp + Q.y * sp; Q.yb = Q.x * sp - Q.y * cp;          // rs_curve.py:286-287, :456-457
}
// the same with sin / cos of dth = q1[2] - q0[2] supplied by the caller (k_plan evaluates them beside sin / cos of q0[2], on other lanes)
__device__ __forceinline__ void rs_query_cs2(const double q0[3], double c, double s, const double q1[3], double maxc, double sp, double cp, RsQuery &Q) {
  const double dx = q1[0] - q0[0], dy = q1[1] - q0[1], dth = q1[2] - q0[2];
  Q.x = (c * dx + s * dy) * maxc; Q.y = (-s * dx + c * dy) * maxc; Q.phi = dth;
  Q.sp = sp; Q.cp = cp;
  Q.xb = Q.x * cp + Q.y * sp; Q.yb = Q.x * sp - Q.y * cp;
}
__device__ __forceinline__ void rs_query(const double q0[3], const double q1[3], double maxc, RsQuery &Q) {
  double s0, c0; d_sincos(q0[2], s0, c0);
  rs_query_cs(q0, c0, s0, q1, maxc, Q);
}

// one word instance -> (valid, t, u, v)
__device__ __forceinline__ bool rs_eval_instance(int inst, const RsQuery &Q, double &t, double &u, double &v) {
  if (inst < 2) return rs_SLS(Q.x, inst ? -Q.y : Q.y, inst ? -Q.phi : Q.phi, t, u, v);
  const int r = (inst - 2) & 3, fam = (inst - 2) >> 2;   // fam 0 LSL,1 LSR,2 LRL,3 LRLb,4 LRLRn,5 LRLRp,6 LRSL,7 LRSR,8 LRSLb,9 LRSRb,10 LRSLR
  const bool back = (fam == 3 || fam == 8 || fam == 9);
  const double bx = back ? Q.xb : Q.x, by = back ? Q.yb : Q.y;
  const bool neg = (r == 1 || r == 2);
  const double x = (r & 1) ? -bx : bx, y = (r & 2) ? -by : by, phi = neg ? -Q.phi : Q.phi, sp = neg ? -Q.sp : Q.sp, cp = Q.cp;
  switch (fam) {
    case 0: return rs_LSL(x, y, phi, sp, cp, t, u, v);
    case 1: return rs_LSR(x, y, phi, sp, cp, t, u, v);
    case 2: case 3: return rs_LRL(x, y, phi, sp, cp, t, u, v);
    case 4: return rs_LRLRn(x, y, phi, sp, cp, t, u, v);
    case 5: return rs_LRLRp(x, y, phi, sp, cp, t, u, v);
    case 6: case 8: return rs_LRSL(x, y, phi, sp, cp, t, u, v);
    case 7: case 9: return rs_LRSR(x, y, phi, sp, cp, t, u, v);
    default: return rs_LRSLR(x, y, phi, sp, cp, t, u, v);
  }
}

// instance -> (ctype id, lengths[], n, np-type mask); arrangement per rs_curve.py:200-534
__device__ __forceinline__ int rs_arrange_inl(int inst, double t, double u, double v, int xy_np, int phi_np,
                                              double *l, int &ct, unsigned &mask) {
  const unsigned P = phi_np ? 1u : 0u;
  if (inst < 2) { l[0] = t; l[1] = u; l[2] = v; ct = inst ? CT_SRS : CT_SLS; mask = (xy_np ? 1u : 0u) | (P << 1); return 3; }
  const int r = (inst - 2) & 3, fam = (inst - 2) >> 2, hi = r >> 1;
  const double s = (r & 1) ? -1.0 : 1.0;
  switch (fam) {
    case 0: l[0] = s * t; l[1] = s * u; l[2] = s * v; ct = hi ? CT_RSR : CT_LSL; mask = P << 2; return 3;
    case 1: l[0] = s * t; l[1] = s * u; l[2] = s * v; ct = hi ? CT_RSL : CT_LSR; mask = P << 2; return 3;
    case 2: l[0] = s * t; l[1] = s * u; l[2] = s * v; ct = hi ? CT_RLR : CT_LRL; mask = P << 2; return 3;
    case 3: l[0] = s * v; l[1] = s * u; l[2] = s * t; ct = hi ? CT_RLR : CT_LRL; mask = P; return 3;
    case 4: l[0] = s * t; l[1] = s * u; l[2] = s * -u; l[3] = s * v; ct = hi ? CT_RLRL : CT_LRLR; mask = P << 3; return 4;
    case 5: l[0] = s * t; l[1] = s * u; l[2] = s * u; l[3] = s * v; ct = hi ? CT_RLRL : CT_LRLR; mask = P << 3; return 4;
    case 6: l[0] = s * t; l[1] = s * -AVP_HALF_PI; l[2] = s * u; l[3] = s * v; ct = hi ? CT_RLSR : CT_LRSL; mask = P << 3; return 4;
    case 7: l[0] = s * t; l[1] = s * -AVP_HALF_PI; l[2] = s * u; l[3] = s * v; ct = hi ? CT_RLSL : CT_LRSR; mask = P << 3; return 4;
    case 8: l[0] = s * v; l[1] = s * u; l[2] = s * -AVP_HALF_PI; l[3] = s * t; ct = hi ? CT_RSLR : CT_LSRL; mask = P; return 4;
    case 9: l[0] = s * v; l[1] = s * u; l[2] = s * -AVP_HALF_PI; l[3] = s * t; ct = hi ? CT_LSLR : CT_RSRL; mask = P; return 4;
    default: l[0] = s * t; l[1] = s * -AVP_HALF_PI; l[2] = s * u; l[3] = s * -AVP_HALF_PI; l[4] = s * v; ct = hi ? CT_RLSRL : CT_LRSLR; mask = P << 4; return 5;
  }
}

// out of line for the callers off the hot path (one copy per kernel); the rs items of k_plan inline the body: their length arrays
// stay in registers
__device__ __noinline__ int rs_arrange(int inst, double t, double u, double v, int xy_np, int phi_np,
                                       double *l, int &ct, unsigned &mask) {
  return rs_arrange_inl(inst, t, u, v, xy_np, phi_np, l, ct, mask);
}

// Candidate results of the 46 instances (filled in parallel), then the sequential
// set_path / calc_optimal_path semantics (rs_curve.py:137-156, :99-110).
struct RsCand { double t, u, v, L; };   // L = sum(|lengths|) of the arranged word (rs_curve.py:148)

struct RsBest { int ok; int degenerate; int n; int ct; double len[5]; double L; /* normalised */ int inst; /* winning word instance */ };

// The same with the arranged word kept (k_plan): the selection reads lengths / ctype / np-type mask instead of re-deriving them
// with rs_arrange for every comparison (88 k warp-cycles per pop in round 1's profile).
struct RsCandX { double t, u, v, L; double len[5]; int32_t n, ct; uint32_t mask; int32_t pad; };

// instances that can share a ctype form 11 groups (same family pair), in instance order
__device__ __constant__ int8_t rs_grp_begin[12] = {0, 1, 2, 6, 10, 18, 26, 30, 34, 38, 42, 46};
#define RS_NGROUP 11
struct RsGroupBest { int inst; int degenerate; double Lm; };    // inst < 0: nothing retained

// L of one arranged candidate (computed where the candidate is evaluated, in parallel)
__device__ __forceinline__ double rs_cand_L(int inst, const RsCand &c, int xy_np, int phi_np) {
  double l[5], a[5]; int ct; unsigned mask;
  const int n = rs_arrange(inst, c.t, c.u, c.v, xy_np, phi_np, l, ct, mask);
  for (int i = 0; i < n; ++i) a[i] = fabs(l[i]);
  return py_sum(a, n, mask);
}

// set_path + calc_optimal_path restricted to one group (rs_curve.py:137-156, :99-110)
__device__ __forceinline__ void rs_select_group(const RsCand *cand, unsigned long long valid, int g, int xy_np, int phi_np,
                                                double maxc, RsGroupBest &out) {
  unsigned long long retained = 0ull;
  out.inst = -1; out.degenerate = 0; out.Lm = 0.0;
  const int beg = rs_grp_begin[g], end = rs_grp_begin[g + 1];
  for (int inst = beg; inst < end; ++inst) {
    if (!((valid >> inst) & 1ull)) continue;
    double l[5]; int ct; unsigned mask;
    const int n = rs_arrange(inst, cand[inst].t, cand[inst].u, cand[inst].v, xy_np, phi_np, l, ct, mask);
    bool dup = false;
    for (int e = beg; e < inst; ++e) {
      if (!((retained >> e) & 1ull)) continue;
      double le[5], d[5]; int cte; unsigned me;
      rs_arrange(e, cand[e].t, cand[e].u, cand[e].v, xy_np, phi_np, le, cte, me);
      if (cte != ct) continue;
      for (int i = 0; i < n; ++i) d[i] = le[i] - l[i];
      if (py_sum(d, n, mask) <= 0.01) { dup = true; break; }     // rs_curve.py:143-146
    }
    if (dup) continue;
    const double L = cand[inst].L;
    if (L >= 1000.0) continue;                                    // MAX_LENGTH
    if (!(L >= 0.01)) { out.degenerate = 1; continue; }           // assert (rs_curve.py:153)
    retained |= (1ull << inst);
    const double Lm = L / maxc;
    if (out.inst < 0 || Lm <= out.Lm) { out.inst = inst; out.Lm = Lm; }     // last <= wins
  }
}

// rs_select_group on candidates that carry their arranged word
__device__ __forceinline__ void rs_select_group_x(const RsCandX *cand, unsigned long long valid, int g, double maxc, RsGroupBest &out) {
  unsigned long long retained = 0ull;
  out.inst = -1; out.degenerate = 0; out.Lm = 0.0;
  const int beg = rs_grp_begin[g], end = rs_grp_begin[g + 1];
  for (int inst = beg; inst < end; ++inst) {
    if (!((valid >> inst) & 1ull)) continue;
    const RsCandX &c = cand[inst];
    bool dup = false;
    for (int e = beg; e < inst; ++e) {
      if (!((retained >> e) & 1ull)) continue;
      const RsCandX &o = cand[e];
      if (o.ct != c.ct) continue;
      double d[5];
#pragma unroll
      for (int i = 0; i < 5; ++i) d[i] = (i < c.n) ? o.len[i] - c.len[i] : 0.0;
      if (py_sum(d, c.n, c.mask) <= 0.01) { dup = true; break; }     // rs_curve.py:143-146
    }
    if (dup) continue;
    const double L = c.L;
    if (L >= 1000.0) continue;                                    // MAX_LENGTH
    if (!(L >= 0.01)) { out.degenerate = 1; continue; }           // assert (rs_curve.py:153)
    retained |= (1ull << inst);
    const double Lm = L / maxc;
    if (out.inst < 0 || Lm <= out.Lm) { out.inst = inst; out.Lm = Lm; }     // last <= wins
  }
}

// combine the group winners in order (equivalent to the sequential scan over all retained words)
__device__ __forceinline__ void rs_combine_groups(const RsGroupBest *gb, const RsCand *cand, int xy_np, int phi_np, RsBest &best) {
  best.ok = 0; best.degenerate = 0; best.inst = -1;
  int bi = -1; double minL = 0.0;
  for (int g = 0; g < RS_NGROUP; ++g) {
    best.degenerate |= gb[g].degenerate;
    if (gb[g].inst < 0) continue;
    if (bi < 0 || gb[g].Lm <= minL) { bi = gb[g].inst; minL = gb[g].Lm; }
  }
  if (bi >= 0) {
    unsigned mask;
    best.ok = 1; best.n = rs_arrange(bi, cand[bi].t, cand[bi].u, cand[bi].v, xy_np, phi_np, best.len, best.ct, mask);
    best.L = cand[bi].L; best.inst = bi;
  }
}

// single-thread selection (API kernels)
__device__ __forceinline__ void rs_select(RsCand *cand, unsigned long long valid, int xy_np, int phi_np,
                                          double maxc, RsBest &best) {
  RsGroupBest gb[RS_NGROUP];
  for (int g = 0; g < RS_NGROUP; ++g) rs_select_group(cand, valid, g, xy_np, phi_np, maxc, gb[g]);
  rs_combine_groups(gb, cand, xy_np, phi_np, best);
}

// rs_curve.py:597-624
__device__ __noinline__ void rs_interpolate(double l, char m, double maxc, double ox, double oy, double oyaw,
                                               double &px, double &py, double &pyaw, int &dir) {
  if (m == 'S') {
    double so, co; d_sincos(oyaw, so, co);
    px = ox + l / maxc * co; py = oy + l / maxc * so; pyaw = oyaw;
  } else {
    double sl, cl; d_sincos(l, sl, cl);
    const double ldx = sl / maxc;
    double ldy = 0.0;
    if (m == 'L') ldy = (1.0 - cl) / maxc; else if (m == 'R') ldy = (1.0 - cl) / (-maxc);
    double cm, sm; d_sincos(-oyaw, sm, cm);
    const double gdx = cm * ldx + sm * ldy, gdy = -sm * ldx + cm * ldy;
    px = ox + gdx; py = oy + gdy;
  }
  if (m == 'L') pyaw = oyaw + l; else if (m == 'R') pyaw = oyaw - l;
  dir = (l > 0.0) ? 1 : -1;
}

// rs_curve.py:537-594 + the global transform of calc_all_paths (:124-130); single thread.
// Returns the number of course points, or -1 on buffer overflow.
__device__ __noinline__ int rs_course(const RsBest &w, double maxc, double step_size, const double q0[3], int cap,
                                      double *X, double *Y, double *YAW, int32_t *DIR) {
  const double step = step_size * maxc;
  const int point_num = (int)(w.L / step) + w.n + 3;
  if (point_num > cap) return -1;
  for (int i = 0; i < point_num; ++i) { X[i] = 0.0; Y[i] = 0.0; YAW[i] = 0.0; DIR[i] = 0; }
  const char *mode = rs_ct_names[w.ct];
  int ind = 1;
  DIR[0] = (w.len[0] > 0.0) ? 1 : -1;
  double d = (w.len[0] > 0.0) ? step : -step, pd = d, ll = 0.0;
  for (int i = 0; i < w.n; ++i) {
    const char m = mode[i]; const double l = w.len[i];
    d = (l > 0.0) ? step : -step;
    const double ox = X[ind], oy = Y[ind], oyaw = YAW[ind];
    ind -= 1;
    if (i >= 1 && (w.len[i - 1] * w.len[i]) > 0) pd = -d - ll; else pd = d - ll;
    while (fabs(pd) <= fabs(l)) {
      ind += 1; if (ind >= point_num) return -1;
      rs_interpolate(pd, m, maxc, ox, oy, oyaw, X[ind], Y[ind], YAW[ind], DIR[ind]);
      pd += d;
    }
    ll = l - pd - d;
    ind += 1; if (ind >= point_num) return -1;
    rs_interpolate(l, m, maxc, ox, oy, oyaw, X[ind], Y[ind], YAW[ind], DIR[ind]);
  }
  int n = point_num;
  while (n > 0 && X[n - 1] == 0.0) --n;
  const double cm = d_cos(-q0[2]), sm = d_sin(-q0[2]);
  for (int i = 0; i < n; ++i) {
    const double ix = X[i], iy = Y[i];
    X[i] = cm * ix + sm * iy + q0[0]; Y[i] = -sm * ix + cm * iy + q0[1];
    YAW[i] = pi_2_pi(YAW[i] + q0[2]);
  }
  return n;
}
