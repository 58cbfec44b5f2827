/* avp_sincos.h -- fp64 sin/cos that return the SAME BITS as the host libm the reference runs on.
 *
 * Why this exists: the reference computes successor poses with np.cos/np.sin
 * (reference hybrid_a_star.py:146-151) and then de-duplicates successors by exact
 * floating-point equality of (x, y, theta) against the closed and open lists
 * (hybrid_a_star.py:155-172).  Whether fl(fl(x + a) - a) == x holds depends on the low
 * bits of a = 1.5*cos(theta'), so a device sin/cos that is "only" <=1 ulp accurate changes
 * which successors are merged and therefore every node index downstream.
 *
 * What it is: a restatement of the table-driven algorithm used by glibc 2.39's x86-64
 * FMA build of sin()/cos() for |x| < 105414350 (table step 1/128, degree-5/6 correction
 * polynomials, 4-constant Cody-Waite reduction by pi/2), with every fused multiply-add
 * written out explicitly so that host (gcc -ffp-contract=off) and device (nvcc -fmad=false)
 * builds evaluate the identical sequence of IEEE operations.  It is proven, not assumed:
 * tests/test_sincos_bits.py compares it with libm on >1e8 arguments (0 mismatches).
 * np.sin/np.cos on float64 dispatch to the same libm entry points in this image
 * (checked by tests/golden/gen_leaf_golden.py).
 *
 * Arguments with |x| >= 105414350 (never produced on the planner path: every angle is
 * wrapped to a few multiples of pi) fall back to the toolchain's sin/cos.
 */
#ifndef AVP_SINCOS_H
#define AVP_SINCOS_H

#include <math.h>
#include <stdint.h>
#include <string.h>
#include "avp_sincos_tab.h"

#if defined(__CUDACC__)
#define AVP_HD __host__ __device__ __forceinline__
#else
#define AVP_HD static inline
#endif

#if defined(__CUDACC__)
__device__ static const double avp_sincos_tab_d[AVP_SINCOS_TAB_N] = {AVP_SINCOS_TAB_VALUES};
#endif
#if defined(__CUDA_ARCH__)
#define AVP_FMA(a, b, c) __fma_rn((a), (b), (c))
#define AVP_ADD(a, b) __dadd_rn((a), (b))
#define AVP_MUL(a, b) __dmul_rn((a), (b))
#define AVP_SCT(i) (avp_sincos_tab_d[(i)])
#else
#define AVP_FMA(a, b, c) fma((a), (b), (c))
#define AVP_ADD(a, b) ((a) + (b))
#define AVP_MUL(a, b) ((a) * (b))
static const double avp_sincos_tab_h[AVP_SINCOS_TAB_N] = {AVP_SINCOS_TAB_VALUES};
#define AVP_SCT(i) (avp_sincos_tab_h[(i)])
#endif

AVP_HD uint64_t avp_d2u(double x) {
#if defined(__CUDA_ARCH__)
  return (uint64_t)__double_as_longlong(x);
#else
  uint64_t u; memcpy(&u, &x, 8); return u;
#endif
}
AVP_HD double avp_u2d(uint64_t u) {
#if defined(__CUDA_ARCH__)
  return __longlong_as_double((long long)u);
#else
  double x; memcpy(&x, &u, 8); return x;
#endif
}
AVP_HD double avp_copysign(double mag, double sgn) {
  return avp_u2d((avp_d2u(mag) & 0x7fffffffffffffffULL) | (avp_d2u(sgn) & 0x8000000000000000ULL));
}

/* constants (values read back from the installed libm; see tools/extract_glibc_sincos.py) */
#define AVP_SC_BIG    0x1.8p+45                   /* rounds |x| to a multiple of 1/128 */
#define AVP_SC_TOINT  0x1.8p+52
#define AVP_SC_HPINV  0x1.45f306dc9c883p-1        /* 2/pi */
#define AVP_SC_HP0    0x1.921fb54442d18p+0        /* pi/2 hi */
#define AVP_SC_HP1    0x1.1a62633145c07p-54       /* pi/2 lo */
#define AVP_SC_MP1    0x1.921fb58000000p+0
#define AVP_SC_MP2   -0x1.dde973c000000p-27
#define AVP_SC_PP3   -0x1.cb3b398000000p-55
#define AVP_SC_PP4   -0x1.d747f23e32ed7p-83
#define AVP_SC_S1    -0x1.5555555555555p-3
#define AVP_SC_S2     0x1.1111111110ecep-7
#define AVP_SC_S3    -0x1.a01a019db08b8p-13
#define AVP_SC_S4     0x1.71de27b9a7ed9p-19
#define AVP_SC_S5    -0x1.addffc2fcdf59p-26
#define AVP_SC_SN3   -0x1.5555555555515p-3
#define AVP_SC_SN5    0x1.11110e829872fp-7
#define AVP_SC_CS2    0x1.0p-1
#define AVP_SC_CS4   -0x1.5555555555535p-5
#define AVP_SC_CS6    0x1.6c16bedd9e239p-10

/* x + t, t = ((P(xx)*x - 0.5*dx)*xx + dx), P the odd Taylor tail of sin */
AVP_HD double avp_sc_taylor(double xx, double x, double dx) {
  double p = AVP_FMA(AVP_SC_S5, xx, AVP_SC_S4);
  p = AVP_FMA(p, xx, AVP_SC_S3);
  p = AVP_FMA(p, xx, AVP_SC_S2);
  p = AVP_FMA(p, xx, AVP_SC_S1);
  double t = AVP_FMA(AVP_FMA(p, x, -AVP_MUL(0.5, dx)), xx, dx);
  return AVP_ADD(x, t);
}

/* sine of (x + dx), |x| < ~0.86, |dx| tiny */
AVP_HD double avp_sc_do_sin(double x, double dx) {
  const double xold = x;
  const double ax = fabs(x);
  if (ax < 0.126) return avp_sc_taylor(AVP_MUL(x, x), x, dx);
  if (x <= 0.0) dx = -dx;
  const double u = AVP_ADD(AVP_SC_BIG, ax);
  const int k = (int)((uint32_t)avp_d2u(u) << 2);
  x = AVP_ADD(ax, -AVP_ADD(u, -AVP_SC_BIG));
  const double xx = AVP_MUL(x, x);
  const double s = AVP_ADD(x, AVP_FMA(AVP_MUL(x, xx), AVP_FMA(xx, AVP_SC_SN5, AVP_SC_SN3), dx));
  const double c = AVP_FMA(x, dx, AVP_MUL(xx, AVP_FMA(xx, AVP_FMA(xx, AVP_SC_CS6, AVP_SC_CS4), AVP_SC_CS2)));
  const double sn = AVP_SCT(k), ssn = AVP_SCT(k + 1), cs = AVP_SCT(k + 2), ccs = AVP_SCT(k + 3);
  const double cor = AVP_FMA(s, cs, AVP_FMA(-c, sn, AVP_FMA(s, ccs, ssn)));
  return avp_copysign(AVP_ADD(sn, cor), xold);
}

/* cosine of (x + dx) */
AVP_HD double avp_sc_do_cos(double x, double dx) {
  if (x < 0.0) dx = -dx;
  const double ax = fabs(x);
  const double u = AVP_ADD(AVP_SC_BIG, ax);
  const int k = (int)((uint32_t)avp_d2u(u) << 2);
  x = AVP_ADD(AVP_ADD(ax, -AVP_ADD(u, -AVP_SC_BIG)), dx);
  const double xx = AVP_MUL(x, x);
  const double s = AVP_FMA(AVP_MUL(x, xx), AVP_FMA(xx, AVP_SC_SN5, AVP_SC_SN3), x);
  const double c = AVP_MUL(xx, AVP_FMA(xx, AVP_FMA(xx, AVP_SC_CS6, AVP_SC_CS4), AVP_SC_CS2));
  const double sn = AVP_SCT(k), ssn = AVP_SCT(k + 1), cs = AVP_SCT(k + 2), ccs = AVP_SCT(k + 3);
  const double cor = AVP_FMA(-s, sn, AVP_FMA(-c, cs, AVP_FMA(-s, ssn, ccs)));
  return AVP_ADD(cs, cor);
}

/* x = n*(pi/2) + (a + da); returns (n + koff) & 3 */
AVP_HD int avp_sc_reduce(double x, double *a, double *da, int koff) {
  const double t = AVP_FMA(x, AVP_SC_HPINV, AVP_SC_TOINT);
  const double xn = AVP_ADD(t, -AVP_SC_TOINT);
  const int n = ((int)(uint32_t)avp_d2u(t) + koff) & 3;
  const double y = AVP_FMA(-xn, AVP_SC_MP2, AVP_FMA(-xn, AVP_SC_MP1, x));
  const double t2 = AVP_FMA(-xn, AVP_SC_PP3, y);
  double db = AVP_FMA(-xn, AVP_SC_PP3, AVP_ADD(y, -t2));
  const double b = AVP_FMA(-xn, AVP_SC_PP4, t2);
  db = AVP_ADD(db, AVP_FMA(-xn, AVP_SC_PP4, AVP_ADD(t2, -b)));
  *a = b;
  *da = db;
  return n;
}

AVP_HD double avp_sc_do_sincos(double a, double da, int n) {
  const double r = (n & 1) ? avp_sc_do_cos(a, da) : avp_sc_do_sin(a, da);
  return (n & 2) ? -r : r;
}

AVP_HD double avp_sin(double x) {
  const uint32_t k = (uint32_t)(avp_d2u(x) >> 32) & 0x7fffffffu;
  if (k < 0x3e500000u) return x;                       /* |x| < 2^-26 */
  if (k < 0x3feb6000u) return avp_sc_do_sin(x, 0.0);   /* |x| < 0.855469 */
  if (k < 0x400368fdu) {                               /* |x| < 2.426265 */
    const double t = AVP_ADD(AVP_SC_HP0, -fabs(x));
    return avp_copysign(avp_sc_do_cos(t, AVP_SC_HP1), x);
  }
  if (k < 0x419921fbu) {                               /* |x| < 105414350 */
    double a, da;
    const int n = avp_sc_reduce(x, &a, &da, 0);
    return avp_sc_do_sincos(a, da, n);
  }
  return sin(x);
}

AVP_HD double avp_cos(double x) {
  const uint32_t k = (uint32_t)(avp_d2u(x) >> 32) & 0x7fffffffu;
  if (k < 0x3e400000u) return 1.0;                     /* |x| < 2^-27 */
  if (k < 0x3feb6000u) return avp_sc_do_cos(x, 0.0);
  if (k < 0x400368fdu) {
    const double y = AVP_ADD(AVP_SC_HP0, -fabs(x));
    const double a = AVP_ADD(y, AVP_SC_HP1);
    const double da = AVP_ADD(AVP_ADD(y, -a), AVP_SC_HP1);
    return avp_sc_do_sin(a, da);
  }
  if (k < 0x419921fbu) {
    double a, da;
    const int n = avp_sc_reduce(x, &a, &da, 1);
    return avp_sc_do_sincos(a, da, n);
  }
  return cos(x);
}

#endif /* AVP_SINCOS_H */
