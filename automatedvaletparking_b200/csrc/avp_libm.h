/* avp_libm.h -- fp64 atan2 / asin / acos / tan / pow(x,2) returning the SAME BITS as the host
 * libm (glibc 2.39, x86-64 FMA variants) that the reference's math.* calls resolve to.
 *
 * Why: Reeds-Shepp lengths (rs_curve.py:159-534) enter f = g + max(dijkstra/100, rs.L);
 * distinct nodes regularly have bit-identical rs.L in the reference (Case1: nodes 194 and
 * 294), and the heap order of such ties is structural.  A 1-ulp different atan2 breaks the
 * tie and changes the expanded-node order, so "node-index bit-exact" needs these functions
 * bit-exact, not merely accurate.
 *
 * Each function restates the algorithm glibc uses on the finite, non-degenerate inputs the
 * planner produces (table-driven IBM Accurate Mathematical Library kernels), with every
 * fused multiply-add explicit; host and device builds evaluate the identical sequence of
 * IEEE operations.  tests/csrc/libm_bits.c proves bit-equality against libm on >1e8 inputs.
 * Tables are read back from the installed libm by tools/extract_glibc_tables.py.
 * Non-finite inputs fall back to the toolchain's function.
 */
#ifndef AVP_LIBM_H
#define AVP_LIBM_H

#include "avp_sincos.h"
#include "avp_libm_tab.h"

#if defined(__CUDACC__)
__device__ static const double avp_atan_cij_d[AVP_ATAN_CIJ_N] = {AVP_ATAN_CIJ_VALUES};
__device__ static const double avp_asncs_d[AVP_ASNCS_N] = {AVP_ASNCS_VALUES};
__device__ static const double avp_inroot_d[AVP_INROOT_N] = {AVP_INROOT_VALUES};
__device__ static const double avp_powtwo_d[AVP_POWTWO_N] = {AVP_POWTWO_VALUES};
__device__ static const double avp_powlog_d[AVP_POWLOG_N] = {AVP_POWLOG_VALUES};
__device__ static const double avp_xfg_d[AVP_XFG_N] = {AVP_XFG_VALUES};
__device__ static const unsigned long long avp_exptab_d[AVP_EXPTAB_N] = {AVP_EXPTAB_VALUES};
#endif
#if defined(__CUDA_ARCH__)
#define AVP_CIJ(i) (avp_atan_cij_d[(i)])
#define AVP_ASN(i) (avp_asncs_d[(i)])
#define AVP_INROOT(i) (avp_inroot_d[(i)])
#define AVP_POWTWO(i) (avp_powtwo_d[(i)])
#define AVP_POWLOG(i) (avp_powlog_d[(i)])
#define AVP_XFG(i) (avp_xfg_d[(i)])
#define AVP_EXPTAB(i) (avp_exptab_d[(i)])
#define AVP_DIV(a, b) __ddiv_rn((a), (b))
#else
static const double avp_atan_cij_h[AVP_ATAN_CIJ_N] = {AVP_ATAN_CIJ_VALUES};
static const double avp_asncs_h[AVP_ASNCS_N] = {AVP_ASNCS_VALUES};
static const double avp_inroot_h[AVP_INROOT_N] = {AVP_INROOT_VALUES};
static const double avp_powtwo_h[AVP_POWTWO_N] = {AVP_POWTWO_VALUES};
#define AVP_CIJ(i) (avp_atan_cij_h[(i)])
#define AVP_ASN(i) (avp_asncs_h[(i)])
#define AVP_INROOT(i) (avp_inroot_h[(i)])
static const double avp_powlog_h[AVP_POWLOG_N] = {AVP_POWLOG_VALUES};
static const unsigned long long avp_exptab_h[AVP_EXPTAB_N] = {AVP_EXPTAB_VALUES};
#define AVP_POWTWO(i) (avp_powtwo_h[(i)])
static const double avp_xfg_h[AVP_XFG_N] = {AVP_XFG_VALUES};
#define AVP_POWLOG(i) (avp_powlog_h[(i)])
#define AVP_XFG(i) (avp_xfg_h[(i)])
#define AVP_EXPTAB(i) (avp_exptab_h[(i)])
#define AVP_DIV(a, b) ((a) / (b))
#endif
#define AVP_SUB(a, b) AVP_ADD((a), -(b))

/* ---------------------------------------------------------------- atan2 (e_atan2.c) */
#define AVP_AT_HPI    0x1.921fb54442d18p+0
#define AVP_AT_HPI1   0x1.1a62633145c07p-54
#define AVP_AT_OPI    0x1.921fb54442d18p+1
#define AVP_AT_OPI1   0x1.1a62633145c07p-53
#define AVP_AT_D3    -0x1.5555555555555p-2
#define AVP_AT_D5     0x1.99999999997fdp-3
#define AVP_AT_D7    -0x1.24924923f7603p-3
#define AVP_AT_D9     0x1.c71c6e5129a3bp-4
#define AVP_AT_D11   -0x1.7458022b13c25p-4
#define AVP_AT_D13    0x1.375f08b31cbcep-4

AVP_HD double avp_at_poly(double v) {      /* d3 + v*(d5 + v*(d7 + v*(d9 + v*(d11 + v*d13)))) */
  double p = AVP_FMA(v, AVP_AT_D13, AVP_AT_D11);
  p = AVP_FMA(v, p, AVP_AT_D9);
  p = AVP_FMA(v, p, AVP_AT_D7);
  p = AVP_FMA(v, p, AVP_AT_D5);
  return AVP_FMA(v, p, AVP_AT_D3);
}
AVP_HD int avp_at_index(double u) {        /* i = (TWO52 + TWO8*u) - TWO52 - 16 */
  const double t = AVP_SUB(AVP_FMA(u, 0x1.0p+8, 0x1.0p+52), 0x1.0p+52);
  return (int)t - 16;
}
AVP_HD double avp_at_tpoly(int i, double v) {   /* c2 + v*(c3 + v*(c4 + v*(c5 + v*c6))) */
  double p = AVP_FMA(v, AVP_CIJ(7 * i + 6), AVP_CIJ(7 * i + 5));
  p = AVP_FMA(v, p, AVP_CIJ(7 * i + 4));
  p = AVP_FMA(v, p, AVP_CIJ(7 * i + 3));
  return AVP_FMA(v, p, AVP_CIJ(7 * i + 2));
}

AVP_HD double avp_atan2(double y, double x) {
  const uint32_t ux = (uint32_t)(avp_d2u(x) >> 32), uy = (uint32_t)(avp_d2u(y) >> 32);
  const uint32_t lx = (uint32_t)avp_d2u(x), ly = (uint32_t)avp_d2u(y);
  if ((ux & 0x7ff00000u) == 0x7ff00000u || (uy & 0x7ff00000u) == 0x7ff00000u) return atan2(y, x);   /* inf / nan */
  if ((uy & 0x7fffffffu) == 0 && ly == 0) {                    /* y = +-0 */
    if ((int32_t)uy >= 0) return ((int32_t)ux >= 0) ? 0.0 : AVP_AT_OPI;
    return ((int32_t)ux >= 0) ? -0.0 : -AVP_AT_OPI;
  }
  if ((ux & 0x7fffffffu) == 0 && lx == 0) return (y > 0) ? AVP_AT_HPI : -AVP_AT_HPI;   /* x = +-0 */
  double ax = (x < 0) ? -x : x, ay = (y < 0) ? -y : y;
  const int32_t de = (int32_t)(uy & 0x7ff00000u) - (int32_t)(ux & 0x7ff00000u);
  if (de >= 59768832) return (y > 0) ? AVP_AT_HPI : -AVP_AT_HPI;
  if (de <= -59768832) {
    if (x > 0) return avp_copysign(AVP_DIV(ay, ax), y);       /* subnormal-result scaling path not needed here */
    return (y > 0) ? AVP_AT_OPI : -AVP_AT_OPI;
  }
  if (ax < 0x1.0p-500 || ay < 0x1.0p-500) { ax = AVP_MUL(ax, 0x1.0p+500); ay = AVP_MUL(ay, 0x1.0p+500); }
  if (ax > 0x1.0p+500 || ay > 0x1.0p+500) { ax = AVP_MUL(ax, 0x1.0p-500); ay = AVP_MUL(ay, 0x1.0p-500); }
  double u, du;
  if (ay < ax) {
    u = AVP_DIV(ay, ax);
    const double v = AVP_MUL(u, ax), vv = AVP_FMA(u, ax, -v);
    du = AVP_DIV(AVP_SUB(AVP_SUB(ay, v), vv), ax);
  } else {
    u = AVP_DIV(ax, ay);
    const double v = AVP_MUL(u, ay), vv = AVP_FMA(u, ay, -v);
    du = AVP_DIV(AVP_SUB(AVP_SUB(ax, v), vv), ay);
  }
  double z;
  if (x > 0) {
    if (ay < ax) {                                   /* (i) atan(ay/ax) */
      if (u < 0x1.0p-4) {
        const double v = AVP_MUL(u, u);
        const double zz = AVP_FMA(AVP_MUL(u, v), avp_at_poly(v), du);
        z = AVP_ADD(u, zz);
      } else {
        const int i = avp_at_index(u);
        const double t3 = AVP_SUB(u, AVP_CIJ(7 * i));
        const double v = AVP_ADD(du, t3);
        const double dv = (fabs(t3) > fabs(du)) ? AVP_ADD(AVP_SUB(t3, v), du) : AVP_ADD(AVP_SUB(du, v), t3);
        const double t2 = AVP_CIJ(7 * i + 2);
        double p = AVP_FMA(v, AVP_CIJ(7 * i + 6), AVP_CIJ(7 * i + 5));
        p = AVP_FMA(v, p, AVP_CIJ(7 * i + 4));
        p = AVP_FMA(v, p, AVP_CIJ(7 * i + 3));
        const double zz = AVP_FMA(v, t2, AVP_FMA(dv, t2, AVP_MUL(AVP_MUL(v, v), p)));
        z = AVP_ADD(zz, AVP_CIJ(7 * i + 1));
      }
    } else {                                         /* (ii) pi/2 - atan(ax/ay) */
      if (u < 0x1.0p-4) {
        const double v = AVP_MUL(u, u);
        const double zz = AVP_MUL(AVP_MUL(u, v), avp_at_poly(v));
        const double t2 = AVP_SUB(AVP_AT_HPI, u);
        const double cor = (AVP_AT_HPI > fabs(u)) ? AVP_SUB(AVP_SUB(AVP_AT_HPI, t2), u) : AVP_SUB(AVP_AT_HPI, AVP_ADD(u, t2));
        const double t3 = AVP_SUB(AVP_SUB(AVP_ADD(cor, AVP_AT_HPI1), du), zz);
        z = AVP_ADD(t3, t2);
      } else {
        const int i = avp_at_index(u);
        const double v = AVP_ADD(AVP_SUB(u, AVP_CIJ(7 * i)), du);
        const double zz = AVP_FMA(-v, avp_at_tpoly(i, v), AVP_AT_HPI1);
        z = AVP_ADD(AVP_SUB(AVP_AT_HPI, AVP_CIJ(7 * i + 1)), zz);
      }
    }
  } else {
    if (ax < ay) {                                   /* (iii) pi/2 + atan(ax/ay) */
      if (u < 0x1.0p-4) {
        const double v = AVP_MUL(u, u);
        const double zz = AVP_MUL(AVP_MUL(u, v), avp_at_poly(v));
        const double t2 = AVP_ADD(u, AVP_AT_HPI);
        const double cor = (AVP_AT_HPI > fabs(u)) ? AVP_ADD(AVP_SUB(AVP_AT_HPI, t2), u) : AVP_ADD(AVP_SUB(u, t2), AVP_AT_HPI);
        const double t3 = AVP_ADD(AVP_ADD(AVP_ADD(cor, AVP_AT_HPI1), du), zz);
        z = AVP_ADD(t3, t2);
      } else {
        const int i = avp_at_index(u);
        const double v = AVP_ADD(AVP_SUB(u, AVP_CIJ(7 * i)), du);
        const double zz = AVP_FMA(v, avp_at_tpoly(i, v), AVP_AT_HPI1);
        z = AVP_ADD(AVP_ADD(AVP_AT_HPI, AVP_CIJ(7 * i + 1)), zz);
      }
    } else {                                         /* (iv) pi - atan(ay/ax) */
      if (u < 0x1.0p-4) {
        const double v = AVP_MUL(u, u);
        const double zz = AVP_MUL(AVP_MUL(u, v), avp_at_poly(v));
        const double t2 = AVP_SUB(AVP_AT_OPI, u);
        const double cor = (AVP_AT_OPI > fabs(u)) ? AVP_SUB(AVP_SUB(AVP_AT_OPI, t2), u) : AVP_SUB(AVP_AT_OPI, AVP_ADD(u, t2));
        const double t3 = AVP_SUB(AVP_SUB(AVP_ADD(cor, AVP_AT_OPI1), du), zz);
        z = AVP_ADD(t3, t2);
      } else {
        const int i = avp_at_index(u);
        const double v = AVP_ADD(AVP_SUB(u, AVP_CIJ(7 * i)), du);
        const double zz = AVP_FMA(-v, avp_at_tpoly(i, v), AVP_AT_OPI1);
        z = AVP_ADD(AVP_SUB(AVP_AT_OPI, AVP_CIJ(7 * i + 1)), zz);
      }
    }
  }
  return avp_copysign(z, y);
}

/* ---------------------------------------------------------------- asin / acos (e_asin.c) */
#define AVP_HAVE_ASIN 1
#define AVP_AS_F1 0x1.55555555554f9p-3
#define AVP_AS_F2 0x1.333333336127dp-4
#define AVP_AS_F3 0x1.6db6dae42c0e4p-5
#define AVP_AS_F4 0x1.f1c7e04f4ad99p-6
#define AVP_AS_F5 0x1.6e442c822d419p-6
#define AVP_AS_F6 0x1.292d80f453c72p-6
#define AVP_AS_RT0 0x1.fffffffecc1ddp-1
#define AVP_AS_RT1 0x1.fffffff757304p-2
#define AVP_AS_RT2 0x1.800496769c91ap-2
#define AVP_AS_RT3 0x1.4006318d1dab9p-2

AVP_HD double avp_as_fpoly(double z) {   /* (((((f6*z+f5)*z+f4)*z+f3)*z+f2)*z+f1) */
  double p = AVP_FMA(z, AVP_AS_F6, AVP_AS_F5);
  p = AVP_FMA(z, p, AVP_AS_F4);
  p = AVP_FMA(z, p, AVP_AS_F3);
  p = AVP_FMA(z, p, AVP_AS_F2);
  return AVP_FMA(z, p, AVP_AS_F1);
}
/* table intervals 0.125 <= |x| < 0.96875: t = Taylor tail about the sub-interval origin, *last = asin(origin) */
AVP_HD double avp_as_table(uint32_t k, double ax, double *last) {
  int n, d;
  if (k < 0x3fd00000u) { n = 11 * (int)((k >> 15) & 0x1f); d = 5; }
  else if (k < 0x3fe00000u) { n = 11 * (int)((k >> 14) & 0x3f) + 352; d = 5; }
  else if (k < 0x3fe80000u) { n = 1056 + 3 * (int)((k >> 11) & 0x1fc); d = 6; }
  else if (k < 0x3fed8000u) { n = 992 + 13 * (int)((k >> 13) & 0x7f); d = 7; }
  else if (k < 0x3fee8000u) { n = 884 + 14 * (int)((k >> 13) & 0x7f); d = 8; }
  else { n = 768 + 15 * (int)((k >> 13) & 0x7f); d = 9; }
  const double xx = AVP_SUB(ax, AVP_ASN(n));
  double p = AVP_ASN(n + 1 + d);
  for (int j = d; j >= 2; --j) p = AVP_FMA(xx, p, AVP_ASN(n + j));
  p = AVP_FMA(AVP_MUL(xx, xx), p, AVP_ASN(n + 2 + d));
  *last = AVP_ASN(n + 3 + d);
  return AVP_FMA(xx, AVP_ASN(n + 1), p);
}
/* 0.96875 <= |x| < 1: z = (1-|x|)/2, c ~ sqrt(z) by a table seed and one Newton step */
AVP_HD void avp_as_sqrt_core(double z, double *c_out, double *t_out) {
  const uint32_t kz = (uint32_t)(avp_d2u(z) >> 32);
  double t = AVP_MUL(AVP_INROOT((kz & 0x001fffffu) >> 14), AVP_POWTWO(511 - (int)(kz >> 21)));
  const double r = AVP_FMA(-AVP_MUL(t, t), z, 1.0);
  double q = AVP_FMA(r, AVP_AS_RT3, AVP_AS_RT2);
  q = AVP_FMA(r, q, AVP_AS_RT1);
  q = AVP_FMA(r, q, AVP_AS_RT0);
  t = AVP_MUL(q, t);
  *c_out = AVP_MUL(z, t);
  *t_out = t;
}

AVP_HD double avp_asin(double x) {
  const uint32_t m = (uint32_t)(avp_d2u(x) >> 32), k = m & 0x7fffffffu;
  const int pos = (int32_t)m > 0;
  if (k < 0x3e500000u) return x;
  if (k < 0x3fc00000u) {                                  /* |x| < 0.125 */
    const double x2 = AVP_MUL(x, x);
    return AVP_FMA(avp_as_fpoly(x2), AVP_MUL(x, x2), x);
  }
  if (k < 0x3fef0000u) {                                  /* table intervals */
    double last;
    const double t = avp_as_table(k, pos ? x : -x, &last);
    const double res = AVP_ADD(t, last);
    return pos ? res : -res;
  }
  if (k < 0x3ff00000u) {                                  /* 0.96875 <= |x| < 1 */
    const double z = AVP_MUL(pos ? AVP_SUB(1.0, x) : AVP_ADD(x, 1.0), 0.5);
    double c, t;
    avp_as_sqrt_core(z, &c, &t);
    const double w = AVP_FMA(-c, AVP_MUL(t, 0.5), 1.5);   /* 1.5 - 0.5*t*c */
    const double y = AVP_SUB(AVP_ADD(c, 0x1.0p+24), 0x1.0p+24);
    const double ty = AVP_FMA(w, c, y);                   /* t + y */
    const double cc = AVP_DIV(AVP_FMA(-y, y, z), ty);
    const double p = AVP_MUL(avp_as_fpoly(z), z);
    const double s2 = AVP_ADD(AVP_ADD(y, cc), AVP_ADD(y, cc));            /* 2*(y+cc) */
    const double cor = AVP_FMA(-s2, p, AVP_FMA(-2.0, cc, AVP_AT_HPI1));   /* (hp1 - 2cc) - 2(y+cc)p */
    const double res1 = AVP_FMA(-2.0, y, AVP_AT_HPI);
    const double res = AVP_ADD(cor, res1);
    return pos ? res : -res;
  }
  if (k == 0x3ff00000u && (uint32_t)avp_d2u(x) == 0) return pos ? AVP_AT_HPI : -AVP_AT_HPI;
  return asin(x);                                         /* |x| > 1 or NaN */
}

AVP_HD double avp_acos(double x) {
  const uint32_t m = (uint32_t)(avp_d2u(x) >> 32), k = m & 0x7fffffffu;
  const int pos = (int32_t)m > 0;
  if (k < 0x3c880000u) return AVP_AT_HPI;
  if (k < 0x3fc00000u) {                                  /* |x| < 0.125 */
    const double x2 = AVP_MUL(x, x);
    const double p = avp_as_fpoly(x2);
    const double r = AVP_SUB(AVP_AT_HPI, x);
    const double cor = AVP_FMA(-p, AVP_MUL(x, x2), AVP_ADD(AVP_SUB(AVP_SUB(AVP_AT_HPI, r), x), AVP_AT_HPI1));
    return AVP_ADD(r, cor);
  }
  if (k < 0x3fef0000u) {
    double last;
    const double t = avp_as_table(k, pos ? x : -x, &last);
    if (pos) return AVP_ADD(AVP_SUB(AVP_AT_HPI1, t), AVP_SUB(AVP_AT_HPI, last));
    return AVP_ADD(AVP_ADD(t, AVP_AT_HPI1), AVP_ADD(last, AVP_AT_HPI));
  }
  if (k < 0x3ff00000u) {
    const double z = AVP_MUL(pos ? AVP_SUB(1.0, x) : AVP_ADD(x, 1.0), 0.5);
    double c, t;
    avp_as_sqrt_core(z, &c, &t);
    const double w = AVP_FMA(-c, AVP_MUL(t, 0.5), 1.5);
    const double y1 = AVP_FMA(c, 0x1.0p+27, c);           /* t27*c + c */
    const double y = AVP_FMA(-0x1.0p+27, c, y1);
    const double ty = AVP_FMA(w, c, y);
    const double cc = AVP_DIV(AVP_FMA(-y, y, z), ty);
    const double p = AVP_MUL(avp_as_fpoly(z), z);
    const double cor = AVP_MUL(p, AVP_ADD(y, cc));
    if ((int32_t)m >= 0) { const double r = AVP_ADD(AVP_ADD(cc, cor), y); return AVP_ADD(r, r); }
    const double r = AVP_ADD(AVP_SUB(AVP_SUB(AVP_AT_HPI1, cc), cor), AVP_SUB(AVP_AT_HPI, y));
    return AVP_ADD(r, r);
  }
  if (k == 0x3ff00000u && (uint32_t)avp_d2u(x) == 0) return pos ? 0.0 : AVP_AT_OPI;
  return acos(x);
}

/* ---------------------------------------------------------------- pow(x, 2.0) (e_pow.c)
 * The reference squares with `** 2` on Python floats / numpy scalars (rs_curve.py:172,220,
 * hybrid_a_star.py:308), which is libm pow(x, 2.0): exp(2*log|x|) in double-double, NOT x*x
 * (they differ in ~0.09% of arguments). */
#define AVP_HAVE_POW2 1
AVP_HD double avp_pow2(double x) {
  const uint64_t ix = avp_d2u(x) & 0x7fffffffffffffffULL;
  const uint32_t top = (uint32_t)(ix >> 52);
  if (top - 1u >= 0x7feu - 1u + 1u || top < 0x200u || top > 0x5ffu) return AVP_MUL(x, x);   /* 0, subnormal, inf/nan, |x| outside 2^+-511: not on the planner path */
  const uint64_t tmp = ix - 0x3fe6955500000000ULL;
  const int i = (int)((tmp >> 45) & 127);
  const int k = (int)((int64_t)tmp >> 52);
  const double z = avp_u2d(ix - (tmp & 0xfff0000000000000ULL));
  const double kd = (double)k;
  const double invc = AVP_POWLOG(4 * i), logc = AVP_POWLOG(4 * i + 2), logctail = AVP_POWLOG(4 * i + 3);
  const double r = AVP_FMA(z, invc, -1.0);
  const double t1 = AVP_FMA(kd, 0x1.62e42fefa3800p-1, logc);
  const double t2 = AVP_ADD(r, t1);
  const double lo1 = AVP_FMA(kd, 0x1.ef35793c76730p-45, logctail);
  const double lo2 = AVP_ADD(AVP_SUB(t1, t2), r);
  const double ar = AVP_MUL(r, -0.5), ar2 = AVP_MUL(r, ar), ar3 = AVP_MUL(r, ar2);
  const double hi = AVP_ADD(t2, ar2);
  const double lo3 = AVP_FMA(ar, r, -ar2);
  const double lo4 = AVP_ADD(AVP_SUB(t2, hi), ar2);
  const double a12 = AVP_FMA(r, 0x1.0000000000006p-1, -0x1.5555555555560p-1);
  const double a34 = AVP_FMA(r, -0x1.555555529a47ap-1, 0x1.999999959554ep-1);
  const double a56 = AVP_FMA(r, 0x1.0002b8b263fc3p+0, -0x1.2495b9b4845e9p+0);
  const double q = AVP_FMA(ar2, AVP_FMA(a56, ar2, a34), a12);
  const double lo = AVP_FMA(ar3, q, AVP_ADD(AVP_ADD(AVP_ADD(lo1, lo2), lo3), lo4));
  const double lhi = AVP_ADD(hi, lo);
  const double ltail = AVP_ADD(AVP_SUB(hi, lhi), lo);
  const double ehi = AVP_MUL(2.0, lhi);
  const double elo = AVP_FMA(2.0, ltail, AVP_FMA(lhi, 2.0, -ehi));
  const uint32_t abstop = (uint32_t)(avp_d2u(ehi) >> 52) & 0x7ffu;
  if (abstop < 0x3c9u) return AVP_ADD(ehi, 1.0);                    /* |2 log x| < 2^-54 */
  if (abstop > 0x3c9u + 0x3eu) return AVP_MUL(x, x);                /* overflow/underflow handling: not reached for 2^-511 < |x| < 2^511 */
  const double kdz = AVP_FMA(ehi, 0x1.71547652b82fep+7, 0x1.8p+52);
  const uint64_t ki = avp_d2u(kdz);
  const double kde = AVP_SUB(kdz, 0x1.8p+52);
  double rr = AVP_FMA(kde, -0x1.cf79abc9e3b3ap-47, AVP_FMA(kde, -0x1.62e42fefa0000p-8, ehi));
  rr = AVP_ADD(elo, rr);
  const int idx = 2 * (int)(ki & 127);
  const double tail = avp_u2d(AVP_EXPTAB(idx));
  const uint64_t sbits = AVP_EXPTAB(idx + 1) + (ki << 45);
  const double r2 = AVP_MUL(rr, rr);
  const double c23 = AVP_FMA(rr, 0x1.555555555543cp-3, 0x1.ffffffffffdbdp-2);
  const double c45 = AVP_FMA(rr, 0x1.1111167a4d017p-7, 0x1.55555cf172b91p-5);
  const double tmp2 = AVP_FMA(c45, AVP_MUL(r2, r2), AVP_FMA(c23, r2, AVP_ADD(rr, tail)));
  const double scale = avp_u2d(sbits);
  return AVP_FMA(tmp2, scale, scale);
}

/* ---------------------------------------------------------------- tan (s_tan.c), |x| <= 25 */
#define AVP_HAVE_TAN 1
AVP_HD double avp_tn_poly(double x2) {   /* d3 + x2*(d5 + x2*(d7 + x2*(d9 + x2*d11))) */
  double t = AVP_FMA(x2, 0x1.2385a3cf2e4eap-7, 0x1.664ed49cfc666p-6);
  t = AVP_FMA(x2, t, 0x1.ba1ba1cdb8745p-5);
  t = AVP_FMA(x2, t, 0x1.11111111107c6p-3);
  return AVP_FMA(x2, t, 0x1.5555555555555p-2);
}
AVP_HD double avp_tan(double x) {
  const uint32_t ux = (uint32_t)(avp_d2u(x) >> 32);
  if ((ux & 0x7ff00000u) == 0x7ff00000u) return tan(x);
  const double w = (x < 0.0) ? -x : x;
  if (w <= 0x1.b096c00000000p-27) return x;
  if (w <= 0x1.f212d00000000p-5) {                       /* (II) |x| <= 0.0608 */
    const double x2 = AVP_MUL(x, x);
    return AVP_FMA(AVP_MUL(x, x2), avp_tn_poly(x2), x);
  }
  if (w <= 0x1.92f1a00000000p-1) {                       /* (III) |x| <= 0.787 */
    const int i = (int)AVP_FMA(w, 256.0, -15.5);
    const double z = AVP_SUB(w, AVP_XFG(4 * i)), z2 = AVP_MUL(z, z);
    const double pz = AVP_FMA(AVP_MUL(z, z2), AVP_FMA(z2, 0x1.11112e0a6b45fp-3, 0x1.5555555554dbdp-2), z);
    const double fi = AVP_XFG(4 * i + 1), gi = AVP_XFG(4 * i + 2);
    const double t2 = AVP_DIV(AVP_MUL(AVP_ADD(fi, gi), pz), AVP_SUB(gi, pz));
    const double y = AVP_ADD(t2, fi);
    return AVP_MUL(y, (x < 0.0) ? -1.0 : 1.0);
  }
  if (w <= 25.0) {                                       /* (IV) reduce by pi/2 */
    const double t = AVP_FMA(x, AVP_SC_HPINV, AVP_SC_TOINT);
    const double xn = AVP_SUB(t, AVP_SC_TOINT);
    const int n = (int)((uint32_t)avp_d2u(t) & 1u);
    const double t1 = AVP_FMA(-xn, AVP_SC_MP2, AVP_FMA(-xn, AVP_SC_MP1, x));
    const double a = AVP_FMA(-xn, -0x1.cb3b399d747f2p-55, t1);
    const double da = AVP_FMA(-xn, -0x1.cb3b399d747f2p-55, AVP_SUB(t1, a));
    double ya, yya, sy;
    if (a < 0.0) { ya = -a; yya = -da; sy = -1.0; } else { ya = a; yya = da; sy = 1.0; }
    if (ya <= 0x1.f212d00000000p-5) {                    /* (VII) */
      const double a2 = AVP_MUL(a, a);
      const double t2 = AVP_FMA(AVP_MUL(a, a2), avp_tn_poly(a2), da);
      const double b = AVP_ADD(a, t2);
      if (!n) return b;
      const double db = (fabs(a) > fabs(t2)) ? AVP_ADD(AVP_SUB(a, b), t2) : AVP_ADD(AVP_SUB(t2, b), a);
      const double c = AVP_DIV(1.0, b);
      const double u = AVP_MUL(c, b), uu = AVP_FMA(c, b, -u);
      const double cc = AVP_DIV(AVP_FMA(-db, c, AVP_ADD(AVP_SUB(AVP_SUB(1.0, u), uu), 0.0)), b);
      const double z = AVP_ADD(c, cc), zz = AVP_ADD(AVP_SUB(c, z), cc);
      return -AVP_ADD(zz, z);
    }
    const int i = (int)AVP_FMA(ya, 256.0, -15.5);         /* (VIII) */
    const double z = AVP_ADD(AVP_SUB(ya, AVP_XFG(4 * i)), yya), z2 = AVP_MUL(z, z);
    const double pz = AVP_FMA(AVP_MUL(z, z2), AVP_FMA(z2, 0x1.11112e0a6b45fp-3, 0x1.5555555554dbdp-2), z);
    const double fi = AVP_XFG(4 * i + 1), gi = AVP_XFG(4 * i + 2);
    const double num = AVP_MUL(AVP_ADD(fi, gi), pz);
    if (n) return AVP_MUL(AVP_SUB(gi, AVP_DIV(num, AVP_ADD(pz, fi))), -sy);
    return AVP_MUL(AVP_ADD(AVP_DIV(num, AVP_SUB(gi, pz)), fi), sy);
  }
  return tan(x);
}

#endif /* AVP_LIBM_H */
