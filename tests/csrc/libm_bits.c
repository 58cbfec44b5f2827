/* Host check: csrc/avp_libm.h (same source the device compiles) vs the installed libm, bit for bit.
 * usage: libm_bits <n_random> <seed> [func]   -> mismatch counts, exit 1 if any. */
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
#include <string.h>
#include <math.h>
#include "../../automatedvaletparking_b200/csrc/avp_libm.h"

static uint64_t s[2];
static uint64_t rnd(void) { uint64_t x = s[0], y = s[1]; s[0] = y; x ^= x << 23; s[1] = x ^ y ^ (x >> 17) ^ (y >> 26); return s[1] + y; }
static double u01(void) { return (rnd() >> 11) * 0x1.0p-53; }
static double uni(double a, double b) { return a + (b - a) * u01(); }
static double (*volatile p_atan2)(double, double) = atan2;
#ifdef AVP_HAVE_ASIN
static double (*volatile p_asin)(double) = asin;
static double (*volatile p_acos)(double) = acos;
#endif
#ifdef AVP_HAVE_TAN
static double (*volatile p_tan)(double) = tan;
#endif
#ifdef AVP_HAVE_POW2
static double (*volatile p_pow)(double, double) = pow;
#endif

int main(int argc, char **argv) {
  long n = argc > 1 ? atol(argv[1]) : 10000000; s[0] = argc > 2 ? strtoull(argv[2], 0, 10) : 12345; s[1] = 0x9E3779B97F4A7C15ULL ^ s[0];
  const char *only = argc > 3 ? argv[3] : "";
  long bad = 0, tot = 0;
  if (!*only || !strcmp(only, "atan2")) {
    long b = 0;
    double R[][4] = {{-10, 10, -10, 10}, {-1, 1, -1, 1}, {-1e-3, 1e-3, -50, 50}, {-50, 50, -1e-3, 1e-3}, {-4, 4, -0.3, 0.3}, {0, 3, 1.9, 2.1}, {-1e6, 1e6, -1e6, 1e6}};
    for (unsigned r = 0; r < sizeof(R) / sizeof(R[0]); ++r)
      for (long i = 0; i < n; ++i) {
        double y = uni(R[r][0], R[r][1]), x = uni(R[r][2], R[r][3]);
        if (i % 97 == 0) x = 2.0; if (i % 89 == 0) x = -2.0; if (i % 1013 == 0) y = 0.0; if (i % 1019 == 0) x = 0.0;
        volatile double a = p_atan2(y, x); double g = avp_atan2(y, x);
        if (avp_d2u(a) != avp_d2u(g)) { if (b < 5) printf("atan2 mismatch y=%a x=%a libm=%a avp=%a\n", y, x, a, g); b++; }
        tot++;
      }
    printf("atan2: %ld mismatches\n", b); bad += b;
  }
#ifdef AVP_HAVE_ASIN
  if (!*only || !strcmp(only, "asin")) {
    long b = 0, c = 0;
    for (long i = 0; i < 6 * n; ++i) {
      double x = (i % 3 == 0) ? uni(-1, 1) : (i % 3 == 1 ? uni(-0.13, 0.13) : copysign(1.0 - ldexp(u01(), -(int)(rnd() % 30)), uni(-1, 1)));
      volatile double a = p_asin(x), a2 = p_acos(x);
      if (avp_d2u(a) != avp_d2u(avp_asin(x))) { if (b < 5) printf("asin mismatch x=%a libm=%a avp=%a\n", x, a, avp_asin(x)); b++; }
      if (avp_d2u(a2) != avp_d2u(avp_acos(x))) { if (c < 5) printf("acos mismatch x=%a libm=%a avp=%a\n", x, a2, avp_acos(x)); c++; }
      tot++;
    }
    printf("asin: %ld mismatches, acos: %ld mismatches\n", b, c); bad += b + c;
  }
#endif
#ifdef AVP_HAVE_TAN
  if (!*only || !strcmp(only, "tan")) {
    long b = 0;
    double R[][2] = {{-3.2, 3.2}, {-1.6, 1.6}, {-0.1, 0.1}, {1.5, 1.65}, {-7, 7}, {-100, 100}};
    for (unsigned r = 0; r < sizeof(R) / sizeof(R[0]); ++r)
      for (long i = 0; i < n; ++i) {
        double x = uni(R[r][0], R[r][1]);
        volatile double a = p_tan(x);
        if (avp_d2u(a) != avp_d2u(avp_tan(x))) { if (b < 5) printf("tan mismatch x=%a libm=%a avp=%a\n", x, a, avp_tan(x)); b++; }
        tot++;
      }
    printf("tan: %ld mismatches\n", b); bad += b;
  }
#endif
#ifdef AVP_HAVE_POW2
  if (!*only || !strcmp(only, "pow2")) {
    long b = 0;
    double R[][2] = {{-40, 40}, {-3, 3}, {-1e-3, 1e-3}, {0, 1e4}, {-1e10, 1e10}};
    for (unsigned r = 0; r < sizeof(R) / sizeof(R[0]); ++r)
      for (long i = 0; i < n; ++i) {
        double x = uni(R[r][0], R[r][1]);
        volatile double a = p_pow(x, 2.0);
        if (avp_d2u(a) != avp_d2u(avp_pow2(x))) { if (b < 5) printf("pow2 mismatch x=%a libm=%a avp=%a\n", x, a, avp_pow2(x)); b++; }
        tot++;
      }
    printf("pow(x,2): %ld mismatches\n", b); bad += b;
  }
#endif
  printf("checked %ld calls, %ld mismatches\n", tot, bad);
  return bad ? 1 : 0;
}
