/* Host check: avp_sin/avp_cos (csrc/avp_sincos.h) vs the installed libm, bit for bit.
 * usage: sincos_bits <n_random> <seed>   -> prints mismatch counts, exit 1 if any. */
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
#include <math.h>
#include "../../automatedvaletparking_b200/csrc/avp_sincos.h"

static uint64_t s[2];
static uint64_t rnd(void) { /* xorshift128+ */
  uint64_t x = s[0], y = s[1]; s[0] = y; x ^= x << 23; s[1] = x ^ y ^ (x >> 17) ^ (y >> 26); return s[1] + y;
}
static double u01(void) { return (rnd() >> 11) * 0x1.0p-53; }

int main(int argc, char **argv) {
  long n = argc > 1 ? atol(argv[1]) : 10000000; s[0] = argc > 2 ? strtoull(argv[2], 0, 10) : 12345; s[1] = 0x9E3779B97F4A7C15ULL ^ s[0];
  long bs = 0, bc = 0, tot = 0;
  double ranges[][2] = {{-3.2, 3.2}, {-7, 7}, {-1, 1}, {-0.13, 0.13}, {0.8, 0.9}, {2.3, 2.5}, {-100, 100}, {-1e6, 1e6}, {-1e8, 1e8}};
  int nr = sizeof(ranges) / sizeof(ranges[0]);
  for (int r = 0; r < nr; ++r) {
    for (long i = 0; i < n; ++i) {
      double x = ranges[r][0] + (ranges[r][1] - ranges[r][0]) * u01();
      volatile double a = sin(x), b = cos(x);
      if (avp_d2u(avp_sin(x)) != avp_d2u(a)) { if (bs < 5) printf("sin mismatch x=%a libm=%a avp=%a\n", x, a, avp_sin(x)); bs++; }
      if (avp_d2u(avp_cos(x)) != avp_d2u(b)) { if (bc < 5) printf("cos mismatch x=%a libm=%a avp=%a\n", x, b, avp_cos(x)); bc++; }
      tot++;
    }
  }
  /* tiny / special magnitudes */
  for (int e = -60; e < 27; ++e) for (int j = 0; j < 2000; ++j) {
    double x = ldexp(0.5 + 0.5 * u01(), e) * ((j & 1) ? -1 : 1);
    volatile double a = sin(x), b = cos(x);
    if (avp_d2u(avp_sin(x)) != avp_d2u(a)) bs++;
    if (avp_d2u(avp_cos(x)) != avp_d2u(b)) bc++;
    tot++;
  }
  printf("checked %ld args: sin mismatches %ld, cos mismatches %ld\n", tot, bs, bc);
  return (bs || bc) ? 1 : 0;
}
