/* oracle/check_voxel_wide.c -- TEST INFRASTRUCTURE.
 * csrc/pm_math.cuh voxel_x_wide / voxel_z_wide exist in two forms: the round-1 form (floor OR ceil of u, then a truncating division
 * by 3 with a sign case) and the one-conversion form the render kernel uses now (floor of |.| on the side of the truncation point,
 * one multiply-shift, sign restored).  This program evaluates both, as host C with the same operations, for every float bit pattern
 * p = 0 .. 2^32-1 in steps of argv[1] (default 1 = exhaustive) and also compares them with the literal double form of the
 * reference (PMK:269-274 inverted: trunc(((double)p + 1.5) / 3 * 32) and trunc((double)p / 6 * 32), clamped to [-3, 36], NaN -> 0).
 * Prints "<mismatches old/new> <mismatches new/literal> <patterns>" (expected: 0 0 N). */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static int div3_trunc(int n) { return n >= 0 ? (n * 43691) >> 17 : -((-n * 43691) >> 17); }
static int old_x(float p) {
  float u = fmaf(32.0f, p, 0x1p-48f);
  u = u != u ? -48.0f : fminf(fmaxf(u, -57.0f), 60.0f);
  return div3_trunc(u >= -48.0f ? (int)floorf(u) + 48 : (int)ceilf(u) + 48);
}
static int old_z(float p) {
  float u = 16.0f * p;
  u = u != u ? 0.0f : fminf(fmaxf(u, -9.0f), 108.0f);
  return div3_trunc(u >= 0.0f ? (int)floorf(u) : (int)ceilf(u));
}
static int new_x(float p) {
  float u = fmaf(32.0f, p, 0x1p-48f);
  u = u != u ? -48.0f : fminf(fmaxf(u, -57.0f), 60.0f);
  const int neg = u < -48.0f;
  const int f = (int)floorf(neg ? -u : u);
  const int q = ((neg ? f - 48 : f + 48) * 43691) >> 17;
  return neg ? -q : q;
}
static int new_z(float p) {
  float u = 16.0f * p;
  u = u != u ? 0.0f : fminf(fmaxf(u, -9.0f), 108.0f);
  const int neg = u < 0.0f;
  const int q = ((int)floorf(neg ? -u : u) * 43691) >> 17;
  return neg ? -q : q;
}
static int clampi(double v) { return v != v ? 0 : (int)(v < -3.0 ? -3.0 : v > 36.0 ? 36.0 : trunc(v)); }

int main(int argc, char **argv) {
  long long stride = argc > 1 ? atoll(argv[1]) : 1;
  if (stride < 1) stride = 1;
  long long bad = 0, bad_lit = 0, n = 0;
#pragma omp parallel for reduction(+ : bad, bad_lit, n)
  for (long long i = 0; i <= 0xffffffffLL; i += stride) {
    uint32_t b = (uint32_t)i;
    float p;
    memcpy(&p, &b, 4);
    const int nx = new_x(p), nz = new_z(p);
    bad += (old_x(p) != nx) + (old_z(p) != nz);
    bad_lit += (clampi(((double)p + 1.5) / 3.0 * 32.0) != nx) + (clampi((double)p / 6.0 * 32.0) != nz);
    n++;
  }
  printf("%lld %lld %lld\n", bad, bad_lit, n);
  return bad != 0 || bad_lit != 0;
}
