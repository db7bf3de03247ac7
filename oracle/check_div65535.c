/* oracle/check_div65535.c -- TEST INFRASTRUCTURE.
 * randFloat (PMK:1039-1052) divides (float)(int)u by 65535.  The CUDA kernels replace the IEEE division sequence by
 *   q0 = RN(x c), r = fma(-q0, 65535, x), q = fma(r, c, q0),  c = RN(1/65535)
 * This program compares that with the correctly rounded x / 65535.0f for x = (float)(int)i, i = INT_MIN .. INT_MAX in steps
 * of argv[1] (default 1 = exhaustive, ~40 CPU-seconds), and prints the number of mismatches (expected: 0). */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

int main(int argc, char **argv) {
  long long stride = argc > 1 ? atoll(argv[1]) : 1;
  if (stride < 1) stride = 1;
  const float c = 1.0f / 65535.0f;
  long long bad = 0, n = 0;
#pragma omp parallel for reduction(+ : bad, n)
  for (long long i = -2147483648LL; i <= 2147483647LL; i += stride) {
    float x = (float)(int)i;
    volatile float a = x / 65535.0f;
    float q0 = x * c, r = fmaf(-q0, 65535.0f, x), q = fmaf(r, c, q0);
    bad += q != a;
    n++;
  }
  printf("%lld %lld\n", bad, n);
  return bad != 0;
}
