/* Exhaustive check of the division-free Math.Round(x,4) tail used by the B&B kernels (common.cuh
 * net_round4_fast): for every integer-valued double k with |k| <= 2^31,
 *   q0 = k * RN(1e-4);  r = fma(-1e4, q0, k);  q1 = fma(r, RN(1e-4), q0)
 * equals the correctly rounded k / 1e4 (Markstein's correction step), and rounding twice is the
 * same as rounding once:  rint((k / 1e4) * 1e4) == k.
 * Build: gcc -O2 -ffp-contract=off -fopenmp tools/div1e4_check.c -lm -o /tmp/div1e4_check      */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

int main(int argc, char** argv) {
  const long long lim = argc > 1 ? atoll(argv[1]) : (1LL << 31);
  const long long from = argc > 2 ? atoll(argv[2]) : 0; /* optional start of the range */
  const double inv = 1e-4;
  long long bad_div = 0, bad_idem = 0;
#pragma omp parallel for reduction(+ : bad_div, bad_idem) schedule(static)
  for (long long i = from; i <= lim; i++) {
    for (int s = 0; s < 2; s++) {
      const double k = s ? -(double)i : (double)i;
      const double ref = k / 1e4;
      double q = k;
      if (k != 0.0) {
        const double q0 = k * inv;
        const double r = fma(-1e4, q0, k);
        q = fma(r, inv, q0);
      }
      if (q != ref || signbit(q) != signbit(ref)) bad_div++;
      if (rint(ref * 1e4) != k) bad_idem++;
    }
  }
  printf("checked %lld <= |k| <= %lld: division mismatches %lld, idempotence mismatches %lld\n", from, lim, bad_div, bad_idem);
  return (bad_div || bad_idem) ? 1 : 0;
}
