// blocked_apply.cuh -- the per-element update of the delayed-update sweeps (tableau_blocked.cu,
// tableau_pipelined.cu): s pending rank-1 updates applied in the original pivot order with the reference's
// separate multiply / subtract roundings (PrimalSimplexSolver.cs:193-211).
#pragma once
#include "sweep.cuh"

namespace lpr {

// fast path: a full group (s == KM) on a row that is not one of the pending pivot rows -- 2 DMUL + 2 DADD
// per pending pivot and chunk, nothing else
template <int KM>
__device__ __forceinline__ double2 blk_apply_fast(double2 x, const double2* pr, const double2* fr) {
  double2 fq[KM / 2];
#pragma unroll
  for (int h2 = 0; h2 < KM / 2; h2++) fq[h2] = fr[h2];
#pragma unroll
  for (int u = 0; u < KM; u++) {
    const double f = (u & 1) ? fq[u >> 1].y : fq[u >> 1].x;
    x.x = __dsub_rn(x.x, __dmul_rn(f, pr[u].x));
    x.y = __dsub_rn(x.y, __dmul_rn(f, pr[u].y));
  }
  return x;
}
// two rows at once: four independent dependency chains per thread keep the FP64 pipe busy while a DADD waits
// for its predecessor (one row alone is 2 chains of KM dependent subtractions)
template <int KM>
__device__ __forceinline__ void blk_apply_fast2(double2& xa, double2& xb, const double2* pr, const double2* fra,
                                                const double2* frb) {
#pragma unroll
  for (int h2 = 0; h2 < KM / 2; h2++) {
    const double2 fa = fra[h2], fb = frb[h2];
    const double2 p0 = pr[2 * h2], p1 = pr[2 * h2 + 1];
    const double ma = __dmul_rn(fa.x, p0.x), mb = __dmul_rn(fa.x, p0.y);
    const double mc = __dmul_rn(fb.x, p0.x), md = __dmul_rn(fb.x, p0.y);
    xa.x = __dsub_rn(xa.x, ma);
    xa.y = __dsub_rn(xa.y, mb);
    xb.x = __dsub_rn(xb.x, mc);
    xb.y = __dsub_rn(xb.y, md);
    const double na = __dmul_rn(fa.y, p1.x), nb = __dmul_rn(fa.y, p1.y);
    const double nc = __dmul_rn(fb.y, p1.x), nd = __dmul_rn(fb.y, p1.y);
    xa.x = __dsub_rn(xa.x, na);
    xa.y = __dsub_rn(xa.y, nb);
    xb.x = __dsub_rn(xb.x, nc);
    xb.y = __dsub_rn(xb.y, nd);
  }
}
template <int KM>
__device__ __forceinline__ double2 blk_apply_gen(double2 x, const double2* pr, const double2* fr, int row, int s,
                                                 const int* pu) {
  double2 fq[KM / 2];
#pragma unroll
  for (int h2 = 0; h2 < KM / 2; h2++) fq[h2] = fr[h2];
#pragma unroll
  for (int u = 0; u < KM; u++) {
    const double f = (u & 1) ? fq[u >> 1].y : fq[u >> 1].x;
    double2 y;
    y.x = __dsub_rn(x.x, __dmul_rn(f, pr[u].x));
    y.y = __dsub_rn(x.y, __dmul_rn(f, pr[u].y));
    if (row == pu[u]) y = pr[u];
    if (u < s) x = y;
  }
  return x;
}

}  // namespace lpr
