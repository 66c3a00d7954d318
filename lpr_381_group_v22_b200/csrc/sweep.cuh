// sweep.cuh -- the rank-1 elimination sweep body shared by the tableau and B&B kernels.
#pragma once
#include "select.cuh"

namespace lpr {

// Programmatic dependent launch (sm_90+): every kernel of a pivot chain first waits for its
// predecessor's results, then lets its successor start launching.  Without the launch attribute
// both instructions are no-ops.
__device__ __forceinline__ void pdl_wait_then_release() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

// streaming 128-bit tableau load that does not allocate in L1, so the pivot row and the factor column
// (re-read by every thread through the read-only path) stay L1 resident (tools/sweep_bench.cu: +5%)
__device__ __forceinline__ double2 ld_stream(const double2* p) {
  double2 r;
  asm volatile("ld.global.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
  return r;
}

// L1-cached loads with ordinary coherence rules (ld.global.ca): a line is dropped by the acquire fence of a grid
// barrier, unlike the non-coherent read-only path (__ldg / const __restrict__), which is only refreshed at kernel
// boundaries
__device__ __forceinline__ double ld_ca(const double* p) {
  double r;
  asm volatile("ld.global.ca.f64 %0, [%1];" : "=d"(r) : "l"(p));
  return r;
}
__device__ __forceinline__ double2 ld_ca2(const double2* p) {
  double2 r;
  asm volatile("ld.global.ca.v2.f64 {%0,%1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
  return r;
}

// one TILE = kSweepThreads * UNROLL 16-byte chunks of the flat R x ld/2 array, tile index t.
// COHERENT: the factor column and the pivot row are re-staged by OTHER CTAs of the same (persistent) kernel between
// calls, so they must not come through the non-coherent read-only path (__ldg): ld.global.ca instead.
template <int SKIP, bool OOP, bool EMIT, int UNROLL, bool COHERENT = false>
__device__ __forceinline__ void sweep_tile(const double2* __restrict__ src, double2* __restrict__ dst,
                                           const double* __restrict__ f, const double2* __restrict__ prow2,
                                           double* __restrict__ cn, double* __restrict__ rhsb,
                                           unsigned long long n, unsigned ldv, int p, unsigned long long t,
                                           double eps, unsigned rhs_chunk, int rhs_odd, unsigned e_chunk, int e_odd,
                                           unsigned tix = threadIdx.x) {
  // tix: this thread's position among the kSweepThreads threads working on tile t (a wider CTA runs several tiles)
  constexpr unsigned TILE = kSweepThreads * UNROLL;
  const unsigned long long q0 = t * TILE + tix;
  unsigned row = (unsigned)(q0 / ldv);
  unsigned c = (unsigned)(q0 - (unsigned long long)row * ldv);
  double2 x[UNROLL];
  double fv[UNROLL];
  unsigned rw[UNROLL], cc[UNROLL];
  bool act[UNROLL];
#pragma unroll
  for (int k = 0; k < UNROLL; k++) {
    const unsigned long long q = q0 + (unsigned long long)k * kSweepThreads;
    rw[k] = row;
    cc[k] = c;
    act[k] = q < n;
    if (act[k]) {
      if (SKIP != 0) {  // the skip decision needs f before the load is issued
        fv[k] = COHERENT ? ld_ca(f + row) : __ldg(f + row);
        bool sk = ((int)row != p) && ((SKIP == 1) ? (fabs(fv[k]) <= eps) : (fabs(fv[k]) < eps));
        if (sk && !OOP) act[k] = false;
      }
      if (act[k]) x[k] = ld_stream(src + q);
    }
    c += kSweepThreads;  // ldv may be < 256: wrap as often as needed
    while (c >= ldv) { c -= ldv; row++; }
  }
#pragma unroll
  for (int k = 0; k < UNROLL; k++) {
    if (!act[k]) continue;
    const unsigned long long q = q0 + (unsigned long long)k * kSweepThreads;
    if (SKIP == 0) fv[k] = COHERENT ? ld_ca(f + rw[k]) : __ldg(f + rw[k]);  // L1-resident: loaded late (registers)
    const double2 pr = COHERENT ? ld_ca2(prow2 + cc[k]) : __ldg(prow2 + cc[k]);
    double2 y;
    bool sk = false;
    if (SKIP != 0 && OOP && (int)rw[k] != p) sk = (SKIP == 1) ? (fabs(fv[k]) <= eps) : (fabs(fv[k]) < eps);
    if ((int)rw[k] == p) {
      y = pr;
    } else if (sk) {
      y = x[k];
    } else {
      y.x = __dsub_rn(x[k].x, __dmul_rn(fv[k], pr.x));
      y.y = __dsub_rn(x[k].y, __dmul_rn(fv[k], pr.y));
      if (OOP) {
        if (y.x == 0.0) y.x = 0.0;
        if (y.y == 0.0) y.y = 0.0;
      }
    }
    dst[q] = y;
    if (EMIT) {
      if (cc[k] == e_chunk) cn[rw[k]] = e_odd ? y.y : y.x;
      if (cc[k] == rhs_chunk) rhsb[rw[k]] = rhs_odd ? y.y : y.x;
    }
  }
}

// the tiles of one tableau dealt to the CTAs of the grid in contiguous spans
template <int SKIP, bool OOP, bool EMIT, int UNROLL>
__device__ __forceinline__ void sweep_body(const double2* __restrict__ src, double2* __restrict__ dst,
                                           const double* __restrict__ f, const double2* __restrict__ prow2,
                                           double* __restrict__ cn, double* __restrict__ rhsb, int R, int C,
                                           int ld, int p, int e_next, double eps, int reverse) {
  const unsigned ldv = (unsigned)(ld >> 1);
  const unsigned long long n = (unsigned long long)R * ldv;
  constexpr unsigned TILE = kSweepThreads * UNROLL;
  const unsigned long long ntiles = (n + TILE - 1) / TILE;
  const unsigned long long t0 = ntiles * blockIdx.x / gridDim.x, t1 = ntiles * (blockIdx.x + 1) / gridDim.x;
  const unsigned rhs_chunk = (unsigned)((C - 1) >> 1);
  const int rhs_odd = (C - 1) & 1;
  const unsigned e_chunk = e_next >= 0 ? (unsigned)(e_next >> 1) : 0xffffffffu;
  const int e_odd = e_next & 1;
  for (unsigned long long tt = t0; tt < t1; tt++) {
    const unsigned long long t = reverse ? (ntiles - 1 - tt) : tt;  // full mirror: last tiles first
    sweep_tile<SKIP, OOP, EMIT, UNROLL>(src, dst, f, prow2, cn, rhsb, n, ldv, p, t, eps, rhs_chunk, rhs_odd,
                                        e_chunk, e_odd);
  }
}


}  // namespace lpr
