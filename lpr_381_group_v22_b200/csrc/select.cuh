// select.cuh -- block-wide selection primitives shared by the tableau / B&B / revised kernels.
#pragma once
#include "common.cuh"

namespace lpr {

constexpr int kSelThreads = 1024;
constexpr int kSweepThreads = 256;
constexpr double kPosInf = __builtin_huge_val();

#define TAT(T, ld, i, j) (T)[(size_t)(i) * (size_t)(ld) + (size_t)(j)]

// Sequential "running best with hysteresis" scan (accept k iff val_k < best - eps, best starts at
// b0) evaluated in parallel: the first index of the minimum is the answer unless an earlier
// candidate could have blocked it, in which case thread 0 replays the scan literally.  Used for
// PrimalSimplexSolver2.cs:102-141, DualSimplex.cs:27-70, SensitivityAnalyzer.cs:139-196 and
// RevisedPrimalSimplexSolver.cs:104-121.  All threads of the block must call it.
template <class Cand>
__device__ int block_hyst_min(int n, Cand cand, double b0, double eps, MinIdx* sm, int* smi) {
  __shared__ int sh_res;
  MinIdx m = minidx_identity();
  for (int k = threadIdx.x; k < n; k += blockDim.x) {
    double val;
    if (cand(k, val) && val == val) m = minidx_combine(m, MinIdx{val, k});
  }
  m = block_minidx(m, sm);
  if (m.i == INT_MAX) return -1;
  if (!(m.v < __dsub_rn(b0, eps))) return -1;
  int bad = 0;
  for (int k = threadIdx.x; k < m.i; k += blockDim.x) {
    double val;
    if (cand(k, val) && val == val && !(m.v < __dsub_rn(val, eps))) bad++;
  }
  bad = block_sum_int(bad, smi);
  if (bad == 0) return m.i;
  // literal replay of the sequential scan, 32 candidates at a time by warp 0: inside a chunk the next element the scan
  // accepts is the first lane (at or after the last accepted one) whose value beats the current best by more than eps
  // -- earlier lanes were tested against the same best and rejected -- so one ballot finds it; acceptances are rare
  // (each lowers the best by more than eps), a chunk without one costs a single ballot.  Integer-valued tableaux tie
  // all the time (cfg5: nearly every pivot), which made the single-thread replay the most expensive step of a pivot.
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    double best = b0;
    int idx = -1;
    for (int base = 0; base < n; base += 32) {
      const int k = base + lane;
      double val = 0.0;
      const bool ok = k < n && cand(k, val);
      int from = 0;
      while (true) {
        const bool acc = ok && lane >= from && val < __dsub_rn(best, eps);
        const unsigned mask = __ballot_sync(0xffffffffu, acc);
        if (!mask) break;
        const int l = __ffs(mask) - 1;
        best = __shfl_sync(0xffffffffu, val, l);
        idx = base + l;
        from = l + 1;
      }
    }
    if (lane == 0) sh_res = idx;
  }
  __syncthreads();
  int r = sh_res;
  __syncthreads();
  return r;
}

// The same sequential scan (accept k iff val_k < fl(best - eps)), evaluated by a prefix-minimum scan instead of a
// reduction followed by a replay.  With m_k = min of the valid values before k:
//   * best_k >= min(b0, m_k), so val_k < fl(min(b0, m_k) - eps) is accepted whatever happened before ("sure");
//   * fl(best_k - eps) <= min(fl(b0 - eps), m_k) (every earlier value was either accepted, so best <= it, or rejected,
//     so fl(best - eps) <= it; best never rises), so val_k >= that minimum is rejected whatever happened before.
// What is left -- a new prefix minimum by less than eps -- is rare (near-ties of rationals).  The scan's state after
// the LAST sure acceptance k* is that element, everything behind it that is not "unsure" is rejected, so the answer
// is k* unless unsure elements follow it, in which case the span [first unsure after k*, last unsure] is replayed
// literally.  Only the first kScanThreads threads scan (a thread owns a contiguous run of candidates: a serial pass
// for the local minimum, one warp scan, a second serial pass to classify): a block-wide instruction costs 8 issue
// cycles per scheduler at 1024 threads, which made the all-threads form slower than the replay it replaced.
// All threads of the block must call it; blockDim.x >= kScanThreads.
constexpr int kScanThreads = 256;
// minimum of two non-NaN doubles: fmin() costs ten instructions for its NaN rules, this is a compare and two selects
__device__ __forceinline__ double dmin(double x, double y) { return y < x ? y : x; }
__device__ __forceinline__ void scan_threads_barrier() {  // named barrier 1: the scanning warps only
  asm volatile("bar.sync 1, %0;" ::"n"(kScanThreads) : "memory");
}
// the scan proper, executed by threads 0 .. kScanThreads-1 ONLY (the other warps of the CTA are free to do something
// else meanwhile: k_persist's barrier warp drains its stores); every scanning thread returns the answer
template <class Cand>
__device__ int hyst_min_scan_threads(int n, Cand cand, double b0, double eps) {
  __shared__ double s_tot[kScanThreads / 32];
  __shared__ int s_kstar, s_ulo, s_uhi;
  constexpr unsigned kFull = 0xffffffffu;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int per = ((n + kScanThreads - 1) / kScanThreads) | 1;  // odd: fewer shared-memory bank conflicts
  const int k0 = tid * per, k1 = min(n, k0 + per);
  if (tid == 0) {
    s_kstar = -1;
    s_ulo = INT_MAX;
    s_uhi = -1;
  }
  const double q0 = __dsub_rn(b0, eps);
  int last_sure = -1, first_unsure = INT_MAX, last_unsure = -1;
  constexpr int kRegs = 8;  // candidates per thread kept in registers (n <= 2048): straight-line code, loads issued together
  if (per <= kRegs) {
    double v[kRegs];  // an invalid candidate is +inf: never accepted, never a new minimum
#pragma unroll
    for (int j = 0; j < kRegs; j++) {
      v[j] = kPosInf;
      double t;
      if (j < per && k0 + j < n && cand(k0 + j, t) && t == t) v[j] = t;
    }
    double lm = dmin(dmin(dmin(v[0], v[1]), dmin(v[2], v[3])), dmin(dmin(v[4], v[5]), dmin(v[6], v[7])));
    double inc = lm;  // inclusive prefix minimum over the threads of the warp
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const double t = __shfl_up_sync(kFull, inc, o);
      if (lane >= o) inc = dmin(inc, t);
    }
    double m = __shfl_up_sync(kFull, inc, 1);  // minimum of the valid values before k0 (inside the warp so far)
    if (lane == 0) m = kPosInf;
    if (lane == 31) s_tot[warp] = inc;
    scan_threads_barrier();
#pragma unroll
    for (int w = 0; w < kScanThreads / 32 - 1; w++) {
      const double t = s_tot[w];
      if (w < warp) m = dmin(m, t);
    }
#pragma unroll
    for (int j = 0; j < kRegs; j++) {  // only m = dmin(m, v) is a dependent chain; the thresholds hang off it
      const double mk = m;
      m = dmin(m, v[j]);
      const bool sure = v[j] < __dsub_rn(dmin(b0, mk), eps);
      const bool unsure = !sure && v[j] < dmin(q0, mk);
      if (sure) last_sure = k0 + j;
      if (unsure) {
        first_unsure = min(first_unsure, k0 + j);
        last_unsure = k0 + j;
      }
    }
  } else {
    double lm = kPosInf;
    for (int k = k0; k < k1; k++) {
      double v;
      if (cand(k, v) && v == v) lm = dmin(lm, v);
    }
    double inc = lm;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const double t = __shfl_up_sync(kFull, inc, o);
      if (lane >= o) inc = dmin(inc, t);
    }
    double m = __shfl_up_sync(kFull, inc, 1);
    if (lane == 0) m = kPosInf;
    if (lane == 31) s_tot[warp] = inc;
    scan_threads_barrier();
    for (int w = 0; w < warp; w++) m = dmin(m, s_tot[w]);
    double sure_below = __dsub_rn(dmin(b0, m), eps), unsure_below = dmin(q0, m);
    for (int k = k0; k < k1; k++) {
      double v;
      if (!(cand(k, v) && v == v)) continue;
      if (v < sure_below) {
        last_sure = k;
      } else if (v < unsure_below) {
        if (first_unsure == INT_MAX) first_unsure = k;
        last_unsure = k;
      }
      if (v < m) {
        m = v;
        sure_below = __dsub_rn(dmin(b0, m), eps);
        unsure_below = dmin(q0, m);
      }
    }
  }
  if (last_sure >= 0) atomicMax(&s_kstar, last_sure);
  if (last_unsure >= 0) {
    atomicMin(&s_ulo, first_unsure);
    atomicMax(&s_uhi, last_unsure);
  }
  scan_threads_barrier();
  int idx = s_kstar;
  double best = b0;
  if (idx >= 0) cand(idx, best);
  const int lo = max(s_ulo, idx + 1), hi = s_uhi;
  scan_threads_barrier();  // the exchange words are reused by the next call
  if (lo <= hi) {
    // literal replay of [lo, hi], 32 candidates per ballot (every scanning warp, redundantly: no further exchange)
    for (int base = lo; base <= hi; base += 32) {
      const int k = base + lane;
      double val = 0.0;
      const bool ok = k <= hi && cand(k, val) && val == val;
      int from = 0;
      while (true) {
        const bool acc = ok && lane >= from && val < __dsub_rn(best, eps);
        const unsigned mask = __ballot_sync(kFull, acc);
        if (!mask) break;
        const int l = __ffs(mask) - 1;
        best = __shfl_sync(kFull, val, l);
        idx = base + l;
        from = l + 1;
      }
    }
  }
  return idx;
}
// block-wide form: all threads of the block must call it; blockDim.x >= kScanThreads
template <class Cand>
__device__ int block_hyst_min_scan(int n, Cand cand, double b0, double eps) {
  __shared__ int s_res;
  if (threadIdx.x < kScanThreads) {
    const int r = hyst_min_scan_threads(n, cand, b0, eps);
    if (threadIdx.x == 0) s_res = r;
  }
  __syncthreads();
  const int r = s_res;
  __syncthreads();  // s_res is reused by the next call
  return r;
}

// first index of the minimum over valid candidates
template <class Cand>
__device__ int block_first_min(int n, Cand cand, MinIdx* sm, double* vout = nullptr) {
  MinIdx m = minidx_identity();
  for (int k = threadIdx.x; k < n; k += blockDim.x) {
    double val;
    if (cand(k, val)) m = minidx_combine(m, MinIdx{val, k});
  }
  m = block_minidx(m, sm);
  if (vout) *vout = m.v;
  return m.i == INT_MAX ? -1 : m.i;
}


}  // namespace lpr
