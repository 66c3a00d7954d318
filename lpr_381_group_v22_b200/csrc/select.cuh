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
