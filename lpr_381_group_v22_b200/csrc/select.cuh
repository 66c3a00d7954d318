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
  if (threadIdx.x == 0) {
    double best = b0;
    int idx = -1;
    for (int k = 0; k < n; k++) {
      double val;
      if (cand(k, val) && val < __dsub_rn(best, eps)) {
        best = val;
        idx = k;
      }
    }
    sh_res = idx;
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
