// common.cuh -- shared host/device helpers of liblprb200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cfloat>
#include <climits>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>

#include "../../include/lprb200.h"

namespace lpr {

// ---- error plumbing (thread local message, no exceptions across the ABI) -----------------
std::string& last_error();
int fail(int code, const char* fmt, ...);
extern std::atomic<int64_t> g_launches;
inline void count_launch(int n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

#define LPR_CUDA(expr)                                                                     \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess)                                                                 \
      return ::lpr::fail(_e == cudaErrorMemoryAllocation ? LPR_E_NOMEM : LPR_E_CUDA,       \
                         "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                         __LINE__);                                                        \
  } while (0)

#define LPR_LAUNCH_CHECK()                                                  \
  do {                                                                      \
    ::lpr::count_launch();                                                  \
    cudaError_t _e = cudaGetLastError();                                    \
    if (_e != cudaSuccess)                                                  \
      return ::lpr::fail(LPR_E_CUDA, "kernel launch failed: %s (%s:%d)",    \
                         cudaGetErrorString(_e), __FILE__, __LINE__);       \
  } while (0)

int select_device(int device);  // cudaSetDevice + sanity (fails loudly without a GPU)
int sm_count(int device);

// ---- synthetic generator (bit identical to oracle/lpr_oracle.cpp orc_u01) -----------------
__host__ __device__ inline uint64_t splitmix64(uint64_t x) {
  uint64_t z = x + 0x9E3779B97F4A7C15ULL;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}
__host__ __device__ inline double u01(uint64_t seed, uint64_t stream, uint64_t idx) {
  return (double)(splitmix64(seed + (stream << 40) + idx) >> 11) * 0x1.0p-53;
}

// ---- (value, index) reductions with deterministic lowest-index tie break -------------------
struct MinIdx {
  double v;
  int i;  // INT_MAX = no candidate
};
__device__ __forceinline__ MinIdx minidx_identity() { return MinIdx{__longlong_as_double(0x7ff0000000000000LL), INT_MAX}; }
__device__ __forceinline__ MinIdx minidx_combine(MinIdx a, MinIdx b) {
  // invalid entries (i == INT_MAX) lose; otherwise smaller value wins, ties -> lower index.
  if (b.i == INT_MAX) return a;
  if (a.i == INT_MAX) return b;
  if (b.v < a.v || (b.v == a.v && b.i < a.i)) return b;
  return a;
}
__device__ __forceinline__ MinIdx warp_minidx(MinIdx x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    MinIdx y;
    y.v = __shfl_xor_sync(0xffffffffu, x.v, o);
    y.i = __shfl_xor_sync(0xffffffffu, x.i, o);
    x = minidx_combine(x, y);
  }
  return x;
}
// block-wide; every thread gets the result.  smem must hold 32 MinIdx.
__device__ __forceinline__ MinIdx block_minidx(MinIdx x, MinIdx* smem) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  x = warp_minidx(x);
  __syncthreads();  // protect smem reuse across consecutive calls
  if (lane == 0) smem[w] = x;
  __syncthreads();
  MinIdx r = (lane < nw) ? smem[lane] : minidx_identity();
  r = warp_minidx(r);
  return r;
}
__device__ __forceinline__ int block_sum_int(int x, int* smem) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
  __syncthreads();
  if (lane == 0) smem[w] = x;
  __syncthreads();
  int r = (lane < nw) ? smem[lane] : 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
  return r;
}

// IEEE x / y, round to nearest.  div.rn.f64 expands to a reciprocal iteration guarded by a test on the DIVIDEND's
// exponent (SASS: FSETP.GEU |x.hi|, 6.58e-37) and CALLs a slow-path subroutine when it is tiny -- which includes
// x == 0, the most common entry of a pivot row (slack columns) and of a degenerate RHS.  One zero lane sends the whole
// warp through the subroutine (k_persist: 5 700 of 21 600 clocks per pivot were spent there).  0 / y is a zero with
// the sign of x XOR y for every non-zero, non-NaN y, so that case is answered directly; same bits otherwise.
__device__ __forceinline__ double ddiv(double x, double y) {
  if (x == 0.0 && fabs(y) > 0.0)
    return __longlong_as_double((__double_as_longlong(x) ^ __double_as_longlong(y)) & (long long)0x8000000000000000ULL);
  return __ddiv_rn(x, y);
}

// Math.Round(x, 4) of .NET Framework (BranchBoundSimplexSolver.cs:540-550): x*1e4, banker's
// round, /1e4 when |x| < 1e16.  rint() in the default rounding mode is value identical to
// COMDouble::Round for every finite double (tests/test_oracle_golden.py checks the oracle's
// literal floor(x+0.5) form against rint on the edge cases).
__device__ __forceinline__ double net_round4_div(double x) {  // the literal form
  if (fabs(x) < 1e16) {
    x = __dmul_rn(x, 1e4);
    x = rint(x);
    x = __ddiv_rn(x, 1e4);
  }
  return x;
}
// Same value without the IEEE division: for an integer-valued k with |k| <= 2^31 the quotient k / 1e4 is
// q1 = fma(r, RN(1e-4), q0) with q0 = k * RN(1e-4), r = fma(-1e4, q0, k) (Markstein's correction step;
// tools/div1e4_check.c compares it with k / 1e4 for every such k, both signs: 0 mismatches).  |x| < 2e5 keeps
// |k| below 2^31; anything larger (and NaN) takes the literal form.  k == 0 returns k so -0.0 stays -0.0.
__device__ __forceinline__ double net_round4(double x) {
  if (fabs(x) < 2e5) {
    const double k = rint(__dmul_rn(x, 1e4));
    if (k == 0.0) return k;
    const double q0 = __dmul_rn(k, 1e-4);
    const double r = __fma_rn(-1e4, q0, k);
    return __fma_rn(r, 1e-4, q0);
  }
  return net_round4_div(x);
}
// Round(Round(x, 4), 4): on |x| < 2e5 the second rounding is the identity (rint((k / 1e4) * 1e4) == k for every
// |k| <= 2^31, same exhaustive check), elsewhere it is applied literally.
__device__ __forceinline__ double net_round4_twice(double x) {
  return fabs(x) < 2e5 ? net_round4(x) : net_round4_div(net_round4_div(x));
}
// |Round(x, 4) - 1| <= 1e-6  <=>  rint(x * 1e4) == 1e4 (neighbouring quotients are 1e-4 away; |x| >= 1e16 and NaN
// are never near 1): the "is a rounded 1" test of the basic-variable scans without the quotient.
__device__ __forceinline__ bool net_round4_is_one(double x) { return rint(__dmul_rn(x, 1e4)) == 1e4; }
// CuttingPlaneSolver.cs:12-17
__device__ __forceinline__ double net_frac(double a) {
  double f = __dsub_rn(a, floor(a));
  if (fabs(f) < 1e-9 || fabs(__dsub_rn(1.0, f)) < 1e-9) return 0.0;
  return f;
}

inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

}  // namespace lpr
