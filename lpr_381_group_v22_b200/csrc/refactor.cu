// refactor.cu -- periodic refactorisation of B^-1 for the revised simplex (north star: "the only step
// allowed tensor cores (FP64 DMMA)"; the reference never refactorises).
//
// B^-1 is maintained by product-form updates (revised.cu k_update) whose rounding errors accumulate.
// The refresh recomputes it from the basis columns with one Newton-Schulz step
//        R  = I - B X            X = current B^-1
//        X' = X + X R            (= X (2I - B X): quadratic error contraction, no pivoting needed
//                                 because X is already accurate to ~1e-10)
// i.e. two m x m x m FP64 GEMMs on the tensor cores (mma.sync.m8n8k4.f64 -- tcgen05 has no FP64 kind),
// 4 m^3 flop = 2.2 TFLOP at m = 8192.  B is gathered from [A | I] by the basis list.  Results move by
// O(accumulated drift) only, far below the 1e-9 tolerance of this path.
#include <algorithm>
#include <vector>

#include "common.cuh"

namespace lpr {

constexpr int GM = 64, GN = 64, GK = 16;  // CTA tile
constexpr int BS_STRIDE = GN + 4;         // conflict-free fragment reads (see below)

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N)); }

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

// C = alpha * A * B + beta * D (+ I when add_identity).  All matrices n x n row-major with leading
// dimension ld, n a multiple of 64.  128 threads = 2x2 warps, each warp a 32x32 tile = 4x4 DMMA tiles.
// smem layouts chosen so that the m8n8k4 fragment loads are bank-conflict free:
//   As[kb][row][4]   (A fragment: lane -> row = lane>>2, k = lane&3  => address 4*row + k)
//   Bs[k][BS_STRIDE] (B fragment: lane -> k = lane&3, col = lane>>2  => address 68*k + col == 4k+col mod 16)
__global__ void __launch_bounds__(128) k_dgemm(int n, int ld, double alpha, const double* __restrict__ A,
                                               const double* __restrict__ B, double beta,
                                               const double* __restrict__ D, int add_identity,
                                               double* __restrict__ C) {
  __shared__ __align__(16) double As[2][GK / 4][GM][4];
  __shared__ __align__(16) double Bs[2][GK][BS_STRIDE];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = (warp >> 1) * 32, wn = (warp & 1) * 32;
  const int m0 = blockIdx.y * GM, n0 = blockIdx.x * GN;
  double acc[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

  auto load_tile = [&](int buf, int k0) {
#pragma unroll
    for (int q = 0; q < 4; q++) {  // A tile: 64 rows x 16 k = 512 16-byte chunks
      const int c = tid + q * 128;
      const int r = c >> 3, kk = (c & 7) * 2;
      cp_async16(&As[buf][kk >> 2][r][kk & 3], A + (size_t)(m0 + r) * ld + k0 + kk);
    }
#pragma unroll
    for (int q = 0; q < 4; q++) {  // B tile: 16 k x 64 cols = 512 chunks
      const int c = tid + q * 128;
      const int k = c >> 5, nn = (c & 31) * 2;
      cp_async16(&Bs[buf][k][nn], B + (size_t)(k0 + k) * ld + n0 + nn);
    }
    cp_async_commit();
  };

  const int nk = n / GK;
  load_tile(0, 0);
  for (int kt = 0; kt < nk; kt++) {
    const int buf = kt & 1;
    if (kt + 1 < nk) {
      load_tile(buf ^ 1, (kt + 1) * GK);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
#pragma unroll
    for (int kb = 0; kb < GK / 4; kb++) {
      double a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; i++) a[i] = As[buf][kb][wm + i * 8 + (lane >> 2)][lane & 3];
#pragma unroll
      for (int j = 0; j < 4; j++) b[j] = Bs[buf][kb * 4 + (lane & 3)][wn + j * 8 + (lane >> 2)];
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const int row = m0 + wm + i * 8 + (lane >> 2);
      const int col = n0 + wn + j * 8 + (lane & 3) * 2;
      double2 out;
      out.x = alpha * acc[i][j][0];
      out.y = alpha * acc[i][j][1];
      if (beta != 0.0) {
        const double2 d = *reinterpret_cast<const double2*>(D + (size_t)row * ld + col);
        out.x += beta * d.x;
        out.y += beta * d.y;
      }
      if (add_identity) {
        if (row == col) out.x += 1.0;
        if (row == col + 1) out.y += 1.0;
      }
      *reinterpret_cast<double2*>(C + (size_t)row * ld + col) = out;
    }
}

// Bm = basis columns of [A | I], padded to np x np with an identity tail; Xp = padded copy of B^-1
__global__ void k_gather_basis(double* Bm, double* Xp, int np, int m, int n, const double* A, int ldA,
                               const double* Binv, int ldB, const int* basis) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = blockIdx.y;
  if (j >= np) return;
  double bv, xv;
  if (i < m && j < m) {
    const int var = basis[j];
    bv = (var < n) ? A[(size_t)i * ldA + var] : ((i == var - n) ? 1.0 : 0.0);
    xv = Binv[(size_t)i * ldB + j];
  } else {
    bv = xv = (i == j) ? 1.0 : 0.0;
  }
  Bm[(size_t)i * np + j] = bv;
  Xp[(size_t)i * np + j] = xv;
}
__global__ void k_scatter_binv(double* Binv, int ldB, int m, const double* Xn, int np) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = blockIdx.y;
  if (j < m) Binv[(size_t)i * ldB + j] = Xn[(size_t)i * np + j];
}
__global__ void k_absmax(const double* R, size_t count, double* out) {
  double mx = 0.0;
  for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < count; k += (size_t)gridDim.x * blockDim.x)
    mx = fmax(mx, fabs(R[k]));
  for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<unsigned long long*>(out), __double_as_longlong(mx));
}

// host entry used by revised.cu
int refactor_binv(cudaStream_t stream, int m, int n, const double* A, int ldA, double* Binv, int ldB,
                  const int* basis, double* residual_out, double* flops_out) {
  const int np = round_up(m, 64);
  const size_t bytes = sizeof(double) * (size_t)np * np;
  double *Bm = nullptr, *Xp = nullptr, *Rm = nullptr, *d_res = nullptr;
  cudaError_t e = cudaMalloc(&Bm, bytes);
  if (e == cudaSuccess) e = cudaMalloc(&Xp, bytes);
  if (e == cudaSuccess) e = cudaMalloc(&Rm, bytes);
  if (e == cudaSuccess) e = cudaMalloc(&d_res, sizeof(double));
  if (e != cudaSuccess) {
    cudaFree(Bm); cudaFree(Xp); cudaFree(Rm); cudaFree(d_res);
    return fail(LPR_E_NOMEM, "refactorisation workspace (3 x %zu MB) allocation failed", bytes >> 20);
  }
  dim3 gg((np + 255) / 256, np);
  k_gather_basis<<<gg, 256, 0, stream>>>(Bm, Xp, np, m, n, A, ldA, Binv, ldB, basis);
  count_launch();
  dim3 grid(np / GN, np / GM);
  // R = I - B X
  k_dgemm<<<grid, 128, 0, stream>>>(np, np, -1.0, Bm, Xp, 0.0, nullptr, 1, Rm);
  count_launch();
  cudaMemsetAsync(d_res, 0, sizeof(double), stream);
  k_absmax<<<296, 256, 0, stream>>>(Rm, (size_t)np * np, d_res);
  count_launch();
  // X' = X + X R   (written over Bm)
  k_dgemm<<<grid, 128, 0, stream>>>(np, np, 1.0, Xp, Rm, 1.0, Xp, 0, Bm);
  count_launch();
  dim3 gs((m + 255) / 256, m);
  k_scatter_binv<<<gs, 256, 0, stream>>>(Binv, ldB, m, Bm, np);
  count_launch();
  double res = 0.0;
  e = cudaMemcpyAsync(&res, d_res, sizeof(double), cudaMemcpyDeviceToHost, stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
  if (e == cudaSuccess) e = cudaGetLastError();
  cudaFree(Bm); cudaFree(Xp); cudaFree(Rm); cudaFree(d_res);
  if (e != cudaSuccess) return fail(LPR_E_CUDA, "refactorisation failed: %s", cudaGetErrorString(e));
  if (residual_out) *residual_out = res;
  if (flops_out) *flops_out = 4.0 * (double)np * np * np;
  return LPR_OK;
}

}  // namespace lpr
