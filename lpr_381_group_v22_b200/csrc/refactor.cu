// refactor.cu -- periodic refactorisation of B^-1 for the revised simplex (north star: "the only step
// allowed tensor cores (FP64 DMMA)"; the reference never refactorises, SURVEY Q7, but keeps B for exactly this,
// RevisedPrimalSimplexSolver.cs:200-212).
//
// B^-1 is maintained by product-form updates (revised.cu k_update) whose rounding errors accumulate.  Two paths:
//
//  (1) cheap refresh, used while the current inverse X is still a good one (||I - B X||_inf < 1e-5): one Newton-Schulz
//      step   R = I - B X,   X' = X + X R   (quadratic error contraction) = two m^3 FP64 GEMMs;
//  (2) full refactorisation FROM THE BASIS COLUMNS ALONE, used when X is unusable (the guard above fails, or the
//      caller asks): blocked Gauss-Jordan inversion with partial pivoting.  With M = B and X = I, for every 64-column
//      panel J:  pivot rows by LU with partial pivoting on a scratch copy of the panel (one launch per column, the last
//      CTA to finish searches the next pivot);  the row swaps are applied to M and X;  with D = M[K,J] (the 64 x 64
//      pivot block) and F = M[O,J] (all other rows) the panel's Gauss-Jordan transform is
//            V[K,:] <- D^-1 V[K,:]         V[O,:] <- V[O,:] - F (D^-1 V[K,:])          for V = [M | X]
//      i.e. a small inverse, a 64-row product and ONE rank-64 GEMM update of the whole of V per panel -- the GEMM
//      carries 3 m^3 of the flops and runs on the FP64 tensor cores.  A final residual check polishes the result with
//      one Newton-Schulz step when needed.
// Both GEMM kernels are mma.sync.m8n8k4.f64 (tcgen05 has no FP64 kind).  B is gathered from [A | I] by the basis list.
// The workspace (3 padded m x m matrices + panel scratch) lives in the handle: nothing is allocated per call.
#include <algorithm>
#include <vector>

#include "common.cuh"

#include "refactor.cuh"

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
// infinity norm of the n x n residual: max over rows of sum_j |R[i][j]| (one warp per row).  ||I - B X||_inf < 1 is a
// sufficient condition for the Newton-Schulz refresh to contract (||E'|| <= ||E||^2); the largest entry alone is not.
__global__ void k_inf_norm(const double* R, int n, int ld, double* out) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
  double mx = 0.0;
  for (int i = warp; i < n; i += nwarps) {
    double s = 0.0;
    for (int j = lane; j < n; j += 32) s += fabs(R[(size_t)i * ld + j]);
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (!(s <= mx)) mx = s;  // NaN propagates
  }
  if (lane == 0) {
    if (mx != mx) mx = __longlong_as_double(0x7ff0000000000000LL);  // NaN -> +inf: never accepted by the guard
    atomicMax(reinterpret_cast<unsigned long long*>(out), __double_as_longlong(mx));
  }
}

// ---------------------------------------------------------------------------------------------------------------
// full refactorisation: blocked Gauss-Jordan inversion with partial pivoting
// ---------------------------------------------------------------------------------------------------------------
constexpr int NB = 64;  // panel width = GEMM tile height: the pivot block K is exactly one tile row
constexpr double kRefreshGuard = 1e-5;  // ||I - B X||_inf below which the Newton-Schulz refresh is used

__global__ void k_set_identity(double* X, int np) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = blockIdx.y;
  if (j < np) X[(size_t)i * np + j] = (i == j) ? 1.0 : 0.0;
}
// W[i][0..NB) = M[i][kb..kb+NB) for rows i >= kb (scratch copy the pivot search eliminates in)
__global__ void k_panel_copy(const double* __restrict__ M, int np, int kb, double* __restrict__ W) {
  const int i = kb + blockIdx.x;
  if (threadIdx.x < NB) W[(size_t)i * NB + threadIdx.x] = M[(size_t)i * np + kb + threadIdx.x];
}
// argmax |W[i][j]| over rows i >= kb that are not pivot rows yet (lowest row on ties); all threads of ONE CTA
__device__ void panel_pivot_search(const double* W, int np, int kb, int j, const int* used, int* piv_rows, int* singular) {
  __shared__ MinIdx sm[32];
  MinIdx m = minidx_identity();
  for (int i = kb + threadIdx.x; i < np; i += blockDim.x) {
    if (used[i]) continue;
    const double v = -fabs(__ldcg(W + (size_t)i * NB + j));  // min of -|v| = max of |v|
    m = minidx_combine(m, MinIdx{v, i});
  }
  m = block_minidx(m, sm);
  if (threadIdx.x == 0) {
    piv_rows[kb + j] = m.i == INT_MAX ? kb + j : m.i;
    if (m.i == INT_MAX || !(m.v < 0.0)) *singular = 1;  // the whole column is zero (or NaN): B is singular
  }
}
__global__ void __launch_bounds__(256) k_panel_first(const double* W, int np, int kb, int* used, int* piv_rows,
                                                     int* singular, unsigned* ticket) {
  for (int i = kb + threadIdx.x; i < np; i += blockDim.x) used[i] = 0;
  if (threadIdx.x == 0) *ticket = 0;
  __syncthreads();
  panel_pivot_search(W, np, kb, 0, used, piv_rows, singular);
}
// LU step for column j of the panel (pivot row r = piv_rows[kb + j]): every other candidate row i gets
// W[i][c] -= (W[i][j] / W[r][j]) * W[r][c] for c > j.  One warp per row.  The last CTA to finish marks r used and
// searches the pivot of column j + 1 in the updated scratch panel.
__global__ void __launch_bounds__(256) k_panel_step(double* W, int np, int kb, int j, int* used, int* piv_rows,
                                                    int* singular, unsigned* ticket) {
  __shared__ int s_last;
  const int r = piv_rows[kb + j];
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
  const double piv = W[(size_t)r * NB + j];
  const double p0 = W[(size_t)r * NB + lane], p1 = W[(size_t)r * NB + 32 + lane];
  for (int i = kb + warp; i < np; i += nwarps) {
    if (i == r || used[i]) continue;
    double* row = W + (size_t)i * NB;
    const double l = row[j] / piv;
    if (l != 0.0) {
      if (lane > j) row[lane] -= l * p0;
      if (32 + lane > j) row[32 + lane] -= l * p1;
    }
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  if (threadIdx.x == 0) {
    used[r] = 1;
    *ticket = 0;
  }
  __syncthreads();
  if (j + 1 < NB) panel_pivot_search(W, np, kb, j + 1, used, piv_rows, singular);
}
// pivot rows (indices into the row order at the start of the panel) -> sequential swap list: after "swap rows kb + t
// and sw[kb + t]" for t = 0 .. j, position kb + j holds the chosen pivot row of column j
__global__ void k_build_swaps(const int* piv_rows, int kb, int* sw) {
  if (threadIdx.x || blockIdx.x) return;
  for (int j = 0; j < NB; j++) {
    int pos = piv_rows[kb + j];
    for (int t = 0; t < j; t++) {  // where the earlier swaps have moved that row
      const int a = kb + t, b = sw[a];
      if (pos == a) pos = b;
      else if (pos == b) pos = a;
    }
    sw[kb + j] = pos;
  }
}
// apply the panel's swaps to columns [col0, col0 + ncols) of Mat; one thread per column, swaps in order
__global__ void k_swap_rows(double* Mat, int ld, int col0, int ncols, int kb, const int* __restrict__ sw) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncols) return;
  double* col = Mat + col0 + c;
  for (int j = 0; j < NB; j++) {
    const int a = kb + j, b = sw[a];
    if (a != b) {
      const double t = col[(size_t)a * ld];
      col[(size_t)a * ld] = col[(size_t)b * ld];
      col[(size_t)b * ld] = t;
    }
  }
}
// Dinv = inverse of the pivot block D = M[kb.., kb..] (NB x NB), Gauss-Jordan with partial pivoting in shared memory
__global__ void __launch_bounds__(1024) k_block_inverse(const double* __restrict__ M, int np, int kb, double* __restrict__ Dinv,
                                                        int* singular) {
  extern __shared__ double a_dyn[];  // NB x (2 NB + 1) doubles = 66 KB: opt-in dynamic shared memory
  double(*a)[2 * NB + 1] = reinterpret_cast<double(*)[2 * NB + 1]>(a_dyn);
  __shared__ int s_p;
  const int tid = threadIdx.x;
  for (int q = tid; q < NB * NB; q += blockDim.x) {
    const int i = q / NB, j = q % NB;
    a[i][j] = M[(size_t)(kb + i) * np + kb + j];
    a[i][NB + j] = (i == j) ? 1.0 : 0.0;
  }
  __syncthreads();
  for (int k = 0; k < NB; k++) {
    if (tid < 32) {  // pivot: largest |a[i][k]|, i >= k
      double best = -1.0;
      int bi = k;
      for (int i = k + tid; i < NB; i += 32) {
        const double v = fabs(a[i][k]);
        if (v > best) { best = v; bi = i; }
      }
      for (int o = 16; o > 0; o >>= 1) {
        const double ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
      }
      if (tid == 0) {
        s_p = bi;
        if (!(best > 0.0)) *singular = 1;
      }
    }
    __syncthreads();
    const int p = s_p;
    if (p != k)
      for (int j = tid; j < 2 * NB; j += blockDim.x) {
        const double t = a[k][j];
        a[k][j] = a[p][j];
        a[p][j] = t;
      }
    __syncthreads();
    const double piv = a[k][k];
    __syncthreads();
    for (int j = tid; j < 2 * NB; j += blockDim.x) a[k][j] = a[k][j] / piv;
    __syncthreads();
    for (int q = tid; q < NB * 2 * NB; q += blockDim.x) {
      const int i = q / (2 * NB), j = q % (2 * NB);
      if (i != k && j != k) a[i][j] -= a[i][k] * a[k][j];
    }
    __syncthreads();
    for (int i = tid; i < NB; i += blockDim.x)
      if (i != k) a[i][k] = 0.0;
    __syncthreads();
  }
  for (int q = tid; q < NB * NB; q += blockDim.x) Dinv[q] = a[q / NB][NB + q % NB];
}
// RK[q][c] = sum_t Dinv[q][t] * V[kb + t][col0 + c]   (NB x ncols; one thread per column, Dinv in shared memory)
__global__ void __launch_bounds__(128) k_pivot_rows(const double* __restrict__ Dinv, const double* __restrict__ V, int ld,
                                                    int col0, int ncols, int kb, double* __restrict__ RK, int ldr) {
  __shared__ double d[NB * NB];
  for (int q = threadIdx.x; q < NB * NB; q += blockDim.x) d[q] = Dinv[q];
  __syncthreads();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncols) return;
  double v[NB];
#pragma unroll
  for (int t = 0; t < NB; t++) v[t] = V[(size_t)(kb + t) * ld + col0 + c];
  for (int q = 0; q < NB; q++) {
    double s = 0.0;
#pragma unroll
    for (int t = 0; t < NB; t++) s += d[q * NB + t] * v[t];
    RK[(size_t)q * ldr + c] = s;
  }
}
__global__ void k_store_pivot_rows(double* V, int ld, int col0, int ncols, int kb, const double* __restrict__ RK, int ldr) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const int q = blockIdx.y;
  if (c < ncols) V[(size_t)(kb + q) * ld + col0 + c] = RK[(size_t)q * ldr + c];
}
// V[i][col0 + j] -= sum_q F[i][q] * RK[q][j] for every row block except the pivot block (blockIdx.y == kb / 64), where
// F = M[:, kb .. kb + 64) (leading dimension ldf).  64 x 64 tile per CTA, K = 64, FP64 DMMA; same fragment layouts as
// k_dgemm.  ncols a multiple of 64.
__global__ void __launch_bounds__(128) k_rank64_update(double* __restrict__ V, int ld, int col0, const double* __restrict__ F,
                                                       int ldf, int kb, const double* __restrict__ RK, int ldr) {
  if ((int)blockIdx.y * GM == kb) return;
  constexpr int KH = NB / 2;  // K is consumed in two halves: 16 KB + 17 KB of static shared memory
  __shared__ __align__(16) double As[KH / 4][GM][4];
  __shared__ __align__(16) double Bs[KH][BS_STRIDE];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = (warp >> 1) * 32, wn = (warp & 1) * 32;
  const int m0 = blockIdx.y * GM, n0 = blockIdx.x * GN;
  double acc[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) acc[i][j][0] = acc[i][j][1] = 0.0;
  for (int half = 0; half < 2; half++) {
    const int k0 = half * KH;
#pragma unroll
    for (int q = 0; q < 8; q++) {  // A tile: 64 rows x 32 k = 1024 16-byte chunks
      const int c = tid + q * 128;
      const int r = c >> 4, kk = (c & 15) * 2;
      cp_async16(&As[kk >> 2][r][kk & 3], F + (size_t)(m0 + r) * ldf + kb + k0 + kk);
    }
#pragma unroll
    for (int q = 0; q < 8; q++) {  // B tile: 32 k x 64 cols = 1024 chunks
      const int c = tid + q * 128;
      const int k = c >> 5, nn = (c & 31) * 2;
      cp_async16(&Bs[k][nn], RK + (size_t)(k0 + k) * ldr + n0 + nn);
    }
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
#pragma unroll
    for (int kq = 0; kq < KH / 4; kq++) {
      double a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; i++) a[i] = As[kq][wm + i * 8 + (lane >> 2)][lane & 3];
#pragma unroll
      for (int j = 0; j < 4; j++) b[j] = Bs[kq * 4 + (lane & 3)][wn + j * 8 + (lane >> 2)];
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
      const int col = col0 + n0 + wn + j * 8 + (lane & 3) * 2;
      double2* p = reinterpret_cast<double2*>(V + (size_t)row * ld + col);
      double2 x = *p;
      x.x -= acc[i][j][0];
      x.y -= acc[i][j][1];
      *p = x;
    }
}

// X <- inverse of the np x np matrix in Mw (destroyed), blocked Gauss-Jordan with partial pivoting
constexpr size_t kBlockInvSmem = sizeof(double) * NB * (2 * NB + 1);
static int invert_blocked_gj(cudaStream_t stream, int np, double* Mw, double* X, RefactorWs& ws, double* flops) {
  // per device, so set on every call (cheap)
  LPR_CUDA(cudaFuncSetAttribute(k_block_inverse, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBlockInvSmem));
  k_set_identity<<<dim3((np + 255) / 256, np), 256, 0, stream>>>(X, np);
  count_launch();
  LPR_CUDA(cudaMemsetAsync(ws.singular, 0, sizeof(int), stream));
  int sms = 148;
  {
    int dev = 0;
    cudaGetDevice(&dev);
    sms = sm_count(dev);
  }
  double fl = 0.0;
  for (int kb = 0; kb < np; kb += NB) {
    const int rows = np - kb;
    k_panel_copy<<<rows, NB, 0, stream>>>(Mw, np, kb, ws.W);
    k_panel_first<<<1, 256, 0, stream>>>(ws.W, np, kb, ws.used, ws.piv_rows, ws.singular, ws.ticket);
    const int g = std::max(1, std::min(sms * 2, (rows + 7) / 8));
    for (int j = 0; j < NB; j++)
      k_panel_step<<<g, 256, 0, stream>>>(ws.W, np, kb, j, ws.used, ws.piv_rows, ws.singular, ws.ticket);
    k_build_swaps<<<1, 1, 0, stream>>>(ws.piv_rows, kb, ws.sw);
    const int mcols = np - kb;  // columns of M that still matter (earlier ones are unit vectors with zeros below kb)
    k_swap_rows<<<(mcols + 255) / 256, 256, 0, stream>>>(Mw, np, kb, mcols, kb, ws.sw);
    k_swap_rows<<<(np + 255) / 256, 256, 0, stream>>>(X, np, 0, np, kb, ws.sw);
    k_block_inverse<<<1, 1024, kBlockInvSmem, stream>>>(Mw, np, kb, ws.Dinv, ws.singular);
    count_launch(7 + NB);
    const int rest = np - kb - NB;  // columns of M to the right of the panel
    if (rest > 0) {
      k_pivot_rows<<<(rest + 127) / 128, 128, 0, stream>>>(ws.Dinv, Mw, np, kb + NB, rest, kb, ws.RK, 2 * np);
      k_rank64_update<<<dim3(rest / GN, np / GM), 128, 0, stream>>>(Mw, np, kb + NB, Mw, np, kb, ws.RK, 2 * np);
      k_store_pivot_rows<<<dim3((rest + 255) / 256, NB), 256, 0, stream>>>(Mw, np, kb + NB, rest, kb, ws.RK, 2 * np);
      count_launch(3);
    }
    k_pivot_rows<<<(np + 127) / 128, 128, 0, stream>>>(ws.Dinv, X, np, 0, np, kb, ws.RK + np, 2 * np);
    k_rank64_update<<<dim3(np / GN, np / GM), 128, 0, stream>>>(X, np, 0, Mw, np, kb, ws.RK + np, 2 * np);
    k_store_pivot_rows<<<dim3((np + 255) / 256, NB), 256, 0, stream>>>(X, np, 0, np, kb, ws.RK + np, 2 * np);
    count_launch(3);
    fl += 2.0 * (double)(np - NB) * NB * ((double)rest + np) + 2.0 * NB * NB * ((double)rest + np);
  }
  int sing = 0;
  LPR_CUDA(cudaMemcpyAsync(&sing, ws.singular, sizeof(int), cudaMemcpyDeviceToHost, stream));
  LPR_CUDA(cudaStreamSynchronize(stream));
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(LPR_E_CUDA, "blocked inversion failed: %s", cudaGetErrorString(e));
  if (sing) return fail(LPR_E_STATE, "refactorisation: the basis matrix is singular");
  if (flops) *flops += fl;
  return LPR_OK;
}

// host entry used by revised.cu.  mode 0 = automatic (Newton-Schulz refresh when max |I - B X| < 0.5, else the full
// path), 1 = refresh only, 2 = full refactorisation from the basis columns.  *path_out: 1 refresh, 2 full.
int refactor_binv(cudaStream_t stream, int m, int n, const double* A, int ldA, double* Binv, int ldB,
                  const int* basis, RefactorWs& ws, int mode, double* residual_out, double* residual_after_out,
                  double* flops_out, int* path_out) {
  const int np = round_up(m, 64);
  const size_t bytes = sizeof(double) * (size_t)np * np;
  int rc0 = refactor_ws_ensure(ws, m);
  if (rc0) return rc0;
  double *Bm = ws.Bm, *Xp = ws.Xp, *Rm = ws.Rm;
  dim3 gg((np + 255) / 256, np), grid(np / GN, np / GM);
  double flops = 0.0, res0 = 0.0, res1 = 0.0;
  int path = 0;
  auto residual = [&](const double* X, double* out_host, int slot) -> int {  // Rm = I - B X, max |Rm|
    k_dgemm<<<grid, 128, 0, stream>>>(np, np, -1.0, Bm, X, 0.0, nullptr, 1, Rm);
    cudaMemsetAsync(ws.d_res + slot, 0, sizeof(double), stream);
    k_inf_norm<<<296, 256, 0, stream>>>(Rm, np, np, ws.d_res + slot);
    count_launch(2);
    flops += 2.0 * (double)np * np * np;
    LPR_CUDA(cudaMemcpyAsync(out_host, ws.d_res + slot, sizeof(double), cudaMemcpyDeviceToHost, stream));
    LPR_CUDA(cudaStreamSynchronize(stream));
    return LPR_OK;
  };
  k_gather_basis<<<gg, 256, 0, stream>>>(Bm, Xp, np, m, n, A, ldA, Binv, ldB, basis);
  count_launch();
  int rc = LPR_OK;
  const double* result = nullptr;
  if (mode != 2) {
    if ((rc = residual(Xp, &res0, 0))) return rc;
    // refresh only while it finishes the job in ONE step (error e -> e^2): beyond 1e-5 a chain of refreshes costs more
    // than the full path, beyond 1 it diverges
    if (mode == 1 || (res0 == res0 && res0 < kRefreshGuard)) {
      // X' = X + X R   (written over Bm: B is not needed any more on this path)
      k_dgemm<<<grid, 128, 0, stream>>>(np, np, 1.0, Xp, Rm, 1.0, Xp, 0, Bm);
      count_launch();
      flops += 2.0 * (double)np * np * np;
      result = Bm;
      path = 1;
      res1 = res0 * res0;  // bound (||E'|| <= ||E||^2), not measured: the refresh is the cheap path
    }
  }
  if (!result) {
    // full path: Rm <- copy of B (destroyed by the elimination), Xp <- B^-1
    LPR_CUDA(cudaMemcpyAsync(Rm, Bm, bytes, cudaMemcpyDeviceToDevice, stream));
    if ((rc = invert_blocked_gj(stream, np, Rm, Xp, ws, &flops))) return rc;
    path = 2;
    if ((rc = residual(Xp, &res1, 1))) return rc;
    if (res1 > 1e-12 && res1 < 0.5) {  // polish: one Newton-Schulz step, X' = X + X R (into Bm)
      k_dgemm<<<grid, 128, 0, stream>>>(np, np, 1.0, Xp, Rm, 1.0, Xp, 0, Bm);
      count_launch();
      flops += 2.0 * (double)np * np * np;
      result = Bm;
    } else {
      result = Xp;
    }
    if (mode == 2) res0 = res1;
  }
  dim3 gs((m + 255) / 256, m);
  k_scatter_binv<<<gs, 256, 0, stream>>>(Binv, ldB, m, result, np);
  count_launch();
  cudaError_t e = cudaStreamSynchronize(stream);
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) return fail(LPR_E_CUDA, "refactorisation failed: %s", cudaGetErrorString(e));
  if (residual_out) *residual_out = res0;
  if (residual_after_out) *residual_after_out = res1;
  if (flops_out) *flops_out = flops;
  if (path_out) *path_out = path;
  return LPR_OK;
}

// allocate the workspace for order m (idempotent); lpr_rev_refactor_ex calls it before it starts its timer
int refactor_ws_ensure(RefactorWs& ws, int m) {
  const int np = round_up(m, 64);
  const size_t bytes = sizeof(double) * (size_t)np * np;
  if (ws.np == np) return LPR_OK;
  refactor_ws_free(ws);
  cudaError_t e = cudaMalloc(&ws.Bm, bytes);
  if (e == cudaSuccess) e = cudaMalloc(&ws.Xp, bytes);
  if (e == cudaSuccess) e = cudaMalloc(&ws.Rm, bytes);
  if (e == cudaSuccess) e = cudaMalloc(&ws.W, sizeof(double) * (size_t)np * NB);
  if (e == cudaSuccess) e = cudaMalloc(&ws.RK, sizeof(double) * (size_t)NB * 2 * np);
  if (e == cudaSuccess) e = cudaMalloc(&ws.Dinv, sizeof(double) * NB * NB);
  if (e == cudaSuccess) e = cudaMalloc(&ws.d_res, sizeof(double) * 2);
  if (e == cudaSuccess) e = cudaMalloc(&ws.used, sizeof(int) * np);
  if (e == cudaSuccess) e = cudaMalloc(&ws.piv_rows, sizeof(int) * np);
  if (e == cudaSuccess) e = cudaMalloc(&ws.sw, sizeof(int) * np);
  if (e == cudaSuccess) e = cudaMalloc(&ws.singular, sizeof(int));
  if (e == cudaSuccess) e = cudaMalloc(&ws.ticket, sizeof(unsigned));
  if (e != cudaSuccess) {
    refactor_ws_free(ws);
    return fail(LPR_E_NOMEM, "refactorisation workspace (3 x %zu MB) allocation failed", bytes >> 20);
  }
  ws.np = np;
  return LPR_OK;
}

void refactor_ws_free(RefactorWs& ws) {
  cudaFree(ws.Bm); cudaFree(ws.Xp); cudaFree(ws.Rm); cudaFree(ws.W); cudaFree(ws.RK); cudaFree(ws.Dinv);
  cudaFree(ws.d_res); cudaFree(ws.used); cudaFree(ws.piv_rows); cudaFree(ws.sw); cudaFree(ws.singular);
  cudaFree(ws.ticket);
  ws = RefactorWs();
}

}  // namespace lpr
