// revised.cu -- RevisedPrimalSimplexSolver (Simplex/RevisedPrimalSimplexSolver.cs:82-287) with A and
// B^-1 resident in HBM.  Per iteration (reference schedule, de-duplicated -- DESIGN.md "revised"):
//   k_price   : partial column sums  sum_i y_i A[i,j]            reads A        8mn bytes
//   k_rc      : rc_j = c_j - sum(partials)                       (:96-102)
//   k_enter   : largest rc > 1e-9 with 1e-9 hysteresis, lowest index (:104-121); gathers a_e
//   k_dir     : u = B^-1 a_e and x_B = B^-1 b in ONE pass        reads B^-1     8m^2 bytes (:89,:149-151)
//   k_ratio   : infeasibility check (:90-91), Bland ratio test (:153-176), eta column (:264-272)
//   k_update  : B^-1 <- E B^-1 (:274, :426-441), fused y' = c_B' B^-1' partial sums (:93)
//                                                                 reads+writes B^-1 16m^2 bytes
//   k_y       : y = sum(partials)
// The B^-1 update is bit-identical to the reference's MultiplyMatrices(E, B^-1) given the same u;
// the dot products use fixed-shape parallel trees (deterministic) instead of the reference's
// sequential sums, which is why this path is held to 1e-9 relative, not bit equality.
#include <algorithm>
#include <cstdlib>
#include <string>
#include <vector>

#include "common.cuh"

namespace lpr {

struct RevState {
  int status;
  int enter;       // entering variable (0..n+m-1) or -1
  int leave_row;   // leaving row
  int leave_var;
  int do_update;   // 1 when k_update must apply the staged eta
  int need_final;  // optimal detected: x_B must be (re)computed and checked
  long long iter;
  long long max_iter;
  double pivot;
};

struct RevView {
  int m, n, ldA, ldB;
  const double* A;   // m x ldA
  double* Binv;      // m x ldB
  const double* b;   // m
  const double* c;   // n (negated when minimising, :51)
  double* cB;        // m
  double* xB;        // m
  double* y;         // m
  double* rc;        // n
  double* u;         // m
  double* ekey;      // n+m : entering-candidate keys (k_enter scratch)
  double* ab;        // 2*m interleaved (a_e[i], b[i])
  double* ecoef;     // m : eta column (1/piv at r, -u_i/piv elsewhere)
  double* brow;      // ldB : copy of row r of B^-1 before the update
  double* ppart;     // price partials  PS x ldA
  double* ypart;     // y partials      YS x ldB
  int PS, YS;
  int* basis;        // m
  int* isbasic;      // n+m
  RevState* st;
  int* log;          // (leaveRow, enter, leaveVar)
  long long log_cap;
  int dense;         // 1 = never skip zero multipliers (roofline measurements)
};

constexpr int kT = 256;
constexpr double kInfD = __builtin_huge_val();

// streaming 128-bit load that does not allocate in L1 (the small vectors re-read by every thread stay there)
__device__ __forceinline__ double2 ld_stream2(const double2* p) {
  double2 r;
  asm volatile("ld.global.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
  return r;
}
// PDL: wait for the producer kernel, then let the consumer start launching (no-ops without the attribute)
__device__ __forceinline__ void rev_pdl() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

// ---- generic (value,index) helpers are in common.cuh ------------------------------------------

// k_price: grid (colTiles, PS).  Thread owns one double2 column chunk and walks its row range.
__global__ void __launch_bounds__(kT) k_price(RevView v) {
  rev_pdl();
  const RevState* st = v.st;
  if (st->status != LPR_RUNNING) return;
  const int ldv = v.ldA >> 1;
  const int chunk = blockIdx.x * kT + threadIdx.x;
  const int rs = blockIdx.y;
  const int r0 = (int)((long long)v.m * rs / v.PS), r1 = (int)((long long)v.m * (rs + 1) / v.PS);
  if (chunk >= ldv) return;
  const double2* __restrict__ A2 = reinterpret_cast<const double2*>(v.A) + chunk;
  double2 acc = make_double2(0.0, 0.0);
  constexpr int U = 8;
  for (int r = r0; r < r1; r += U) {
    double yi[U];
    double2 a[U];
#pragma unroll
    for (int k = 0; k < U; k++) {
      const int row = r + k;
      yi[k] = 0.0;
      if (row < r1) {
        yi[k] = v.y[row];
        if (yi[k] != 0.0 || v.dense) a[k] = A2[(size_t)row * ldv];
      }
    }
#pragma unroll
    for (int k = 0; k < U; k++) {
      const int row = r + k;
      if (row < r1 && (yi[k] != 0.0 || v.dense)) {  // Dot :443-448: s += y_i * A_ij (mul, then add)
        acc.x = __dadd_rn(acc.x, __dmul_rn(yi[k], a[k].x));
        acc.y = __dadd_rn(acc.y, __dmul_rn(yi[k], a[k].y));
      }
    }
  }
  reinterpret_cast<double2*>(v.ppart)[(size_t)rs * ldv + chunk] = acc;
}

// fixed-order sum of the row-split partials with the loads batched 8 at a time
__device__ __forceinline__ double sum_partials(const double* part, int splits, size_t stride, int j) {
  double s = 0.0;
  for (int r0 = 0; r0 < splits; r0 += 8) {
    double t[8];
#pragma unroll
    for (int q = 0; q < 8; q++) t[q] = (r0 + q < splits) ? part[(size_t)(r0 + q) * stride + j] : 0.0;
#pragma unroll
    for (int q = 0; q < 8; q++)
      if (r0 + q < splits) s = __dadd_rn(s, t[q]);
  }
  return s;
}

__global__ void __launch_bounds__(kT) k_rc(RevView v) {
  rev_pdl();
  if (v.st->status != LPR_RUNNING) return;
  const int j = blockIdx.x * kT + threadIdx.x;
  if (j >= v.n) return;
  v.rc[j] = __dsub_rn(v.c[j], sum_partials(v.ppart, v.PS, v.ldA, j));  // :98
}

__global__ void __launch_bounds__(kT) k_y_force(RevView v) {
  const int j = blockIdx.x * kT + threadIdx.x;
  if (j >= v.m) return;
  v.y[j] = sum_partials(v.ypart, v.YS, v.ldB, j);
}

__global__ void __launch_bounds__(kT) k_y(RevView v) {
  rev_pdl();
  if (v.st->status != LPR_RUNNING) return;
  const int j = blockIdx.x * kT + threadIdx.x;
  if (j >= v.m) return;
  v.y[j] = sum_partials(v.ypart, v.YS, v.ldB, j);
}

// hysteresis scan (see tableau.cu block_hyst_min) specialised for "largest rc" (:104-121)
template <class Cand>
__device__ int rev_hyst_min(int n, Cand cand, double b0, double eps, MinIdx* sm, int* smi) {
  __shared__ int sh_res;
  MinIdx m = minidx_identity();
#pragma unroll 8
  for (int k = threadIdx.x; k < n; k += blockDim.x) {
    double val;
    if (cand(k, val) && val == val) m = minidx_combine(m, MinIdx{val, k});
  }
  m = block_minidx(m, sm);
  if (m.i == INT_MAX) return -1;
  if (!(m.v < __dsub_rn(b0, eps))) return -1;
  int bad = 0;
#pragma unroll 8
  for (int k = threadIdx.x; k < m.i; k += blockDim.x) {
    double val;
    if (cand(k, val) && val == val && !(m.v < __dsub_rn(val, eps))) bad++;
  }
  bad = block_sum_int(bad, smi);
  if (bad == 0) return m.i;
  if (threadIdx.x < 32) {  // literal replay, one ballot per 32 candidates (see block_hyst_min in select.cuh)
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

__global__ void __launch_bounds__(1024) k_enter(RevView v) {
  __shared__ MinIdx sm[32];
  __shared__ int smi[32];
  rev_pdl();
  RevState* st = v.st;
  if (st->status != LPR_RUNNING) return;
  const int n = v.n, m = v.m;
  const double EPS = 1e-9;
  // candidate keys (-rc for eligible non-basic variables, +inf otherwise) are materialised once into the
  // u scratch-free `rc`-sized staging area: all loads of a thread are issued before any use
  double* key = v.ekey;
  for (int k0 = 0; k0 < n + m; k0 += 1024 * 8) {
    double r[8];
    int ib[8];
#pragma unroll
    for (int q = 0; q < 8; q++) {
      const int k = k0 + q * 1024 + threadIdx.x;
      ib[q] = (k < n + m) ? v.isbasic[k] : 1;
      r[q] = (k < n + m) ? ((k < n) ? v.rc[k] : -v.y[k - n]) : 0.0;  // :100-102 slack rc = -y_k
    }
#pragma unroll
    for (int q = 0; q < 8; q++) {
      const int k = k0 + q * 1024 + threadIdx.x;
      if (k < n + m) key[k] = (!ib[q] && r[q] > EPS) ? -r[q] : kInfD;
    }
  }
  __syncthreads();
  // val = -rc so that "rc > best + EPS" becomes "val < best - EPS"; first candidate always accepted
  int e = rev_hyst_min(n + m, [&](int k, double& val) {
    val = key[k];
    return val < kInfD;
  }, kInfD, EPS, sm, smi);
  // interleave (a_e, b) for the direction pass; slack columns need no gather
  if (e >= 0 && e < n) {
    for (int i0 = 0; i0 < m; i0 += 1024 * 8) {  // strided DRAM gather: 8 independent loads in flight per thread
      double t[8];
#pragma unroll
      for (int q = 0; q < 8; q++) {
        const int i = i0 + q * 1024 + threadIdx.x;
        t[q] = (i < m) ? v.A[(size_t)i * v.ldA + e] : 0.0;
      }
#pragma unroll
      for (int q = 0; q < 8; q++) {
        const int i = i0 + q * 1024 + threadIdx.x;
        if (i < m) v.ab[2 * i] = t[q];
      }
    }
  }
  if (threadIdx.x == 0) {
    st->enter = e;
    st->need_final = (e < 0);
    st->do_update = 0;
  }
}

// k_dir: one warp per row: u_i = B^-1[i,:] . a_e  and  x_B[i] = B^-1[i,:] . b  (fixed tree)
__global__ void __launch_bounds__(kT) k_dir(RevView v) {
  rev_pdl();
  const RevState* st = v.st;
  if (st->status != LPR_RUNNING) return;
  const int e = st->enter;
  const int m = v.m, n = v.n;
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * kT + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * kT) >> 5;
  const bool structural = (e >= 0 && e < n);
  const int ldv = v.ldB >> 1;
  const double2* __restrict__ B2 = reinterpret_cast<const double2*>(v.Binv);
  const double4* __restrict__ ab4 = reinterpret_cast<const double4*>(v.ab);  // (a_j, b_j, a_j+1, b_j+1)
  const int mv = (m + 1) >> 1;
  for (int i = warp; i < m; i += nwarps) {
    double su = 0.0, sx = 0.0;
    const double2* row = B2 + (size_t)i * ldv;
    for (int c = lane; c < mv; c += 32) {
      double2 bv = row[c];
      double4 w = ab4[c];
      const bool two = (2 * c + 1 < m);
      if (structural) {
        su = __dadd_rn(su, __dmul_rn(bv.x, w.x));
        if (two) su = __dadd_rn(su, __dmul_rn(bv.y, w.z));
      }
      sx = __dadd_rn(sx, __dmul_rn(bv.x, w.y));
      if (two) sx = __dadd_rn(sx, __dmul_rn(bv.y, w.w));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      su = __dadd_rn(su, __shfl_xor_sync(0xffffffffu, su, o));
      sx = __dadd_rn(sx, __shfl_xor_sync(0xffffffffu, sx, o));
    }
    if (lane == 0) {
      v.xB[i] = sx;
      if (e >= 0) v.u[i] = structural ? su : v.Binv[(size_t)i * v.ldB + (e - n)];  // :149-151
    }
  }
}

__global__ void __launch_bounds__(1024) k_ratio(RevView v) {
  __shared__ MinIdx sm[32];
  __shared__ int smi[32];
  __shared__ int sh_r;
  rev_pdl();
  RevState* st = v.st;
  if (st->status != LPR_RUNNING) return;
  const int m = v.m, n = v.n;
  const double EPS = 1e-9;
  const int tid = threadIdx.x;
  const int e = st->enter;
  const long long iter = st->iter;
  // :90-91 any x_B < -EPS => "Infeasible basis"
  int neg = 0;
#pragma unroll 8
  for (int i = tid; i < m; i += blockDim.x)
    if (v.xB[i] < -EPS) neg++;
  neg = block_sum_int(neg, smi);
  if (neg) {
    if (tid == 0) { st->status = LPR_INFEASIBLE; st->do_update = 0; }
    return;
  }
  if (e < 0) {  // :124-146 optimal
    if (tid == 0) { st->status = LPR_OPTIMAL; st->do_update = 0; }
    return;
  }
  if (st->max_iter >= 0 && iter >= st->max_iter) {
    if (tid == 0) { st->status = LPR_ITER_LIMIT; st->do_update = 0; }
    return;
  }
  // :153-176 ratio test.  Simple case: the unique minimum is separated from every other ratio
  // by more than the tolerance window; otherwise thread 0 replays the scan literally.
  MinIdx mn = minidx_identity();
#pragma unroll 8
  for (int i = tid; i < m; i += blockDim.x) {
    double ui = v.u[i];
    if (ui > EPS) {
      double r = ddiv(v.xB[i], ui);
      if (r == r) mn = minidx_combine(mn, MinIdx{r, i});
    }
  }
  mn = block_minidx(mn, sm);
  if (mn.i == INT_MAX || !(mn.v < DBL_MAX)) {
    // no candidate (or only non-finite ratios: replay decides)
    if (mn.i == INT_MAX) {
      if (tid == 0) { st->status = LPR_UNBOUNDED; st->do_update = 0; }  // :178-179
      return;
    }
  }
  int bad = 0;
#pragma unroll 8
  for (int i = tid; i < m; i += blockDim.x) {
    double ui = v.u[i];
    if (i != mn.i && ui > EPS) {
      double r = ddiv(v.xB[i], ui);
      if (!(mn.v < __dsub_rn(r, EPS)) || fabs(__dsub_rn(r, mn.v)) <= EPS || !(r == r)) bad++;
    }
  }
  bad = block_sum_int(bad, smi);
  if (tid == 0) {
    int leave = mn.i;
    if (bad || !(mn.v < DBL_MAX)) {
      leave = -1;
      double best = DBL_MAX;
      for (int i = 0; i < m; i++) {
        double ui = v.u[i];
        if (ui > EPS) {
          double r = ddiv(v.xB[i], ui);
          if (r < __dsub_rn(best, EPS) ||
              (fabs(__dsub_rn(r, best)) <= EPS && (leave == -1 || v.basis[i] < v.basis[leave]))) {
            best = r;
            leave = i;
          }
        }
      }
    }
    sh_r = leave;
  }
  __syncthreads();
  const int r = sh_r;
  if (r < 0) {
    if (tid == 0) { st->status = LPR_UNBOUNDED; st->do_update = 0; }
    return;
  }
  const double pivot = v.u[r];
  if (fabs(pivot) < EPS) {  // :267
    if (tid == 0) { st->status = LPR_PIVOT_TOO_SMALL; st->do_update = 0; }
    return;
  }
  // eta column :269-272 and a copy of the pivot row of B^-1
  for (int i0 = 0; i0 < v.ldB; i0 += 1024 * 8) {
    double tu[8], tb[8];
#pragma unroll
    for (int q = 0; q < 8; q++) {
      const int i = i0 + q * 1024 + tid;
      tu[q] = (i < m) ? v.u[i] : 0.0;
      tb[q] = (i < m) ? v.Binv[(size_t)r * v.ldB + i] : 0.0;
    }
#pragma unroll
    for (int q = 0; q < 8; q++) {
      const int i = i0 + q * 1024 + tid;
      if (i < m) v.ecoef[i] = (i == r) ? ddiv(1.0, pivot) : ddiv(-tu[q], pivot);
      if (i < v.ldB) v.brow[i] = tb[q];
    }
  }
  if (tid == 0) {
    const int leaveVar = v.basis[r];
    if (v.log && iter < v.log_cap) {
      v.log[3 * iter + 0] = r;
      v.log[3 * iter + 1] = e;
      v.log[3 * iter + 2] = leaveVar;
    }
    v.basis[r] = e;  // :194-212
    v.isbasic[e] = 1;
    v.isbasic[leaveVar] = 0;
    v.cB[r] = (e < n) ? v.c[e] : 0.0;
    st->leave_row = r;
    st->leave_var = leaveVar;
    st->pivot = pivot;
    st->do_update = 1;
    st->iter = iter + 1;
  }
}

// k_update: grid (colTiles, YS).  B^-1 <- E B^-1 exactly as MultiplyMatrices (:426-441) does it
// (k ascending, |E_ik| < 1e-9 skipped, accumulation into a zero matrix), and the partial sums of
// the next dual vector y' = c_B' B^-1' over this CTA's row range.
__global__ void __launch_bounds__(kT) k_update(RevView v) {
  rev_pdl();
  const RevState* st = v.st;
  if (!st->do_update || st->status != LPR_RUNNING) return;
  const int r = st->leave_row;
  const int m = v.m;
  const double EPS = 1e-9;
  const int ldv = v.ldB >> 1;
  const int chunk = blockIdx.x * kT + threadIdx.x;
  const int rs = blockIdx.y;
  const int r0 = (int)((long long)m * rs / v.YS), r1 = (int)((long long)m * (rs + 1) / v.YS);
  if (chunk >= ldv) return;
  double2* __restrict__ B2 = reinterpret_cast<double2*>(v.Binv) + chunk;
  const double2 br = reinterpret_cast<const double2*>(v.brow)[chunk];
  double2 acc = make_double2(0.0, 0.0);
  constexpr int U = 8;
  for (int q = r0; q < r1; q += U) {
    double ei[U], cb[U];
    double2 x[U];
#pragma unroll
    for (int k = 0; k < U; k++) {
      const int row = q + k;
      if (row < r1) {
        ei[k] = v.ecoef[row];
        cb[k] = v.cB[row];
        x[k] = B2[(size_t)row * ldv];
      }
    }
#pragma unroll
    for (int k = 0; k < U; k++) {
      const int row = q + k;
      if (row < r1) {
        double2 y;
        const bool use = !(fabs(ei[k]) < EPS);
        if (row == r) {
          if (use) {
            y.x = __dadd_rn(0.0, __dmul_rn(ei[k], br.x));
            y.y = __dadd_rn(0.0, __dmul_rn(ei[k], br.y));
          } else {
            y.x = 0.0;
            y.y = 0.0;
          }
        } else {
          y.x = __dadd_rn(0.0, x[k].x);  // 0.0 + 1.0*B[i][j]
          y.y = __dadd_rn(0.0, x[k].y);
          if (use) {                      // (+) e_i * B[r][j]; fp addition commutes so i<r / i>r agree
            y.x = __dadd_rn(y.x, __dmul_rn(ei[k], br.x));
            y.y = __dadd_rn(y.y, __dmul_rn(ei[k], br.y));
          }
        }
        B2[(size_t)row * ldv] = y;
        if (cb[k] != 0.0 || v.dense) {  // y'_j += cB'_i * B'[i][j]  (:412-424)
          acc.x = __dadd_rn(acc.x, __dmul_rn(cb[k], y.x));
          acc.y = __dadd_rn(acc.y, __dmul_rn(cb[k], y.y));
        }
      }
    }
  }
  reinterpret_cast<double2*>(v.ypart)[(size_t)rs * ldv + chunk] = acc;
}

// y partials from the current B^-1 (after a refactorisation): same summation shape as k_update
__global__ void __launch_bounds__(kT) k_ypart(RevView v) {
  const int m = v.m;
  const int ldv = v.ldB >> 1;
  const int chunk = blockIdx.x * kT + threadIdx.x;
  const int rs = blockIdx.y;
  const int r0 = (int)((long long)m * rs / v.YS), r1 = (int)((long long)m * (rs + 1) / v.YS);
  if (chunk >= ldv) return;
  const double2* __restrict__ B2 = reinterpret_cast<const double2*>(v.Binv) + chunk;
  double2 acc = make_double2(0.0, 0.0);
  for (int row = r0; row < r1; row++) {
    const double cb = v.cB[row];
    if (cb != 0.0 || v.dense) {
      const double2 x = B2[(size_t)row * ldv];
      acc.x = __dadd_rn(acc.x, __dmul_rn(cb, x.x));
      acc.y = __dadd_rn(acc.y, __dmul_rn(cb, x.y));
    }
  }
  reinterpret_cast<double2*>(v.ypart)[(size_t)rs * ldv + chunk] = acc;
}

__global__ void k_rev_reset(RevState* st, long long max_iter) {
  st->status = LPR_RUNNING;
  st->enter = -1;
  st->leave_row = -1;
  st->leave_var = -1;
  st->do_update = 0;
  st->need_final = 0;
  st->iter = 0;
  st->max_iter = max_iter;
  st->pivot = 0.0;
}

// (re)initialise the slack basis: B^-1 = I, cB = 0, y = 0 (:63-79)
__global__ void k_rev_init(RevView v) {
  const int m = v.m, n = v.n;
  const size_t total = (size_t)m * v.ldB;
  for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += (size_t)gridDim.x * blockDim.x) {
    const int i = (int)(k / v.ldB), j = (int)(k % v.ldB);
    v.Binv[k] = (i == j) ? 1.0 : 0.0;
  }
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x) {
    v.cB[i] = 0.0;
    v.y[i] = 0.0;
    v.xB[i] = 0.0;
    v.u[i] = 0.0;
    v.basis[i] = n + i;
    v.ab[2 * i] = 0.0;
    v.ab[2 * i + 1] = v.b[i];
  }
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n + m; k += gridDim.x * blockDim.x)
    v.isbasic[k] = (k >= n) ? 1 : 0;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    v.ab[2 * m] = 0.0;  // padding pair for odd m
    v.ab[2 * m + 1] = 0.0;
  }
}

__global__ void k_rev_gen(double* A, int ldA, double* b, double* c, double* c_orig, int m, int n, uint64_t seed) {
  const int row = blockIdx.y;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < ldA; j += gridDim.x * blockDim.x) {
    A[(size_t)row * ldA + j] = (j < n) ? 0.1 + u01(seed, 0, (uint64_t)row * (uint64_t)n + (uint64_t)j) : 0.0;
    if (row == 0 && j < n) {
      double cj = 1.0 + u01(seed, 2, (uint64_t)j);
      c[j] = cj;
      c_orig[j] = cj;
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) b[row] = ((double)n / 4.0) * (1.0 + u01(seed, 1, (uint64_t)row));
}

// x_B = B^-1 b alone (same warp-per-row tree as k_dir, so the value is the one the next k_dir will produce): the
// POST-pivot basic solution of the per-iteration snapshot (:218)
__global__ void __launch_bounds__(kT) k_xb_only(RevView v) {
  const int m = v.m;
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * kT + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * kT) >> 5;
  const int ldv = v.ldB >> 1;
  const double2* __restrict__ B2 = reinterpret_cast<const double2*>(v.Binv);
  const double4* __restrict__ ab4 = reinterpret_cast<const double4*>(v.ab);
  const int mv = (m + 1) >> 1;
  for (int i = warp; i < m; i += nwarps) {
    double sx = 0.0;
    const double2* row = B2 + (size_t)i * ldv;
    for (int c = lane; c < mv; c += 32) {
      double2 bv = row[c];
      double4 w = ab4[c];
      sx = __dadd_rn(sx, __dmul_rn(bv.x, w.y));
      if (2 * c + 1 < m) sx = __dadd_rn(sx, __dmul_rn(bv.y, w.w));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sx = __dadd_rn(sx, __shfl_xor_sync(0xffffffffu, sx, o));
    if (lane == 0) v.xB[i] = sx;
  }
}

// MultiplyMatrices(BInverse, A) of CaptureSnapshot (:360, :426-441), literally: k ascending, |B^-1[i,k]| < 1e-9 skipped,
// multiply then add into a zero -- one thread per output element, so the printed tableau is the reference's bit for bit
// given the same B^-1.  O(m^2 n): only ever launched for a snapshot (lazy), never inside Solve().
__global__ void __launch_bounds__(256) k_binv_a(RevView v, double* __restrict__ out, int ldo) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = blockIdx.y;
  if (j >= v.n) return;
  const double* __restrict__ brow = v.Binv + (size_t)i * v.ldB;
  double s = 0.0;
  for (int k = 0; k < v.m; k++) {
    const double aik = brow[k];
    if (fabs(aik) < 1e-9) continue;
    s = __dadd_rn(s, __dmul_rn(aik, v.A[(size_t)k * v.ldA + j]));
  }
  out[(size_t)i * ldo + j] = s;
}

// SolutionVector / FinalZ :277-287 (x = max(0, x_B) for basic structurals; z = c_orig . x sequential)
__global__ void k_rev_solution(RevView v, const double* c_orig, double* x, double* z) {
  const int n = v.n, m = v.m;
  for (int j = threadIdx.x; j < n; j += blockDim.x) x[j] = 0.0;
  __syncthreads();
  for (int i = threadIdx.x; i < m; i += blockDim.x) {
    int b = v.basis[i];
    if (b < n) {
      double xb = v.xB[i];
      x[b] = (0.0 > xb) ? 0.0 : xb;  // Math.Max(0.0, xB[i])
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int j = 0; j < n; j++) s = __dadd_rn(s, __dmul_rn(c_orig[j], x[j]));  // Dot :443-448
    *z = s;
  }
}

}  // namespace lpr

#include "refactor.cuh"
using namespace lpr;

struct lpr_rev {
  int device = 0, sms = 148;
  cudaStream_t stream = nullptr;
  int m = 0, n = 0, ldA = 0, ldB = 0, PS = 1, YS = 1;
  double* ekey = nullptr;
  double *A = nullptr, *Binv = nullptr, *b = nullptr, *c = nullptr, *c_orig = nullptr, *cB = nullptr,
         *xB = nullptr, *y = nullptr, *rc = nullptr, *u = nullptr, *ab = nullptr, *ecoef = nullptr,
         *brow = nullptr, *ppart = nullptr, *ypart = nullptr, *x = nullptr, *z = nullptr;
  int *basis = nullptr, *isbasic = nullptr, *log = nullptr;
  long long log_cap = 0;
  RevState* st = nullptr;
  RevState* st_host = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, evb[2] = {nullptr, nullptr};
  float last_ms = 0.f, last_refactor_ms = 0.f;
  double last_refactor_residual = 0.0, last_refactor_residual_after = 0.0, last_refactor_flops = 0.0;
  int last_refactor_path = 0;  // 1 = Newton-Schulz refresh, 2 = full blocked Gauss-Jordan inversion
  RefactorWs rws;              // refactorisation workspace, allocated by the first refactorisation
  bool solved = false;
  // lpr_rev_begin / lpr_rev_step: host copies of the PRE-pivot quantities of the last step, which CaptureSnapshot
  // (:294-387) prints beside the post-pivot tableau
  struct Trace {
    bool valid = false, optimal = false;
    long long iteration = 0;
    int enter = -1, leave_row = -1, leave_var = -1;
    double rc_enter_pre = 0.0;
    std::vector<double> u_pre, xb_pre;
  } trace;
  bool stepping = false, is_min = false;
  RevView view() const {
    RevView v;
    v.m = m; v.n = n; v.ldA = ldA; v.ldB = ldB; v.A = A; v.Binv = Binv; v.b = b; v.c = c; v.cB = cB;
    v.xB = xB; v.y = y; v.rc = rc; v.u = u; v.ekey = ekey; v.ab = ab; v.ecoef = ecoef; v.brow = brow; v.ppart = ppart;
    v.ypart = ypart; v.PS = PS; v.YS = YS; v.basis = basis; v.isbasic = isbasic; v.st = st; v.log = log;
    v.log_cap = log_cap;
    const char* de = getenv("LPR_REV_DENSE");  // read per call: bench.py measures both settings in one process
    v.dense = de ? atoi(de) : 0;
    return v;
  }
};

static int rev_alloc(int device, int m, int n, lpr_rev** out) {
  if (!out) return fail(LPR_E_BADARG, "out is null");
  *out = nullptr;
  if (m < 1 || n < 1) return fail(LPR_E_BADARG, "bad shape m=%d n=%d", m, n);
  int rc = select_device(device);
  if (rc) return rc;
  lpr_rev* h = new (std::nothrow) lpr_rev();
  if (!h) return fail(LPR_E_NOMEM, "host allocation failed");
  h->device = device;
  h->sms = sm_count(device);
  h->m = m;
  h->n = n;
  h->ldA = round_up(n, 16);
  h->ldB = round_up(m, 16);
  // row splits so that (column tiles x splits) is about 4 CTAs per SM
  // ... and, when a nearby split count exists, so that the CTA count is a whole multiple of the SM count: 32
  // column tiles x 19 splits = 608 CTAs on 148 SMs x 4 resident leaves 16 CTAs for a second, nearly empty
  // wave (k_price ran at 4.8 TB/s); 32 x 37 = 1184 = 2 full waves
  auto splits = [&](int ld) {
    int tiles = (ld / 2 + kT - 1) / kT;
    int s = std::max(1, (h->sms * 4 + tiles - 1) / tiles);
    for (int t = s; t <= 2 * s + 1; t++)
      if (((long long)tiles * t) % h->sms == 0) { s = t; break; }
    return std::min(s, std::max(1, m / 8));
  };
  h->PS = splits(h->ldA);
  h->YS = splits(h->ldB);
  cudaError_t e;
#define TRY(x)                                                                                   \
  if ((e = (x)) != cudaSuccess) {                                                                \
    lpr_rev_destroy(h);                                                                          \
    return fail(e == cudaErrorMemoryAllocation ? LPR_E_NOMEM : LPR_E_CUDA, "%s failed: %s", #x, \
                cudaGetErrorString(e));                                                          \
  }
  TRY(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  TRY(cudaMalloc(&h->A, sizeof(double) * (size_t)m * h->ldA));
  TRY(cudaMalloc(&h->Binv, sizeof(double) * (size_t)m * h->ldB));
  TRY(cudaMalloc(&h->b, sizeof(double) * m));
  TRY(cudaMalloc(&h->c, sizeof(double) * n));
  TRY(cudaMalloc(&h->c_orig, sizeof(double) * n));
  TRY(cudaMalloc(&h->cB, sizeof(double) * m));
  TRY(cudaMalloc(&h->xB, sizeof(double) * m));
  TRY(cudaMalloc(&h->y, sizeof(double) * m));
  TRY(cudaMalloc(&h->rc, sizeof(double) * n));
  TRY(cudaMalloc(&h->u, sizeof(double) * m));
  TRY(cudaMalloc(&h->ekey, sizeof(double) * ((size_t)n + m)));
  TRY(cudaMalloc(&h->ab, sizeof(double) * (2 * (size_t)m + 4)));
  TRY(cudaMalloc(&h->ecoef, sizeof(double) * m));
  TRY(cudaMalloc(&h->brow, sizeof(double) * h->ldB));
  TRY(cudaMalloc(&h->ppart, sizeof(double) * (size_t)h->PS * h->ldA));
  TRY(cudaMalloc(&h->ypart, sizeof(double) * (size_t)h->YS * h->ldB));
  TRY(cudaMalloc(&h->x, sizeof(double) * n));
  TRY(cudaMalloc(&h->z, sizeof(double)));
  TRY(cudaMalloc(&h->basis, sizeof(int) * m));
  TRY(cudaMalloc(&h->isbasic, sizeof(int) * ((size_t)n + m)));
  TRY(cudaMalloc(&h->st, sizeof(RevState)));
  TRY(cudaMallocHost(&h->st_host, sizeof(RevState) * 2));
  TRY(cudaEventCreate(&h->ev0));
  TRY(cudaEventCreate(&h->ev1));
  TRY(cudaEventCreateWithFlags(&h->evb[0], cudaEventDisableTiming));
  TRY(cudaEventCreateWithFlags(&h->evb[1], cudaEventDisableTiming));
#undef TRY
  *out = h;
  return LPR_OK;
}

static int rev_ensure_log(lpr_rev* h, long long cap) {
  if (cap <= h->log_cap) return LPR_OK;
  if (h->log) cudaFree(h->log);
  h->log = nullptr;
  h->log_cap = 0;
  LPR_CUDA(cudaMalloc(&h->log, sizeof(int) * 3 * (size_t)cap));
  h->log_cap = cap;
  return LPR_OK;
}

// launch with the programmatic-stream-serialization attribute (PDL); every kernel of the chain starts with
// griddepcontrol.wait so correctness does not depend on it
static bool rev_launch(void (*kernel)(RevView), dim3 grid, int block, cudaStream_t stream, const RevView& v) {
  static const int pdl = getenv("LPR_PDL") ? atoi(getenv("LPR_PDL")) : 1;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(block);
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  count_launch();
  return cudaLaunchKernelEx(&cfg, kernel, v) != cudaSuccess;
}

static int rev_init_basis(lpr_rev* h) {
  k_rev_init<<<h->sms * 4, 256, 0, h->stream>>>(h->view());
  LPR_LAUNCH_CHECK();
  return LPR_OK;
}

extern "C" {

int lpr_rev_destroy(lpr_rev* h) {
  if (!h) return LPR_OK;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  refactor_ws_free(h->rws);
  double* d[] = {h->A, h->Binv, h->b, h->c, h->c_orig, h->cB, h->xB, h->y, h->rc, h->u, h->ab, h->ecoef,
                 h->brow, h->ppart, h->ypart, h->x, h->z};
  for (double* p : d) cudaFree(p);
  cudaFree(h->ekey);
  cudaFree(h->basis);
  cudaFree(h->isbasic);
  cudaFree(h->log);
  cudaFree(h->st);
  if (h->st_host) cudaFreeHost(h->st_host);
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  if (h->evb[0]) cudaEventDestroy(h->evb[0]);
  if (h->evb[1]) cudaEventDestroy(h->evb[1]);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return LPR_OK;
}

int lpr_rev_create(int device, int m, int n, const double* A, const double* b, const double* c,
                   int is_minimization, lpr_rev** out) {
  if (!A || !b || !c) return fail(LPR_E_BADARG, "null model array");
  lpr_rev* h = nullptr;
  int rc = rev_alloc(device, m, n, &h);
  if (rc) return rc;
  std::vector<double> cc(c, c + n);
  if (is_minimization)
    for (auto& x : cc) x = -x;  // :51
  cudaError_t e = cudaMemsetAsync(h->A, 0, sizeof(double) * (size_t)m * h->ldA, h->stream);
  if (e == cudaSuccess)
    e = cudaMemcpy2DAsync(h->A, sizeof(double) * h->ldA, A, sizeof(double) * n, sizeof(double) * n, m,
                          cudaMemcpyHostToDevice, h->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(h->b, b, sizeof(double) * m, cudaMemcpyHostToDevice, h->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(h->c, cc.data(), sizeof(double) * n, cudaMemcpyHostToDevice, h->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(h->c_orig, c, sizeof(double) * n, cudaMemcpyHostToDevice, h->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  if (e != cudaSuccess) {
    lpr_rev_destroy(h);
    return fail(LPR_E_CUDA, "model upload failed: %s", cudaGetErrorString(e));
  }
  rc = rev_init_basis(h);
  if (rc == LPR_OK && cudaStreamSynchronize(h->stream) != cudaSuccess) rc = fail(LPR_E_CUDA, "init failed");
  if (rc) {
    lpr_rev_destroy(h);
    return rc;
  }
  h->is_min = is_minimization != 0;
  *out = h;
  return LPR_OK;
}

int lpr_rev_create_dense_lp(int device, uint64_t seed, int m, int n, lpr_rev** out) {
  lpr_rev* h = nullptr;
  int rc = rev_alloc(device, m, n, &h);
  if (rc) return rc;
  dim3 grid(std::max(1, std::min(64, (h->ldA + 255) / 256)), m);
  k_rev_gen<<<grid, 256, 0, h->stream>>>(h->A, h->ldA, h->b, h->c, h->c_orig, m, n, seed);
  count_launch();
  rc = rev_init_basis(h);
  cudaError_t e = cudaStreamSynchronize(h->stream);
  if (rc == LPR_OK && e != cudaSuccess) rc = fail(LPR_E_CUDA, "dense LP generation failed: %s", cudaGetErrorString(e));
  if (rc) {
    lpr_rev_destroy(h);
    return rc;
  }
  *out = h;
  return LPR_OK;
}

int lpr_rev_solve(lpr_rev* h, int64_t max_iter, int refactor_every, int* status, int64_t* n_iter, int* log,
                  int64_t log_cap) {
  if (!h) return fail(LPR_E_BADARG, "null handle");
  int rc = select_device(h->device);
  if (rc) return rc;
  if (log && log_cap > 0) {
    long long want = std::min<long long>(log_cap, 1LL << 24);
    if (max_iter >= 0) want = std::min<long long>(want, max_iter + 1);
    rc = rev_ensure_log(h, std::max<long long>(1, want));
    if (rc) return rc;
  }
  lpr_rev hv = *h;
  if (!(log && log_cap > 0)) { hv.log = nullptr; hv.log_cap = 0; }
  RevView v = hv.view();
  const int m = h->m, n = h->n;
  dim3 gp((h->ldA / 2 + kT - 1) / kT, h->PS), gu((h->ldB / 2 + kT - 1) / kT, h->YS);
  const int gdir = h->sms * 8;
  static const int batch_env = getenv("LPR_REV_BATCH") ? atoi(getenv("LPR_REV_BATCH")) : 16;
  const int batch = std::max(1, batch_env);

  LPR_CUDA(cudaEventRecord(h->ev0, h->stream));
  rc = rev_init_basis(h);  // Solve() always starts from the slack basis (:63-79)
  if (rc) return rc;
  k_rev_reset<<<1, 1, 0, h->stream>>>(h->st, (long long)max_iter);
  LPR_LAUNCH_CHECK();
  int slot = 0, pending = 0, bsize = std::min(batch, 2);
  long long launched = 0;
  long long next_refactor = refactor_every > 0 ? refactor_every : -1;
  bool finished = false;
  while (!finished) {
    for (int q = 0; q < bsize; q++) {
      if (rev_launch(k_price, gp, kT, h->stream, v) || rev_launch(k_rc, dim3((n + kT - 1) / kT), kT, h->stream, v) ||
          rev_launch(k_enter, dim3(1), 1024, h->stream, v) || rev_launch(k_dir, dim3(gdir), kT, h->stream, v) ||
          rev_launch(k_ratio, dim3(1), 1024, h->stream, v) || rev_launch(k_update, gu, kT, h->stream, v) ||
          rev_launch(k_y, dim3((m + kT - 1) / kT), kT, h->stream, v))
        return fail(LPR_E_CUDA, "revised simplex kernel launch failed: %s", cudaGetErrorString(cudaGetLastError()));
      launched++;
      if (next_refactor > 0 && launched == next_refactor) {
        // periodic refactorisation needs the host to know the run is still alive
        LPR_CUDA(cudaMemcpyAsync(&h->st_host[0], h->st, sizeof(RevState), cudaMemcpyDeviceToHost, h->stream));
        LPR_CUDA(cudaStreamSynchronize(h->stream));
        pending = 0;
        if (h->st_host[0].status != LPR_RUNNING) {
          // the run ended before this refactorisation point: stop here.  (Resetting `pending` and carrying on used
          // to starve the status check below whenever every batch contained a refactorisation point: an endless
          // stream of no-op launches once refactor_every <= batch.)
          finished = true;
          break;
        }
        rc = lpr_rev_refactor(h);
        if (rc) return rc;
        next_refactor += refactor_every;
      }
    }
    if (finished) break;
    LPR_CUDA(cudaMemcpyAsync(&h->st_host[slot], h->st, sizeof(RevState), cudaMemcpyDeviceToHost, h->stream));
    LPR_CUDA(cudaEventRecord(h->evb[slot], h->stream));
    pending++;
    if (pending == 2 || bsize < batch) {
      const int old = (pending == 2) ? (slot ^ 1) : slot;
      LPR_CUDA(cudaEventSynchronize(h->evb[old]));
      pending--;
      if (h->st_host[old].status != LPR_RUNNING) break;
    }
    slot ^= 1;
    bsize = std::min(batch, bsize * 2);
  }
  k_rev_solution<<<1, 1024, 0, h->stream>>>(v, h->c_orig, h->x, h->z);
  LPR_LAUNCH_CHECK();
  LPR_CUDA(cudaEventRecord(h->ev1, h->stream));
  LPR_CUDA(cudaMemcpyAsync(&h->st_host[0], h->st, sizeof(RevState), cudaMemcpyDeviceToHost, h->stream));
  LPR_CUDA(cudaStreamSynchronize(h->stream));
  LPR_CUDA(cudaEventElapsedTime(&h->last_ms, h->ev0, h->ev1));
  const long long it = h->st_host[0].iter;
  if (status) *status = h->st_host[0].status;
  if (n_iter) *n_iter = it;
  if (log && log_cap > 0 && it > 0) {
    long long cnt = std::min<long long>(std::min<long long>(it, log_cap), h->log_cap);
    LPR_CUDA(cudaMemcpy(log, h->log, sizeof(int) * 3 * (size_t)cnt, cudaMemcpyDeviceToHost));
  }
  h->solved = true;
  return LPR_OK;
}

// ---- one iteration at a time, for snapshot-accurate tracing (SURVEY 8b "Snapshots", row b9) ----------------------------
int lpr_rev_begin(lpr_rev* h) {
  if (!h) return fail(LPR_E_BADARG, "null handle");
  int rc = select_device(h->device);
  if (rc) return rc;
  if ((rc = rev_init_basis(h))) return rc;  // Solve() always starts from the slack basis (:63-79)
  k_rev_reset<<<1, 1, 0, h->stream>>>(h->st, -1LL);
  LPR_LAUNCH_CHECK();
  LPR_CUDA(cudaStreamSynchronize(h->stream));
  h->stepping = true;
  h->trace = lpr_rev::Trace();
  return LPR_OK;
}

// One pass of the while loop of Solve() (:86-250).  *status: RUNNING after a pivot ("Iteration k" snapshot available),
// OPTIMAL when no entering variable exists ("Optimal" snapshot available, x and z ready), INFEASIBLE / UNBOUNDED /
// PIVOT_TOO_SMALL where the reference throws (:91, :179, :267).
int lpr_rev_step(lpr_rev* h, int* status, int* enter, int* leave_row, int* leave_var) {
  if (!h) return fail(LPR_E_BADARG, "null handle");
  if (!h->stepping) return fail(LPR_E_STATE, "lpr_rev_step needs lpr_rev_begin first");
  int rc = select_device(h->device);
  if (rc) return rc;
  lpr_rev hv = *h;
  hv.log = nullptr;
  hv.log_cap = 0;
  RevView v = hv.view();
  const int m = h->m, n = h->n;
  dim3 gp((h->ldA / 2 + kT - 1) / kT, h->PS), gu((h->ldB / 2 + kT - 1) / kT, h->YS);
  if (rev_launch(k_price, gp, kT, h->stream, v) || rev_launch(k_rc, dim3((n + kT - 1) / kT), kT, h->stream, v) ||
      rev_launch(k_enter, dim3(1), 1024, h->stream, v) || rev_launch(k_dir, dim3(h->sms * 8), kT, h->stream, v))
    return fail(LPR_E_CUDA, "revised simplex kernel launch failed: %s", cudaGetErrorString(cudaGetLastError()));
  // pre-pivot x_B (k_ratio does not change it, but the snapshot pass overwrites it with the post-pivot one)
  lpr_rev::Trace& tr = h->trace;
  tr.valid = false;
  tr.u_pre.assign(m, 0.0);
  tr.xb_pre.assign(m, 0.0);
  LPR_CUDA(cudaMemcpyAsync(tr.xb_pre.data(), h->xB, sizeof(double) * m, cudaMemcpyDeviceToHost, h->stream));
  if (rev_launch(k_ratio, dim3(1), 1024, h->stream, v))
    return fail(LPR_E_CUDA, "revised simplex kernel launch failed: %s", cudaGetErrorString(cudaGetLastError()));
  LPR_CUDA(cudaMemcpyAsync(&h->st_host[0], h->st, sizeof(RevState), cudaMemcpyDeviceToHost, h->stream));
  LPR_CUDA(cudaStreamSynchronize(h->stream));
  const RevState st = h->st_host[0];
  if (status) *status = st.status;
  if (enter) *enter = st.enter;
  if (leave_row) *leave_row = st.status == LPR_RUNNING ? st.leave_row : -1;
  if (leave_var) *leave_var = st.status == LPR_RUNNING ? st.leave_var : -1;
  if (st.status == LPR_RUNNING && st.do_update) {
    const int e = st.enter;
    LPR_CUDA(cudaMemcpyAsync(tr.u_pre.data(), h->u, sizeof(double) * m, cudaMemcpyDeviceToHost, h->stream));
    double t = 0.0;  // reduced cost of the entering variable BEFORE the pivot: rc_e, or -y_k for slack k (:185-187)
    LPR_CUDA(cudaMemcpyAsync(&t, e < n ? h->rc + e : h->y + (e - n), sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    LPR_CUDA(cudaStreamSynchronize(h->stream));
    tr.rc_enter_pre = e < n ? t : -t;
    if (rev_launch(k_update, gu, kT, h->stream, v) || rev_launch(k_y, dim3((m + kT - 1) / kT), kT, h->stream, v))
      return fail(LPR_E_CUDA, "revised simplex kernel launch failed: %s", cudaGetErrorString(cudaGetLastError()));
    LPR_CUDA(cudaStreamSynchronize(h->stream));
    tr.valid = true;
    tr.optimal = false;
    tr.iteration = st.iter;  // already advanced: "Iteration {iteration + 1}" (:231)
    tr.enter = e;
    tr.leave_row = st.leave_row;
    tr.leave_var = st.leave_var;
  } else if (st.status == LPR_OPTIMAL) {
    k_rev_solution<<<1, 1024, 0, h->stream>>>(v, h->c_orig, h->x, h->z);
    LPR_LAUNCH_CHECK();
    LPR_CUDA(cudaStreamSynchronize(h->stream));
    tr.valid = true;
    tr.optimal = true;
    tr.iteration = st.iter;
    tr.enter = tr.leave_row = tr.leave_var = -1;
    h->solved = true;
    h->stepping = false;
  } else {
    h->stepping = false;
  }
  return LPR_OK;
}

}  // extern "C"
namespace lpr {
void format_revised_snapshot(std::string& sb, const char* title, bool is_min, int m, int n, const double* y,
                             const double* rcX, const double* rcS, int entering, double rc_pre, const double* u_pre,
                             const double* ratios_pre, const int* basis_pre, int leave_row, int leave_var_pre,
                             double z_working, double z_original, const double* BinvA, int64_t ldBA,
                             const double* Binv, int64_t ldB, const double* xB, const int* basis_post);
std::string& thread_text();
}
extern "C" {

// CaptureSnapshot (:294-387) for the step lpr_rev_step has just made: post-pivot duals / reduced costs / B^-1 A | B^-1 |
// RHS table beside the pre-pivot direction and ratio test.  Everything O(m n) and larger is computed HERE, on demand
// (the reference pays MultiplyMatrices(B^-1, A) = O(m^2 n) on every iteration whether or not anybody reads the text).
int lpr_rev_format_snapshot(lpr_rev* h, const char** text, int64_t* len) {
  if (!h || !text) return fail(LPR_E_BADARG, "null argument");
  lpr_rev::Trace& tr = h->trace;
  if (!tr.valid) return fail(LPR_E_STATE, "no step to describe: call lpr_rev_step first");
  int rc = select_device(h->device);
  if (rc) return rc;
  const int m = h->m, n = h->n;
  if ((double)m * ((double)n + m) > 2.7e8) return fail(LPR_E_CAPACITY, "snapshot of a %d x %d model is too large to print", m, n);
  RevView v = h->view();
  dim3 gp((h->ldA / 2 + kT - 1) / kT, h->PS);
  if (!tr.optimal) {  // post-pivot x_B and reduced costs (:218-227); at the optimum the current ones are the post ones
    k_xb_only<<<h->sms * 8, kT, 0, h->stream>>>(v);
    LPR_LAUNCH_CHECK();
    if (rev_launch(k_price, gp, kT, h->stream, v) || rev_launch(k_rc, dim3((n + kT - 1) / kT), kT, h->stream, v))
      return fail(LPR_E_CUDA, "snapshot kernels failed: %s", cudaGetErrorString(cudaGetLastError()));
  }
  double* d_ba = nullptr;
  LPR_CUDA(cudaMalloc(&d_ba, sizeof(double) * (size_t)m * n));
  k_binv_a<<<dim3((n + 255) / 256, m), 256, 0, h->stream>>>(v, d_ba, n);
  count_launch();
  std::vector<double> y(m), rcx(n), xb(m), cb(m), corig(n), ba((size_t)m * n), binv((size_t)m * m), ratios(m);
  std::vector<int> basis(m), basis_pre(m);
  cudaError_t e = cudaMemcpyAsync(ba.data(), d_ba, sizeof(double) * (size_t)m * n, cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(y.data(), h->y, sizeof(double) * m, cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(rcx.data(), h->rc, sizeof(double) * n, cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(xb.data(), h->xB, sizeof(double) * m, cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(cb.data(), h->cB, sizeof(double) * m, cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(corig.data(), h->c_orig, sizeof(double) * n, cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(basis.data(), h->basis, sizeof(int) * m, cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess)
    e = cudaMemcpy2DAsync(binv.data(), sizeof(double) * m, h->Binv, sizeof(double) * h->ldB, sizeof(double) * m, m,
                          cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  cudaFree(d_ba);
  if (e != cudaSuccess) return fail(LPR_E_CUDA, "snapshot readback failed: %s", cudaGetErrorString(e));
  double zw = 0.0;  // Dot(cB, xB) :247 / :143, sequential like :443-448
  for (int i = 0; i < m; i++) zw += cb[i] * xb[i];
  std::vector<double> x(n, 0.0);  // ComputeOriginalZFromCurrentBasis :253-262
  for (int i = 0; i < m; i++)
    if (basis[i] < n) x[basis[i]] = (0.0 > xb[i]) ? 0.0 : xb[i];
  double zo = 0.0;
  for (int j = 0; j < n; j++) zo += corig[j] * x[j];
  basis_pre = basis;
  if (!tr.optimal) {
    basis_pre[tr.leave_row] = tr.leave_var;
    for (int i = 0; i < m; i++)  // :159-175
      ratios[i] = tr.u_pre[i] > 1e-9 ? tr.xb_pre[i] / tr.u_pre[i] : __builtin_huge_val();
  }
  std::string& sb = thread_text();
  sb.clear();
  const std::string title = tr.optimal ? std::string("Optimal") : "Iteration " + std::to_string(tr.iteration);
  format_revised_snapshot(sb, title.c_str(), h->is_min, m, n, y.data(), rcx.data(), nullptr, tr.optimal ? -1 : tr.enter,
                          tr.rc_enter_pre, tr.u_pre.data(), ratios.data(), basis_pre.data(), tr.leave_row, tr.leave_var,
                          zw, zo, ba.data(), n, binv.data(), m, xb.data(), basis.data());
  *text = sb.c_str();
  if (len) *len = (int64_t)sb.size();
  return LPR_OK;
}

// mode 0: Newton-Schulz refresh while max |I - B X| < 0.5, else the full refactorisation; 1: refresh only;
// 2: full refactorisation from the basis columns alone (blocked Gauss-Jordan, refactor.cu)
int lpr_rev_refactor_ex(lpr_rev* h, int mode) {
  if (!h) return fail(LPR_E_BADARG, "null handle");
  if (mode < 0 || mode > 2) return fail(LPR_E_BADARG, "refactorisation mode %d", mode);
  int rc = select_device(h->device);
  if (rc) return rc;
  if ((rc = refactor_ws_ensure(h->rws, h->m))) return rc;  // first call of a handle: allocated outside the timer
  LPR_CUDA(cudaEventRecord(h->ev0, h->stream));
  double res = 0.0, res_after = 0.0, flops = 0.0;
  int path = 0;
  rc = refactor_binv(h->stream, h->m, h->n, h->A, h->ldA, h->Binv, h->ldB, h->basis, h->rws, mode, &res, &res_after,
                     &flops, &path);
  if (rc == LPR_OK) {
    // the dual vector of the next pricing pass must come from the refreshed inverse
    RevView v = h->view();
    dim3 gu((h->ldB / 2 + kT - 1) / kT, h->YS);
    k_ypart<<<gu, kT, 0, h->stream>>>(v);
    count_launch();
    // k_y checks the status word: it is RUNNING or terminal; recompute unconditionally with a local copy
    k_y_force<<<(h->m + kT - 1) / kT, kT, 0, h->stream>>>(v);
    count_launch();
    cudaEventRecord(h->ev1, h->stream);
    if (cudaStreamSynchronize(h->stream) != cudaSuccess) rc = fail(LPR_E_CUDA, "refactor: y recompute failed");
    float ms = 0.f;
    cudaEventElapsedTime(&ms, h->ev0, h->ev1);
    h->last_refactor_ms = ms;
    h->last_refactor_residual = res;
    h->last_refactor_residual_after = res_after;
    h->last_refactor_flops = flops;
    h->last_refactor_path = path;
  }
  return rc;
}
int lpr_rev_refactor(lpr_rev* h) {
  static const int mode = getenv("LPR_REV_REFACTOR_MODE") ? atoi(getenv("LPR_REV_REFACTOR_MODE")) : 0;
  return lpr_rev_refactor_ex(h, mode);
}

#define REV_READ(name, field, count, type)                                                        \
  int name(lpr_rev* h, type* out) {                                                               \
    if (!h || !out) return fail(LPR_E_BADARG, "null argument");                                   \
    int rc = select_device(h->device);                                                            \
    if (rc) return rc;                                                                            \
    LPR_CUDA(cudaMemcpyAsync(out, h->field, sizeof(type) * (size_t)(count), cudaMemcpyDeviceToHost, h->stream)); \
    LPR_CUDA(cudaStreamSynchronize(h->stream));                                                   \
    return LPR_OK;                                                                                \
  }
REV_READ(lpr_rev_read_basis, basis, h->m, int)
REV_READ(lpr_rev_read_x, x, h->n, double)
REV_READ(lpr_rev_read_z, z, 1, double)
REV_READ(lpr_rev_read_y, y, h->m, double)
REV_READ(lpr_rev_read_xb, xB, h->m, double)
#undef REV_READ

int lpr_rev_read_binv(lpr_rev* h, double* binv) {
  if (!h || !binv) return fail(LPR_E_BADARG, "null argument");
  int rc = select_device(h->device);
  if (rc) return rc;
  LPR_CUDA(cudaMemcpy2DAsync(binv, sizeof(double) * h->m, h->Binv, sizeof(double) * h->ldB, sizeof(double) * h->m,
                             h->m, cudaMemcpyDeviceToHost, h->stream));
  LPR_CUDA(cudaStreamSynchronize(h->stream));
  return LPR_OK;
}
int lpr_rev_last_solve_ms(const lpr_rev* h, float* ms) {
  if (!h || !ms) return fail(LPR_E_BADARG, "null argument");
  *ms = h->last_ms;
  return LPR_OK;
}
int lpr_rev_last_refactor_ms(const lpr_rev* h, float* ms) {
  if (!h || !ms) return fail(LPR_E_BADARG, "null argument");
  *ms = h->last_refactor_ms;
  return LPR_OK;
}
int lpr_rev_last_refactor_info(const lpr_rev* h, double* residual, double* flops) {
  if (!h) return fail(LPR_E_BADARG, "null argument");
  if (residual) *residual = h->last_refactor_residual;
  if (flops) *flops = h->last_refactor_flops;
  return LPR_OK;
}
int lpr_rev_last_refactor_path(const lpr_rev* h, int* path, double* residual_after) {
  if (!h) return fail(LPR_E_BADARG, "null argument");
  if (path) *path = h->last_refactor_path;
  if (residual_after) *residual_after = h->last_refactor_residual_after;
  return LPR_OK;
}
// overwrite B^-1 (m x m, row major) -- warm starts, and the tests that damage the inverse to exercise the guard
int lpr_rev_write_binv(lpr_rev* h, const double* binv) {
  if (!h || !binv) return fail(LPR_E_BADARG, "null argument");
  int rc = select_device(h->device);
  if (rc) return rc;
  LPR_CUDA(cudaMemcpy2DAsync(h->Binv, sizeof(double) * h->ldB, binv, sizeof(double) * h->m, sizeof(double) * h->m, h->m,
                             cudaMemcpyHostToDevice, h->stream));
  LPR_CUDA(cudaStreamSynchronize(h->stream));
  return LPR_OK;
}

}  // extern "C"
