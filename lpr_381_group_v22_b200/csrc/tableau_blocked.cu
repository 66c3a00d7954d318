// tableau_blocked.cu -- delayed-update ("blocked") primal tableau simplex, bit-identical to the
// per-pivot path of PrimalSimplexSolver.cs:102-211.
//
// A pivot only needs ONE column (entering), ONE row (leaving) and the objective row / RHS column of the
// current tableau to be chosen.  Those can be produced from the tableau of K0 pivots ago plus the K pending
// rank-1 updates, applied to just that column / row in the original order:
//     x <- (i == p_u) ? prow_u[j] : x - (f_u[i] * prow_u[j])        u = 1..s
// which is exactly the sequence of roundings every element goes through in the reference's in-place
// Pivot (:193-211).  So the full tableau is swept once per K pivots instead of once per pivot, and each
// element still sees the same operations in the same order => same bits.  HBM traffic per pivot drops
// from 16*R*C to 16*R*C/K + O(K*(R+C)).
//   k_blk_select (ceil(ld/256) CTAs): entering column gathered from the stale tableau + pending updates,
//       min-ratio test, pivot row slice with pending updates, objective-row / RHS mirrors, next entering
//       column; the last CTA (atomic ticket) publishes the pivot.
//   k_blk_sweep: applies the s pending updates; thread owns one 16-byte column chunk (its s pivot-row
//       values live in registers) and walks down rows.
#include <cooperative_groups.h>

#include <algorithm>
#include <cstdlib>
#include <vector>

#include "blocked_apply.cuh"
#include "tableau.cuh"

namespace lpr {

constexpr int KMAX = 16;

struct BlkView {
  TabView v;
  double* PR;
  double* F;
  double* row0;
  double* rhs0;
  double* rhs1;
  int* pidx;
  int Rcap;
  MinIdx* cand;
  unsigned* ticket;
  long long* dbg;  // optional phase timestamps of the cluster select (LPR_BLK_TIMING=1), else null
};

__global__ void __launch_bounds__(kSelThreads) k_blk_init(BlkView b) {
  __shared__ MinIdx sm[32];
  const TabView& v = b.v;
  TabState* st = v.st;
  const int R = v.R, C = v.C, ld = v.ld;
  for (int j = threadIdx.x; j < ld; j += blockDim.x) b.row0[j] = (j < C) ? v.T[j] : 0.0;
  for (int i = threadIdx.x; i < R; i += blockDim.x) b.rhs0[i] = TAT(v.T, ld, i, C - 1);
  // FindEnteringVariable :152-167
  int e = block_first_min(C - 1, [&](int j, double& val) { val = v.T[j]; return val < 0.0; }, sm);
  if (threadIdx.x == 0) {
    st->enter = e;
    st->cur = 0;
    st->do_sweep = 0;
    st->group_base = 0;
  }
}

// F layout: F[row * KM + u] (the KM pending factors of one row are contiguous => 128-bit loads).
// All loads of pending data are issued unconditionally (index clamped for u >= s) so that they form one
// batch per pass instead of one memory round trip per pending pivot.
template <int KM>
__global__ void __launch_bounds__(256) k_blk_select(BlkView b, int group_pos) {
  __shared__ MinIdx sm[32];
  __shared__ int sh_last;
  __shared__ double sh_piv;
  const TabView& v = b.v;
  TabState* st = v.st;
  const int tid = threadIdx.x;
  const int R = v.R, C = v.C, ld = v.ld;
  const double* T = v.T;
  const int j = blockIdx.x * blockDim.x + tid;
  pdl_wait_then_release();
  const int status = st->status;
  const long long npiv = st->npiv;
  const long long maxp = st->max_piv;
  const int cur = st->cur;
  const int e = st->enter;
  if (status != LPR_RUNNING) {
    if (group_pos == 0 && blockIdx.x == 0 && tid == 0) st->group_base = npiv;  // nothing pending any more
    return;
  }
  const int s = group_pos;  // pending pivots of this group (every earlier select of the group succeeded)
  const double* rhs = cur ? b.rhs1 : b.rhs0;
  double* rhs_next = cur ? b.rhs0 : b.rhs1;
  int term = LPR_RUNNING, p = -1;
  double f0 = 0.0, piv = 0.0;
  int pu[KM];
#pragma unroll
  for (int u = 0; u < KM; u++) pu[u] = b.pidx[u];  // entries >= s are stale, never used
  if (e < 0) {
    term = LPR_OPTIMAL;
  } else {
    double pe[KM];
#pragma unroll
    for (int u = 0; u < KM; u++) pe[u] = b.PR[(size_t)u * ld + e];
    // entering column of the CURRENT tableau = stale column + pending updates (same roundings as the
    // sweeps would have applied), then FindLeavingVariable :169-191
    MinIdx best = minidx_identity();
    double best_a = 0.0;
    constexpr int UR = 4;
    for (int base = 0; base < R; base += 256 * UR) {
      double col[UR], rv[UR];
      double2 fq[UR][KM / 2];
#pragma unroll
      for (int q = 0; q < UR; q++) {
        const int i = min(base + q * 256 + tid, R - 1);
        col[q] = TAT(T, ld, i, e);
        rv[q] = rhs[i];
        const double2* fr = reinterpret_cast<const double2*>(b.F + (size_t)i * KM);
#pragma unroll
        for (int h2 = 0; h2 < KM / 2; h2++) fq[q][h2] = fr[h2];
      }
#pragma unroll
      for (int q = 0; q < UR; q++) {
        const int i = base + q * 256 + tid;
#pragma unroll
        for (int u = 0; u < KM; u++) {
          const double fu = (u & 1) ? fq[q][u >> 1].y : fq[q][u >> 1].x;
          const double upd = (i == pu[u]) ? pe[u] : __dsub_rn(col[q], __dmul_rn(fu, pe[u]));
          col[q] = (u < s) ? upd : col[q];
        }
        if (i < R) {
          if (i == 0) f0 = col[q];
          if (blockIdx.x == 0) b.F[(size_t)i * KM + s] = col[q];  // factor column of this pivot
          if (i >= 1 && col[q] > 1e-9) {
            const double val = ddiv(rv[q], col[q]);
            if (val >= 0.0 && val < DBL_MAX) {
              MinIdx nb = minidx_combine(best, MinIdx{val, i - 1});
              if (nb.i != best.i) best_a = col[q];
              best = nb;
            }
          }
        }
      }
    }
    const MinIdx mine = best;
    best = block_minidx(best, sm);
    const int k = (best.i == INT_MAX) ? -1 : best.i;
    if (tid == 0) sh_piv = f0;  // thread 0 owns row 0 (base 0, q 0)
    __syncthreads();
    f0 = sh_piv;
    __syncthreads();
    if (k >= 0 && mine.i == k) sh_piv = best_a;
    __syncthreads();
    if (k < 0) term = LPR_UNBOUNDED;
    else if (maxp >= 0 && npiv >= maxp) term = LPR_ITER_LIMIT;
    p = k + 1;
    if (k >= 0) piv = sh_piv;
  }
  MinIdx m = minidx_identity();
  if (term == LPR_RUNNING) {
    // pivot row slice of the current tableau, normalised (:197-199); objective row mirror (:206-208)
    if (j < ld) {
      double pr = 0.0, z = 0.0;
      if (j < C) {
        double x = TAT(T, ld, p, j);
        double pru[KM];
        double2 fp[KM / 2];
        const double2* fr = reinterpret_cast<const double2*>(b.F + (size_t)p * KM);
#pragma unroll
        for (int u = 0; u < KM; u++) pru[u] = b.PR[(size_t)u * ld + j];
#pragma unroll
        for (int h2 = 0; h2 < KM / 2; h2++) fp[h2] = fr[h2];
        const double r0j = b.row0[j];
#pragma unroll
        for (int u = 0; u < KM; u++) {
          const double fu = (u & 1) ? fp[u >> 1].y : fp[u >> 1].x;
          const double upd = (p == pu[u]) ? pru[u] : __dsub_rn(x, __dmul_rn(fu, pru[u]));
          x = (u < s) ? upd : x;
        }
        pr = ddiv(x, piv);
        z = __dsub_rn(r0j, __dmul_rn(f0, pr));
        if (j < C - 1 && z < 0.0) m = MinIdx{z, j};
      }
      b.PR[(size_t)s * ld + j] = pr;
      b.row0[j] = z;
    }
    m = block_minidx(m, sm);
    if (blockIdx.x == 0) {
      // RHS column mirror: new = (i == p) ? prow[C-1] : rhs[i] - f_i * prow[C-1], prow[C-1] = rhs[p]/piv
      const double prc = ddiv(rhs[p], piv);
      for (int i = tid; i < R; i += blockDim.x) {
        const double fi = b.F[(size_t)i * KM + s];  // written by this same thread above
        rhs_next[i] = (i == p) ? prc : __dsub_rn(rhs[i], __dmul_rn(fi, prc));
      }
    }
  }
  if (tid == 0) {
    b.cand[blockIdx.x] = m;
    __threadfence();
    unsigned t = atomicAdd(b.ticket, 1u);
    sh_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!sh_last) return;
  __threadfence();
  if (term == LPR_RUNNING) {
    MinIdx r = minidx_identity();
    for (int q = tid; q < (int)gridDim.x; q += blockDim.x) r = minidx_combine(r, b.cand[q]);
    r = block_minidx(r, sm);
    m = r;
  }
  if (tid == 0) {
    *b.ticket = 0;
    if (group_pos == 0) st->group_base = npiv;
    if (term != LPR_RUNNING) {
      st->status = term;
    } else {
      b.pidx[s] = p;
      st->leave = p;
      st->enter = (m.i == INT_MAX) ? -1 : m.i;  // entering column of the next pivot
      st->npiv = npiv + 1;
      st->cur = cur ^ 1;
      st->pivot = piv;
      if (v.log && npiv < v.log_cap) {
        v.log[2 * npiv] = p;
        v.log[2 * npiv + 1] = e;
      }
      if (v.basis) v.basis[p - 1] = e;  // :142
    }
  }
}

// ---- select v2: rows of the entering column are DEALT to the CTAs (no redundancy), the per-CTA ratio
// candidates are combined after a grid-wide barrier.  All CTAs of this small grid (<= #SMs) are co-resident,
// so a spin barrier on a global counter is safe; the host falls back to the redundant kernel otherwise.
struct RatioCand {
  double val;
  double a;
  int idx;
  int pad;
};

template <int KM>
__global__ void __launch_bounds__(256) k_blk_select2(BlkView b, int group_pos, RatioCand* rc, unsigned* bar,
                                                     double* f0buf) {
  __shared__ MinIdx sm[32];
  __shared__ int sh_last;
  __shared__ double sh_piv;
  const TabView& v = b.v;
  TabState* st = v.st;
  const int tid = threadIdx.x;
  const int R = v.R, C = v.C, ld = v.ld;
  const int G = gridDim.x;
  const double* T = v.T;
  const int j = blockIdx.x * blockDim.x + tid;
  pdl_wait_then_release();
  const int status = st->status;
  const long long npiv = st->npiv;
  const long long maxp = st->max_piv;
  const int cur = st->cur;
  const int e = st->enter;
  if (status != LPR_RUNNING) {
    if (group_pos == 0 && blockIdx.x == 0 && tid == 0) st->group_base = npiv;
    return;
  }
  const int s = group_pos;
  const double* rhs = cur ? b.rhs1 : b.rhs0;
  double* rhs_next = cur ? b.rhs0 : b.rhs1;
  int pu[KM];
#pragma unroll
  for (int u = 0; u < KM; u++) pu[u] = b.pidx[u];
  int term = (e < 0) ? LPR_OPTIMAL : LPR_RUNNING;
  // ---- phase A: this CTA's rows of the entering column ------------------------------------------------
  const int rpc = (R + G - 1) / G;
  const int row_lo = blockIdx.x * rpc, row_hi = min(R, row_lo + rpc);
  constexpr int UR = 2;  // rows per thread kept in registers (rpc <= 512); larger rpc loops
  double colk[UR];
  MinIdx best = minidx_identity();
  double best_a = 0.0;
  if (term == LPR_RUNNING) {
    double pe[KM];
#pragma unroll
    for (int u = 0; u < KM; u++) pe[u] = b.PR[(size_t)u * ld + e];
    for (int base = row_lo; base < row_hi; base += 256 * UR) {
      double rv[UR];
      double2 fq[UR][KM / 2];
#pragma unroll
      for (int q = 0; q < UR; q++) {
        const int i = min(base + q * 256 + tid, R - 1);
        colk[q] = TAT(T, ld, i, e);
        rv[q] = rhs[i];
        const double2* fr = reinterpret_cast<const double2*>(b.F + (size_t)i * KM);
#pragma unroll
        for (int h2 = 0; h2 < KM / 2; h2++) fq[q][h2] = fr[h2];
      }
#pragma unroll
      for (int q = 0; q < UR; q++) {
        const int i = base + q * 256 + tid;
#pragma unroll
        for (int u = 0; u < KM; u++) {
          const double fu = (u & 1) ? fq[q][u >> 1].y : fq[q][u >> 1].x;
          const double upd = (i == pu[u]) ? pe[u] : __dsub_rn(colk[q], __dmul_rn(fu, pe[u]));
          colk[q] = (u < s) ? upd : colk[q];
        }
        if (i < row_hi) {
          if (i == 0) *f0buf = colk[q];
          b.F[(size_t)i * KM + s] = colk[q];  // factor column of this pivot
          if (i >= 1 && colk[q] > 1e-9) {
            const double val = ddiv(rv[q], colk[q]);
            if (val >= 0.0 && val < DBL_MAX) {
              MinIdx nb = minidx_combine(best, MinIdx{val, i - 1});
              if (nb.i != best.i) best_a = colk[q];
              best = nb;
            }
          }
        }
      }
    }
    const MinIdx mine = best;
    best = block_minidx(best, sm);
    if (best.i != INT_MAX && mine.i == best.i) sh_piv = best_a;
    __syncthreads();
    if (tid == 0) {
      RatioCand c;
      c.val = best.v;
      c.idx = best.i;
      c.a = (best.i != INT_MAX) ? sh_piv : 0.0;
      c.pad = 0;
      rc[blockIdx.x] = c;
    }
  }
  // ---- grid barrier (every CTA takes part, also when the run is about to terminate) ----------------------
  if (tid == 0) {
    __threadfence();
    atomicAdd(bar, 1u);
    while (*((volatile unsigned*)bar) < (unsigned)G) {}
    __threadfence();
  }
  __syncthreads();
  int p = -1;
  double piv = 0.0, f0 = 0.0;
  if (term == LPR_RUNNING) {
    MinIdx r = minidx_identity();
    double ra = 0.0;
    for (int q = tid; q < G; q += blockDim.x) {
      const RatioCand c = rc[q];
      MinIdx nb = minidx_combine(r, MinIdx{c.val, c.idx});
      if (nb.i != r.i) ra = c.a;
      r = nb;
    }
    const MinIdx mine = r;
    r = block_minidx(r, sm);
    if (r.i != INT_MAX && mine.i == r.i) sh_piv = ra;
    __syncthreads();
    const int k = (r.i == INT_MAX) ? -1 : r.i;
    if (k < 0) term = LPR_UNBOUNDED;
    else if (maxp >= 0 && npiv >= maxp) term = LPR_ITER_LIMIT;
    p = k + 1;
    if (k >= 0) piv = sh_piv;
    f0 = *((volatile double*)f0buf);
  }
  // ---- phase B: pivot row slice, objective row / RHS mirrors, next entering candidate -------------------
  MinIdx m = minidx_identity();
  if (term == LPR_RUNNING) {
    const double rhsp = rhs[p];
    if (j < ld) {
      double pr = 0.0, z = 0.0;
      if (j < C) {
        double x = TAT(T, ld, p, j);
        double pru[KM];
        double2 fp[KM / 2];
        const double2* fr = reinterpret_cast<const double2*>(b.F + (size_t)p * KM);
#pragma unroll
        for (int u = 0; u < KM; u++) pru[u] = b.PR[(size_t)u * ld + j];
#pragma unroll
        for (int h2 = 0; h2 < KM / 2; h2++) fp[h2] = fr[h2];
        const double r0j = b.row0[j];
#pragma unroll
        for (int u = 0; u < KM; u++) {
          const double fu = (u & 1) ? fp[u >> 1].y : fp[u >> 1].x;
          const double upd = (p == pu[u]) ? pru[u] : __dsub_rn(x, __dmul_rn(fu, pru[u]));
          x = (u < s) ? upd : x;
        }
        pr = ddiv(x, piv);
        z = __dsub_rn(r0j, __dmul_rn(f0, pr));
        if (j < C - 1 && z < 0.0) m = MinIdx{z, j};
      }
      b.PR[(size_t)s * ld + j] = pr;
      b.row0[j] = z;
    }
    m = block_minidx(m, sm);
    // RHS mirror for this CTA's rows: new = (i == p) ? prow[C-1] : rhs[i] - f_i * prow[C-1]
    const double prc = ddiv(rhsp, piv);
    if (rpc <= 256 * UR) {
#pragma unroll
      for (int q = 0; q < UR; q++) {
        const int i = row_lo + q * 256 + tid;
        if (i < row_hi) rhs_next[i] = (i == p) ? prc : __dsub_rn(rhs[i], __dmul_rn(colk[q], prc));
      }
    } else {
      for (int i = row_lo + tid; i < row_hi; i += blockDim.x)
        rhs_next[i] = (i == p) ? prc : __dsub_rn(rhs[i], __dmul_rn(b.F[(size_t)i * KM + s], prc));
    }
  }
  if (tid == 0) {
    b.cand[blockIdx.x] = m;
    __threadfence();
    unsigned t = atomicAdd(b.ticket, 1u);
    sh_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!sh_last) return;
  __threadfence();
  if (term == LPR_RUNNING) {
    MinIdx r = minidx_identity();
    for (int q = tid; q < G; q += blockDim.x) r = minidx_combine(r, b.cand[q]);
    r = block_minidx(r, sm);
    m = r;
  }
  if (tid == 0) {
    *b.ticket = 0;
    *bar = 0;  // every CTA is past the barrier (it reached the ticket)
    if (group_pos == 0) st->group_base = npiv;
    if (term != LPR_RUNNING) {
      st->status = term;
    } else {
      b.pidx[s] = p;
      st->leave = p;
      st->enter = (m.i == INT_MAX) ? -1 : m.i;
      st->npiv = npiv + 1;
      st->cur = cur ^ 1;
      st->pivot = piv;
      if (v.log && npiv < v.log_cap) {
        v.log[2 * npiv] = p;
        v.log[2 * npiv + 1] = e;
      }
      if (v.basis) v.basis[p - 1] = e;  // :142
    }
  }
}

// applies the s = npiv - group_base pending pivots to the whole tableau, once.  Thread owns one 16-byte
// column chunk (its pending pivot-row values live in registers) and walks down rows; the KM factors of
// the CTA's rows are staged once in shared memory and broadcast from there.
// rows per work item RB (shared-memory stage: RB x KM doubles): small items keep the tail wave short

// ---- select v3: ONE launch selects all K pivots of a group.  A thread-block cluster (8 portable / 16
// non-portable CTAs x 512 threads) owns the whole selection: thread g owns rows g, g+T, ... of the entering
// column and columns g, g+T, ... of the pivot row; the two argmin reductions of a pivot go through
// distributed shared memory and the hardware cluster barrier instead of global-memory tickets and kernel
// boundaries (~4 us per pivot instead of ~12).
namespace cg = cooperative_groups;

struct ClusterSlot {
  double val;   // ratio / objective value
  double aux;   // pivot element candidate (phase A), f0 (CTA that owns row 0)
  int idx;
  int has_f0;
};

// one-barrier block argmin: every warp publishes its candidate, every thread re-reduces the <= 32 entries
__device__ __forceinline__ MinIdx block_minidx_1sync(MinIdx x, MinIdx* smem /* 32 entries, private to this call site */) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  x = warp_minidx(x);
  if (lane == 0) smem[w] = x;
  __syncthreads();
  MinIdx r = (lane < nw) ? smem[lane] : minidx_identity();
  return warp_minidx(r);
}

// cluster-wide argmin without a block-level stage: every warp leader stores its candidate into slot
// [crank*nw + w] of EVERY CTA of the cluster (distributed shared memory), one cluster barrier, then each CTA
// reduces the ncta*nw candidates locally.  `aux` rides along (pivot element of the winning row).
struct WarpCand {
  double val;
  double aux;
  int idx;
  int pad;
};
constexpr int kMaxClusterWarps = 16 * 32;  // 16 CTAs x up to 32 warps

__device__ __forceinline__ void cluster_publish(cg::cluster_group& cluster, WarpCand* slots, int ncta, int crank,
                                                MinIdx x, double aux) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const MinIdx mine = x;
  x = warp_minidx(x);
  const unsigned owner = __ballot_sync(0xffffffffu, x.i != INT_MAX && mine.i == x.i);
  const double a = __shfl_sync(0xffffffffu, aux, owner ? (__ffs(owner) - 1) : 0);
  if (lane < ncta) {  // lane r writes this warp's candidate into CTA r
    WarpCand* dst = cluster.map_shared_rank(slots, lane);
    WarpCand c;
    c.val = x.v;
    c.aux = a;
    c.idx = x.i;
    c.pad = 0;
    dst[crank * nw + w] = c;
  }
}
// after cluster.sync(): reduce the local copy (all threads get the result through shared memory)
__device__ __forceinline__ void cluster_collect(const WarpCand* slots, int total, MinIdx* out, double* aux_out,
                                                int* s_i, double* s_d) {
  if (threadIdx.x < 32) {
    MinIdx r = minidx_identity();
    double ra = 0.0;
    for (int k = threadIdx.x; k < total; k += 32) {
      const WarpCand c = slots[k];
      MinIdx nb = minidx_combine(r, MinIdx{c.val, c.idx});
      if (nb.i != r.i) ra = c.aux;
      r = nb;
    }
    const MinIdx mine = r;
    r = warp_minidx(r);
    const unsigned owner = __ballot_sync(0xffffffffu, r.i != INT_MAX && mine.i == r.i);
    const double a = __shfl_sync(0xffffffffu, ra, owner ? (__ffs(owner) - 1) : 0);
    if (threadIdx.x == 0) {
      s_i[0] = r.i;
      s_d[0] = r.v;
      s_d[1] = a;
    }
  }
  __syncthreads();
  out->i = s_i[0];
  out->v = s_d[0];
  *aux_out = s_d[1];
}

template <int KM, int NT>
__global__ void __launch_bounds__(NT) k_blk_select_cluster(BlkView b, int K) {
  cg::cluster_group cluster = cg::this_cluster();
  const int ncta = (int)cluster.num_blocks();
  const int crank = (int)cluster.block_rank();
  const int tid = threadIdx.x;
  const int nthr = ncta * NT;
  const int gid = crank * NT + tid;
  const int total_warps = ncta * (NT / 32);
  __shared__ WarpCand slotA[2][kMaxClusterWarps / 2], slotB[2][kMaxClusterWarps / 2];  // NT <= 512: 16 warps x 16 CTAs
  __shared__ double s_pe[KM], s_fp[KM];
  __shared__ int s_pu[KM];
  __shared__ double s_f0[2];
  __shared__ double s_d[2];
  __shared__ int s_i[1];
  const TabView& v = b.v;
  TabState* st = v.st;
  const int R = v.R, C = v.C, ld = v.ld;
  const double* T = v.T;
  pdl_wait_then_release();
  const int status = st->status;
  long long npiv = st->npiv;
  const long long npiv0 = npiv;
  const long long maxp = st->max_piv;
  int cur = st->cur;
  int e = st->enter;
  if (status != LPR_RUNNING) {
    if (gid == 0) st->group_base = npiv;
    return;
  }
  if (tid < KM) s_pu[tid] = -1;
  __syncthreads();
  int term = LPR_RUNNING;
  int s = 0;
  auto stamp = [&](int q, int k) {
    if (b.dbg && gid == 0) {
      long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      b.dbg[q * 8 + k] = t;
    }
  };
  const bool own_row = gid < R;
  const int j0 = gid, j1 = gid + nthr;          // the (up to) two columns of this thread
  const bool own_c0 = j0 < ld, own_c1 = j1 < ld;
  double fq0[KM];                                 // this thread's row of pending factors, kept in registers
#pragma unroll
  for (int u = 0; u < KM; u++) fq0[u] = 0.0;
  for (int q = 0; q < K; q++, s++) {
    if (e < 0) { term = LPR_OPTIMAL; break; }
    stamp(q, 0);
    const double* rhs = cur ? b.rhs1 : b.rhs0;
    double* rhs_next = cur ? b.rhs0 : b.rhs1;
    const int buf = q & 1;
    // ---- phase A: entering column of the current tableau (stale column + pending updates), ratio test ----
    if (tid < KM) s_pe[tid] = (tid < s) ? __ldcg(b.PR + (size_t)tid * ld + e) : 0.0;  // issued first: L2 trip
    double col0 = 0.0, rv0 = 0.0;
    if (own_row) {
      col0 = TAT(T, ld, gid, e);  // DRAM gather of the stale column
      rv0 = rhs[gid];
    }
    __syncthreads();
    stamp(q, 1);
    MinIdx best = minidx_identity();
    double best_a = 0.0;
    auto ratio_cand = [&](int i, double col, double rv) {
      if (i >= 1 && col > 1e-9) {
        const double val = ddiv(rv, col);
        if (val >= 0.0 && val < DBL_MAX) {
          MinIdx nb = minidx_combine(best, MinIdx{val, i - 1});
          if (nb.i != best.i) best_a = col;
          best = nb;
        }
      }
    };
    if (own_row) {
#pragma unroll
      for (int u = 0; u < KM; u++) {
        const double upd = (gid == s_pu[u]) ? s_pe[u] : __dsub_rn(col0, __dmul_rn(fq0[u], s_pe[u]));
        col0 = (u < s) ? upd : col0;
      }
#pragma unroll
      for (int u = 0; u < KM; u++)
        if (u == s) fq0[u] = col0;          // remember this pivot's factor for the next pivots of the group
      b.F[(size_t)gid * KM + s] = col0;     // and publish it for the sweep
      if (gid == 0) s_f0[0] = col0;         // f0 = T[0, e] (thread 0 of CTA 0)
      ratio_cand(gid, col0, rv0);
    }
    for (int i = gid + nthr; i < R; i += nthr) {  // only when the cluster has fewer threads than rows
      double col = TAT(T, ld, i, e);
      for (int u = 0; u < s; u++) {
        const double fu = b.F[(size_t)i * KM + u];
        col = (i == s_pu[u]) ? s_pe[u] : __dsub_rn(col, __dmul_rn(fu, s_pe[u]));
      }
      b.F[(size_t)i * KM + s] = col;
      ratio_cand(i, col, rhs[i]);
    }
    stamp(q, 2);
    cluster_publish(cluster, slotA[buf], ncta, crank, best, best_a);
    if (crank == 0 && tid == 0) {  // f0 travels with the barrier too: write it into every CTA
      for (int r = 0; r < ncta; r++) *cluster.map_shared_rank(&s_f0[1], r) = s_f0[0];
    }
    cluster.sync();  // release/acquire at cluster scope: also orders the global writes above
    stamp(q, 3);
    MinIdx ra;
    double piv = 0.0;
    cluster_collect(slotA[buf], total_warps, &ra, &piv, s_i, s_d);
    if (ra.i == INT_MAX) { term = LPR_UNBOUNDED; break; }
    if (maxp >= 0 && npiv >= maxp) { term = LPR_ITER_LIMIT; break; }
    const int p = ra.i + 1;
    const double f0 = s_f0[1];
    stamp(q, 4);
    // ---- phase B: pivot row (stale row + pending updates), objective row / RHS mirrors, next entering ----
    if (tid < KM) s_fp[tid] = (tid < s) ? __ldcg(b.F + (size_t)p * KM + tid) : 0.0;  // issued first: L2 trip
    const double rhsp = __ldcg(rhs + p);
    double x0 = 0.0, x1 = 0.0, r00 = 0.0, r01 = 0.0;
    double pru0[KM], pru1[KM];
    if (own_c0 && j0 < C) {
      x0 = TAT(T, ld, p, j0);
      r00 = b.row0[j0];
#pragma unroll
      for (int u = 0; u < KM; u++) pru0[u] = b.PR[(size_t)u * ld + j0];
    }
    if (own_c1 && j1 < C) {
      x1 = TAT(T, ld, p, j1);
      r01 = b.row0[j1];
#pragma unroll
      for (int u = 0; u < KM; u++) pru1[u] = b.PR[(size_t)u * ld + j1];
    }
    __syncthreads();
    stamp(q, 5);
    MinIdx m = minidx_identity();
    auto finish_col = [&](int j, double x, double r0j, const double* pru) {
      double pr = 0.0, z = 0.0;
      if (j < C) {
#pragma unroll
        for (int u = 0; u < KM; u++) {
          const double upd = (p == s_pu[u]) ? pru[u] : __dsub_rn(x, __dmul_rn(s_fp[u], pru[u]));
          x = (u < s) ? upd : x;
        }
        pr = ddiv(x, piv);
        z = __dsub_rn(r0j, __dmul_rn(f0, pr));
        if (j < C - 1 && z < 0.0) m = minidx_combine(m, MinIdx{z, j});
      }
      b.PR[(size_t)s * ld + j] = pr;
      b.row0[j] = z;
    };
    if (own_c0) finish_col(j0, x0, r00, pru0);
    if (own_c1) finish_col(j1, x1, r01, pru1);
    for (int j = gid + 2 * nthr; j < ld; j += nthr) {  // only for very wide tableaux
      double pru[KM];
#pragma unroll
      for (int u = 0; u < KM; u++) pru[u] = (j < C) ? b.PR[(size_t)u * ld + j] : 0.0;
      finish_col(j, (j < C) ? TAT(T, ld, p, j) : 0.0, (j < C) ? b.row0[j] : 0.0, pru);
    }
    {
      const double prc = ddiv(rhsp, piv);
      if (own_row) rhs_next[gid] = (gid == p) ? prc : __dsub_rn(rv0, __dmul_rn(col0, prc));
      for (int i = gid + nthr; i < R; i += nthr)
        rhs_next[i] = (i == p) ? prc : __dsub_rn(rhs[i], __dmul_rn(b.F[(size_t)i * KM + s], prc));
    }
    stamp(q, 6);
    cluster_publish(cluster, slotB[buf], ncta, crank, m, 0.0);
    if (tid == 0) s_pu[s] = p;
    if (gid == 0) {
      b.pidx[s] = p;
      if (v.log && npiv < v.log_cap) {
        v.log[2 * npiv] = p;
        v.log[2 * npiv + 1] = e;
      }
      if (v.basis) v.basis[p - 1] = e;  // :142
      st->pivot = piv;
      st->leave = p;
    }
    cluster.sync();
    stamp(q, 7);
    MinIdx rb;
    double dummy;
    cluster_collect(slotB[buf], total_warps, &rb, &dummy, s_i, s_d);
    e = (rb.i == INT_MAX) ? -1 : rb.i;
    npiv++;
    cur ^= 1;
  }
  // peers may still be writing into / reading from this CTA's shared memory: leave together
  cluster.sync();
  if (gid == 0) {
    st->group_base = npiv0;
    st->npiv = npiv;
    st->cur = cur;
    st->enter = e;
    if (term != LPR_RUNNING) st->status = term;
  }
}

template <int UNROLL, int KM, int kBlkRowsMax>
__global__ void __launch_bounds__(kSweepThreads) k_blk_sweep(BlkView b) {
  __shared__ __align__(16) double sF[kBlkRowsMax * KM];
  __shared__ unsigned sPiv[(kBlkRowsMax + 31) / 32];
  pdl_wait_then_release();
  const TabView& v = b.v;
  const TabState* st = v.st;
  const int s = (int)(st->npiv - st->group_base);
  if (s <= 0) return;
  const int R = v.R, ld = v.ld;
  const int ldv = ld >> 1;
  double2* T2 = reinterpret_cast<double2*>(v.T);
  int pu[KM];
#pragma unroll
  for (int u = 0; u < KM; u++) pu[u] = (u < s) ? b.pidx[u] : -1;
  const bool full = (s == KM);

  // work items: (column group, block of kBlkRowsMax rows), dealt round-robin to the CTAs
  const int nfull = ldv / kSweepThreads;
  const int nrb = (R + kBlkRowsMax - 1) / kBlkRowsMax;
  const int items = nfull * nrb;
  for (int it = blockIdx.x; it < items; it += gridDim.x) {
    const int cg = it / nrb, rb = it - cg * nrb;
    const int r0 = rb * kBlkRowsMax, r1 = min(R, r0 + kBlkRowsMax);
    __syncthreads();
    if (threadIdx.x < (kBlkRowsMax + 31) / 32) sPiv[threadIdx.x] = full ? 0u : 0xffffffffu;  // partial group: slow path
    __syncthreads();
    if (full && threadIdx.x < KM) {
      const int q = pu[threadIdx.x] - r0;  // pu[] is a register array: pick this thread's entry without indexing
      (void)q;
    }
    if (full && threadIdx.x == 0) {
#pragma unroll
      for (int u = 0; u < KM; u++) {
        const int q = pu[u] - r0;
        if (q >= 0 && q < kBlkRowsMax) sPiv[q >> 5] |= 1u << (q & 31);
      }
    }
    for (int t = threadIdx.x; t < (r1 - r0) * KM / 2; t += blockDim.x)
      reinterpret_cast<double2*>(sF)[t] = __ldg(reinterpret_cast<const double2*>(b.F + (size_t)r0 * KM) + t);
    const int chunk = cg * kSweepThreads + threadIdx.x;
    double2 pr[KM];
#pragma unroll
    for (int u = 0; u < KM; u++) pr[u] = reinterpret_cast<const double2*>(b.PR + (size_t)u * ld)[chunk];
    __syncthreads();
    double2* d = T2 + chunk;
    // software pipeline: the loads of the next UNROLL rows are in flight while the current rows go through
    // their KM dependent multiply/subtract pairs (the DP work of a full group is ~2/3 of the memory time)
    double2 xc[UNROLL], xn[UNROLL];
#pragma unroll
    for (int k = 0; k < UNROLL; k++)
      if (r0 + k < r1) xc[k] = ld_stream(d + (size_t)(r0 + k) * ldv);
    for (int r = r0; r < r1; r += UNROLL) {
#pragma unroll
      for (int k = 0; k < UNROLL; k++)
        if (r + UNROLL + k < r1) xn[k] = ld_stream(d + (size_t)(r + UNROLL + k) * ldv);
#pragma unroll
      for (int k = 0; k < UNROLL; k++) {
        const int row = r + k;
        if (row < r1) {
          const double2* fr = reinterpret_cast<const double2*>(sF + (size_t)(row - r0) * KM);
          const bool slow = (sPiv[(row - r0) >> 5] >> ((row - r0) & 31)) & 1u;  // CTA-uniform
          d[(size_t)row * ldv] = slow ? blk_apply_gen<KM>(xc[k], pr, fr, row, s, pu) : blk_apply_fast<KM>(xc[k], pr, fr);
        }
      }
#pragma unroll
      for (int k = 0; k < UNROLL; k++) xc[k] = xn[k];
    }
  }
  // ragged remainder: the last (ldv % 256) chunks of every row, one thread per row
  const int rem0 = nfull * kSweepThreads;
  if (rem0 < ldv) {
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    const int nth = gridDim.x * blockDim.x;
    for (int row = gid; row < R; row += nth) {
      for (int c = rem0; c < ldv; c++) {
        double2 pr[KM];
#pragma unroll
        for (int u = 0; u < KM; u++) pr[u] = reinterpret_cast<const double2*>(b.PR + (size_t)u * ld)[c];
        T2[(size_t)row * ldv + c] = blk_apply_gen<KM>(T2[(size_t)row * ldv + c], pr,
                                                      reinterpret_cast<const double2*>(b.F + (size_t)row * KM), row, s, pu);
      }
    }
  }
}

__global__ void k_state_reset_blk(TabState* st, long long max_piv);

int tab_blocked_alloc(lpr_tab* h, int K) {
  K = (K <= 8) ? 8 : 16;  // storage is sized for the kernel instantiation (KM)
  if (h->blk_k == K) return LPR_OK;
  cudaFree(h->blk_pr); cudaFree(h->blk_f); cudaFree(h->blk_row0); cudaFree(h->blk_rhs[0]); cudaFree(h->blk_rhs[1]);
  cudaFree(h->blk_p);
  h->blk_pr = h->blk_f = h->blk_row0 = h->blk_rhs[0] = h->blk_rhs[1] = nullptr;
  h->blk_p = nullptr;
  h->blk_k = 0;
  LPR_CUDA(cudaMalloc(&h->blk_pr, sizeof(double) * (size_t)K * h->ld));
  LPR_CUDA(cudaMalloc(&h->blk_f, sizeof(double) * (size_t)K * h->Rcap));
  LPR_CUDA(cudaMalloc(&h->blk_row0, sizeof(double) * h->ld));
  LPR_CUDA(cudaMalloc(&h->blk_rhs[0], sizeof(double) * h->Rcap));
  LPR_CUDA(cudaMalloc(&h->blk_rhs[1], sizeof(double) * h->Rcap));
  LPR_CUDA(cudaMalloc(&h->blk_p, sizeof(int) * K));
  LPR_CUDA(cudaMemset(h->blk_pr, 0, sizeof(double) * (size_t)K * h->ld));
  LPR_CUDA(cudaMemset(h->blk_f, 0, sizeof(double) * (size_t)K * h->Rcap));
  LPR_CUDA(cudaMemset(h->blk_p, 0xff, sizeof(int) * K));
  h->blk_k = K;
  return LPR_OK;
}

template <class... KArgs, class... Args>
static cudaError_t launch_pdl_b(void (*kernel)(KArgs...), dim3 grid, dim3 block, cudaStream_t stream, Args... args) {
  static const int pdl = getenv("LPR_PDL") ? atoi(getenv("LPR_PDL")) : 1;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// Solve() with K delayed pivots per tableau sweep.  Same contract as tab_solve_internal for RULE_PRIMAL.
int tab_solve_blocked(lpr_tab* h, int K, int64_t max_pivots, int* status, int64_t* n_pivots, int* pivot_log,
                      int64_t log_cap, bool time_sweeps) {
  K = std::max(2, std::min(K, KMAX));
  int rc = tab_blocked_alloc(h, K);
  if (rc) return rc;
  if (pivot_log && log_cap > 0) {
    long long want = log_cap;
    if (max_pivots >= 0) want = std::min<long long>(want, max_pivots + 1);
    want = std::min<long long>(want, 1LL << 24);
    if ((rc = tab_ensure_log(h, want))) return rc;
  }
  BlkView b;
  b.v = h->view();
  if (!(pivot_log && log_cap > 0)) { b.v.log = nullptr; b.v.log_cap = 0; }
  b.PR = h->blk_pr; b.F = h->blk_f; b.row0 = h->blk_row0; b.rhs0 = h->blk_rhs[0]; b.rhs1 = h->blk_rhs[1];
  b.pidx = h->blk_p; b.Rcap = h->Rcap; b.cand = h->selcand; b.ticket = h->ticket;
  b.dbg = nullptr;
  long long* d_dbg = nullptr;
  if (getenv("LPR_BLK_TIMING")) {
    LPR_CUDA(cudaMalloc(&d_dbg, sizeof(long long) * 8 * KMAX));
    LPR_CUDA(cudaMemset(d_dbg, 0, sizeof(long long) * 8 * KMAX));
    b.dbg = d_dbg;
  }
  // select v2 scratch (per-CTA ratio candidates, barrier counter, f0)
  static const int sel_v2 = getenv("LPR_BLK_SELECT_V2") ? atoi(getenv("LPR_BLK_SELECT_V2")) : 1;
  RatioCand* d_rc = nullptr;
  unsigned* d_bar = nullptr;
  double* d_f0 = nullptr;
  const bool use_v2 = sel_v2 && ((h->ld + 255) / 256 <= h->sms);
  if (use_v2) {
    LPR_CUDA(cudaMalloc(&d_rc, sizeof(RatioCand) * ((h->ld + 255) / 256 + 1)));
    LPR_CUDA(cudaMalloc(&d_bar, sizeof(unsigned)));
    LPR_CUDA(cudaMalloc(&d_f0, sizeof(double)));
    LPR_CUDA(cudaMemsetAsync(d_bar, 0, sizeof(unsigned), h->stream));
  }
  // select v3: one cluster launch per group
  static const int cl_env = getenv("LPR_BLK_CLUSTER") ? atoi(getenv("LPR_BLK_CLUSTER")) : 16;
  int cluster_ctas = 0;
  if (cl_env >= 2) {
    cluster_ctas = std::min(cl_env, 16);
    if (cluster_ctas > 8) {
      cudaError_t ce = cudaFuncSetAttribute(k_blk_select_cluster<8, 512>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
      if (ce == cudaSuccess) ce = cudaFuncSetAttribute(k_blk_select_cluster<16, 512>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
      if (ce == cudaSuccess) ce = cudaFuncSetAttribute(k_blk_select_cluster<8, 256>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
      if (ce == cudaSuccess) ce = cudaFuncSetAttribute(k_blk_select_cluster<16, 256>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
      if (ce != cudaSuccess) {
        cudaGetLastError();
        cluster_ctas = 8;
      }
    }
  }
  auto launch_cluster = [&](int ncta) -> cudaError_t {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(ncta);
    static const int nt = (getenv("LPR_BLK_CLUSTER_THREADS") ? atoi(getenv("LPR_BLK_CLUSTER_THREADS")) : 256) <= 256 ? 256 : 512;
    cfg.blockDim = dim3(nt);
    cfg.stream = h->stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = ncta;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    static const int pdl = getenv("LPR_PDL") ? atoi(getenv("LPR_PDL")) : 1;
    cfg.numAttrs = pdl ? 2 : 1;
    if (nt == 256)
      return (K <= 8) ? cudaLaunchKernelEx(&cfg, k_blk_select_cluster<8, 256>, b, K)
                      : cudaLaunchKernelEx(&cfg, k_blk_select_cluster<16, 256>, b, K);
    return (K <= 8) ? cudaLaunchKernelEx(&cfg, k_blk_select_cluster<8, 512>, b, K)
                    : cudaLaunchKernelEx(&cfg, k_blk_select_cluster<16, 512>, b, K);
  };
  static const int groups_per_batch = std::max(1, getenv("LPR_TAB_BATCH") ? atoi(getenv("LPR_TAB_BATCH")) / 4 : 8);
  const int gsel = (h->ld + 255) / 256;
  // sweep grid: contiguous (column group, row) unit ranges, ~4 waves of resident CTAs
  static const int rb_env = getenv("LPR_BLK_ROWS") ? atoi(getenv("LPR_BLK_ROWS")) : 64;
  const int RB = (rb_env >= 256) ? 256 : (rb_env >= 128 ? 128 : 64);
  const long long items = (long long)(h->ld / 2 / kSweepThreads) * ((h->R + RB - 1) / RB);
  long long gs = std::max<long long>(items, (h->R + kSweepThreads - 1) / kSweepThreads);
  gs = std::max<long long>(1, std::min<long long>(gs, (long long)h->sms * 64));

  std::vector<cudaEvent_t> sw_ev;
  h->last_sweep_us = 0.f;
  LPR_CUDA(cudaEventRecord(h->ev0, h->stream));
  k_state_reset_blk<<<1, 1, 0, h->stream>>>(h->st, (long long)max_pivots);
  LPR_LAUNCH_CHECK();
  k_blk_init<<<1, kSelThreads, 0, h->stream>>>(b);
  LPR_LAUNCH_CHECK();
  int slot = 0, pending = 0, ngroups = 1;
  while (true) {
    for (int g = 0; g < ngroups; g++) {
      if (cluster_ctas >= 2) {
        cudaError_t ce = launch_cluster(cluster_ctas);
        if (ce != cudaSuccess && cluster_ctas > 8) {  // 16-CTA clusters not schedulable here: retry with 8
          cudaGetLastError();
          cluster_ctas = 8;
          ce = launch_cluster(cluster_ctas);
        }
        if (ce != cudaSuccess) {
          cudaGetLastError();
          cluster_ctas = 0;  // fall back to the multi-launch select
        } else {
          count_launch();
        }
      }
      for (int q = 0; q < K && cluster_ctas < 2; q++) {
        cudaError_t le;
        if (use_v2)
          le = (K <= 8) ? launch_pdl_b(k_blk_select2<8>, gsel, 256, h->stream, b, q, d_rc, d_bar, d_f0)
                        : launch_pdl_b(k_blk_select2<16>, gsel, 256, h->stream, b, q, d_rc, d_bar, d_f0);
        else
          le = (K <= 8) ? launch_pdl_b(k_blk_select<8>, gsel, 256, h->stream, b, q)
                        : launch_pdl_b(k_blk_select<16>, gsel, 256, h->stream, b, q);
        if (le != cudaSuccess)
          return fail(LPR_E_CUDA, "blocked select launch failed: %s", cudaGetErrorString(cudaGetLastError()));
        count_launch();
      }
      const bool timed = time_sweeps && sw_ev.size() < 256;
      if (timed) {
        cudaEvent_t a, c2;
        LPR_CUDA(cudaEventCreate(&a));
        LPR_CUDA(cudaEventCreate(&c2));
        sw_ev.push_back(a);
        sw_ev.push_back(c2);
        LPR_CUDA(cudaEventRecord(a, h->stream));
      }
      cudaError_t se;
      if (K <= 8)
        se = RB == 256 ? launch_pdl_b(k_blk_sweep<4, 8, 256>, (int)gs, kSweepThreads, h->stream, b)
           : RB == 128 ? launch_pdl_b(k_blk_sweep<4, 8, 128>, (int)gs, kSweepThreads, h->stream, b)
                       : launch_pdl_b(k_blk_sweep<4, 8, 64>, (int)gs, kSweepThreads, h->stream, b);
      else
        se = RB == 256 ? launch_pdl_b(k_blk_sweep<4, 16, 256>, (int)gs, kSweepThreads, h->stream, b)
           : RB == 128 ? launch_pdl_b(k_blk_sweep<4, 16, 128>, (int)gs, kSweepThreads, h->stream, b)
                       : launch_pdl_b(k_blk_sweep<4, 16, 64>, (int)gs, kSweepThreads, h->stream, b);
      if (se != cudaSuccess)
        return fail(LPR_E_CUDA, "blocked sweep launch failed: %s", cudaGetErrorString(cudaGetLastError()));
      count_launch();
      if (timed) LPR_CUDA(cudaEventRecord(sw_ev.back(), h->stream));
    }
    LPR_CUDA(cudaMemcpyAsync(&h->st_host[slot], h->st, sizeof(TabState), cudaMemcpyDeviceToHost, h->stream));
    LPR_CUDA(cudaEventRecord(h->evb[slot], h->stream));
    pending++;
    if (pending == 2 || ngroups < groups_per_batch) {
      const int old = (pending == 2) ? (slot ^ 1) : slot;
      LPR_CUDA(cudaEventSynchronize(h->evb[old]));
      pending--;
      if (h->st_host[old].status != LPR_RUNNING) break;
    }
    slot ^= 1;
    ngroups = std::min(groups_per_batch, ngroups * 2);
  }
  LPR_CUDA(cudaEventRecord(h->ev1, h->stream));
  LPR_CUDA(cudaMemcpyAsync(&h->st_host[0], h->st, sizeof(TabState), cudaMemcpyDeviceToHost, h->stream));
  LPR_CUDA(cudaStreamSynchronize(h->stream));
  LPR_CUDA(cudaEventElapsedTime(&h->last_ms, h->ev0, h->ev1));
  cudaFree(d_rc);
  cudaFree(d_bar);
  cudaFree(d_f0);
  if (d_dbg) {  // phase timestamps of the LAST group's select (ns, relative to the first stamp)
    long long hd[8 * KMAX];
    cudaMemcpy(hd, d_dbg, sizeof hd, cudaMemcpyDeviceToHost);
    for (int q = 0; q < K && hd[q * 8]; q++) {
      fprintf(stderr, "[blk timing] pivot %2d:", q);
      for (int k = 1; k < 8; k++) fprintf(stderr, " %5lld", hd[q * 8 + k] - hd[q * 8 + k - 1]);
      if (q + 1 < K && hd[(q + 1) * 8]) fprintf(stderr, " | next %5lld", hd[(q + 1) * 8] - hd[q * 8 + 7]);
      fprintf(stderr, "  ns (stageA chainRatioA publishSyncA collectA stageB finishB publishSyncB)\n");
    }
    cudaFree(d_dbg);
  }
  const long long npiv = h->st_host[0].npiv;
  if (time_sweeps) {
    double sum = 0.0;
    int cnt = 0;
    const int real = (int)std::min<long long>((npiv + K - 1) / K, (long long)sw_ev.size() / 2);
    for (int k = std::min(2, real / 2); k < real; k++) {
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, sw_ev[2 * k], sw_ev[2 * k + 1]) == cudaSuccess) { sum += ms; cnt++; }
    }
    if (cnt) h->last_sweep_us = (float)(sum * 1e3 / cnt);
    for (auto ev : sw_ev) cudaEventDestroy(ev);
  }
  if (status) *status = h->st_host[0].status;
  if (n_pivots) *n_pivots = npiv;
  if (pivot_log && log_cap > 0 && npiv > 0) {
    long long cnt = std::min<long long>(std::min<long long>(npiv, log_cap), h->log_cap);
    LPR_CUDA(cudaMemcpy(pivot_log, h->log, sizeof(int) * 2 * (size_t)cnt, cudaMemcpyDeviceToHost));
  }
  return LPR_OK;
}

__global__ void k_state_reset_blk(TabState* st, long long max_piv) {
  st->status = LPR_RUNNING;
  st->enter = -1;
  st->leave = -1;
  st->next_enter = -1;
  st->do_sweep = 0;
  st->cur = 0;
  st->src = 0;
  st->phase = 0;
  st->have_prev = 0;
  st->dropped = 0;
  st->npiv = 0;
  st->max_piv = max_piv;
  st->group_base = 0;
  st->pivot = 0.0;
}

}  // namespace lpr
