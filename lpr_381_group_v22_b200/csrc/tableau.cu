// tableau.cu -- dense simplex tableau in HBM: selection kernels, the rank-1 pivot sweep and the
// device-side Solve() loops of PrimalSimplexSolver / PrimalSimplexSolver2 / DualSimplexSolver /
// SensitivityAnalyzer (reference file:line cited per kernel; semantics in SURVEY.md Appendix A).
//
// Layout: row-major fp64, leading dimension padded to a multiple of 16 doubles so every row
// starts on a 128-byte line; padding columns are kept at 0.  Arithmetic: IEEE binary64 with
// separate multiply and subtract roundings (__dmul_rn/__dsub_rn are never contracted) and
// IEEE division, which makes every element bit-identical to the reference's scalar loops.
#include "tableau.cuh"
#include "select.cuh"
#include "sweep.cuh"

#include <algorithm>
#include <cstdlib>
#include <mutex>
#include <vector>

namespace lpr {

std::string& last_error() {
  static thread_local std::string s;
  return s;
}
int fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  last_error() = buf;
  return code;
}
std::atomic<int64_t> g_launches{0};

int select_device(int device) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0)
    return fail(LPR_E_CUDA, "no CUDA device available (%s); liblprb200 has no CPU fallback",
                e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
  if (device < 0 || device >= n) return fail(LPR_E_BADARG, "device %d out of range [0,%d)", device, n);
  LPR_CUDA(cudaSetDevice(device));
  return LPR_OK;
}
int sm_count(int device) {
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  return sms > 0 ? sms : 148;
}

static int env_int(const char* name, int dflt) {
  const char* s = getenv(name);
  return (s && *s) ? atoi(s) : dflt;
}

// =============================================================================================
// device helpers
// =============================================================================================
// =============================================================================================
// k_select<RULE>: one CTA.  Commits the previous pivot, picks (leave row, enter col) with the
// rule's exact tolerances and tie breaks, stages the normalised pivot row (prow) and the
// pre-update factor column (col[cur]) for the sweep.
// =============================================================================================
enum : int { F_PRINT = 1, F_FUSED = 2 };

template <int RULE>
__global__ void __launch_bounds__(kSelThreads) k_select(TabView v, int flags, int* mask) {
  __shared__ MinIdx sm[32];
  __shared__ int smi[32];
  TabState* st = v.st;
  const int tid = threadIdx.x;
  const int R = v.R, C = v.C, ld = v.ld;
  double* T = v.T;
  pdl_wait_then_release();

  const int status = st->status;
  const int did = st->do_sweep;
  long long npiv = st->npiv;
  const long long maxp = st->max_piv;
  int phase = st->phase;
  __syncthreads();
  if (status != LPR_RUNNING) {
    if (tid == 0) st->do_sweep = 0;
    return;
  }
  if (did) npiv++;
  auto finish = [&](int s) {
    if (tid == 0) {
      st->status = s;
      st->do_sweep = 0;
      st->npiv = npiv;
      st->phase = phase;
    }
  };
  // PrimalSimplexSolver2.cs:75,90 / DualSimplex.cs:94,108: the counter only advances when
  // printSteps is set and is tested AFTER the pivot.
  if (did && (RULE == LPR_RULE_PRIMAL2 || RULE == LPR_RULE_DUAL)) {
    long long iter = (flags & F_PRINT) ? npiv : 0;
    if (maxp >= 0 && iter >= maxp) {
      finish(LPR_ITER_LIMIT);
      return;
    }
  }

  int e = -1, p = -1;
  int skip_basis_update = 0;
  if (RULE == LPR_RULE_PRIMAL) {
    // FindEnteringVariable PrimalSimplexSolver.cs:152-167: min T[0,j] < 0.0, lowest j
    e = block_first_min(C - 1, [&](int j, double& val) { val = T[j]; return val < 0.0; }, sm);
    if (e < 0) { finish(LPR_OPTIMAL); return; }
    // FindLeavingVariable :169-191: a > 1e-9, ratio >= 0, ratio < DBL_MAX, lowest row
    int k = block_first_min(R - 1, [&](int q, double& val) {
      double a = TAT(T, ld, q + 1, e);
      if (!(a > 1e-9)) return false;
      val = ddiv(TAT(T, ld, q + 1, C - 1), a);
      return val >= 0.0 && val < DBL_MAX;
    }, sm);
    if (k < 0) { finish(LPR_UNBOUNDED); return; }
    p = k + 1;
    if (maxp >= 0 && npiv >= maxp) { finish(LPR_ITER_LIMIT); return; }
  } else if (RULE == LPR_RULE_PRIMAL2) {
    const double EPS = 1e-10;
    // FindEnteringColumn PrimalSimplexSolver2.cs:102-117
    e = block_hyst_min(C - 1, [&](int j, double& val) { val = T[j]; return true; }, 0.0, EPS, sm, smi);
    if (e < 0) { finish(LPR_OPTIMAL); return; }
    // FindLeavingRow :120-141 (precedence quirk reduces to ratio > EPS && ratio < best - EPS)
    int k = block_hyst_min(R - 1, [&](int q, double& val) {
      double a = TAT(T, ld, q + 1, e);
      if (!(a > EPS)) return false;
      val = ddiv(TAT(T, ld, q + 1, C - 1), a);
      return val > EPS;
    }, kPosInf, EPS, sm, smi);
    if (k < 0) { finish(LPR_UNBOUNDED); return; }
    p = k + 1;
    if (fabs(TAT(T, ld, p, e)) <= EPS) { finish(LPR_PIVOT_TOO_SMALL); return; }
    skip_basis_update = 1;
  } else if (RULE == LPR_RULE_DUAL) {
    const double EPS = 1e-9;
    // DualSimplex.cs:27-37 most negative RHS among constraint rows
    int k = block_hyst_min(R - 1, [&](int q, double& val) { val = TAT(T, ld, q + 1, C - 1); return true; },
                           0.0, EPS, sm, smi);
    if (k < 0) { finish(LPR_OPTIMAL); return; }
    p = k + 1;
    // :50-70 min |obj_j / a_j| over a_j < -EPS, |obj_j| > EPS
    e = block_hyst_min(C - 1, [&](int j, double& val) {
      double a = TAT(T, ld, p, j);
      if (!(a < -EPS)) return false;
      double num = T[j];
      if (!(fabs(num) > EPS)) return false;
      val = fabs(ddiv(num, a));
      return true;
    }, kPosInf, EPS, sm, smi);
    if (e < 0) { finish(LPR_INFEASIBLE); return; }
    if (fabs(TAT(T, ld, p, e)) <= EPS) { finish(LPR_PIVOT_TOO_SMALL); return; }
    skip_basis_update = 1;
  } else if (RULE == LPR_RULE_SENS) {
    const double EPS = 1e-9;
    // phase pivots counter lives in st->pivot's slot? no: use st->have_prev as the per-phase count
    int phase_piv = st->have_prev + (did ? 1 : 0);
    if (phase == 0) {
      // DualSimplexIfNeeded SensitivityAnalyzer.cs:168-201
      int k = block_hyst_min(R - 1, [&](int q, double& val) { val = TAT(T, ld, q + 1, C - 1); return true; },
                             0.0, EPS, sm, smi);
      if (k < 0) {
        phase = 1;
        phase_piv = 0;
      } else {
        p = k + 1;
        if (maxp >= 0 && phase_piv > maxp) { finish(LPR_ITER_LIMIT); return; }
        e = block_hyst_min(C - 1, [&](int j, double& val) {
          double a = TAT(T, ld, p, j);
          if (!(a < -EPS)) return false;
          val = ddiv(T[j], -a);
          return true;
        }, kPosInf, EPS, sm, smi);
        if (e < 0) { finish(LPR_INFEASIBLE); return; }
      }
    }
    if (phase == 1) {
      // ReOptimize :121-166 with IsOptimal :85-96 (non-basic columns only)
      for (int j = tid; j < C; j += blockDim.x) mask[j] = 0;
      __syncthreads();
      for (int i = tid; i < R - 1; i += blockDim.x) {
        int b = v.basis[i];
        if (b >= 0 && b < C) mask[b] = 1;
      }
      __syncthreads();
      int notopt = 0;
      for (int j = tid; j < C - 1; j += blockDim.x)
        if (!mask[j] && T[j] < -EPS) notopt++;
      notopt = block_sum_int(notopt, smi);
      if (notopt == 0) { finish(LPR_OPTIMAL); return; }
      if (maxp >= 0 && phase_piv > maxp) { finish(LPR_ITER_LIMIT); return; }
      e = block_first_min(C - 1, [&](int j, double& val) { val = T[j]; return !mask[j] && val < 0.0; }, sm);
      if (e < 0) { finish(LPR_OPTIMAL); return; }
      int k = block_hyst_min(R - 1, [&](int q, double& val) {
        double a = TAT(T, ld, q + 1, e);
        if (!(a > EPS)) return false;
        val = ddiv(TAT(T, ld, q + 1, C - 1), a);
        return true;
      }, kPosInf, EPS, sm, smi);
      if (k < 0) { finish(LPR_UNBOUNDED); return; }
      p = k + 1;
    }
    if (fabs(TAT(T, ld, p, e)) < EPS) { finish(LPR_PIVOT_TOO_SMALL); return; }
    if (tid == 0) st->have_prev = phase_piv;
  }

  // ---- stage pivot row and factor column (Pivot: PrimalSimplexSolver.cs:193-200) -----------
  const double piv = TAT(T, ld, p, e);
  const int cur = st->cur;
  double* colb = cur ? v.col[1] : v.col[0];
  for (int j = tid; j < ld; j += blockDim.x) v.prow[j] = (j < C) ? ddiv(TAT(T, ld, p, j), piv) : 0.0;
  for (int i = tid; i < R; i += blockDim.x) colb[i] = TAT(T, ld, i, e);
  if (tid == 0) {
    st->enter = e;
    st->leave = p;
    st->next_enter = -1;
    st->do_sweep = 1;
    st->npiv = npiv;
    st->phase = phase;
    st->pivot = piv;
    if (v.log && npiv < v.log_cap) {
      v.log[2 * npiv] = p;
      v.log[2 * npiv + 1] = e;
    }
    if (!skip_basis_update && v.basis && p >= 1) v.basis[p - 1] = e;  // PrimalSimplexSolver.cs:142
  }
}

// ---------------------------------------------------------------------------------------------
// Fused primal path (RULE_PRIMAL, the bench hot path).  The sweep of pivot t also emits the
// post-update column of the NEXT entering variable and the post-update RHS column into side
// buffers, so selecting pivot t+1 costs O(R + C) contiguous reads instead of two strided
// column gathers, and the tableau is read and written exactly once per pivot.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kSelThreads) k_primal_init(TabView v) {
  __shared__ MinIdx sm[32];
  TabState* st = v.st;
  const double* T = v.T;
  const int R = v.R, C = v.C, ld = v.ld;
  int e = block_first_min(C - 1, [&](int j, double& val) { val = T[j]; return val < 0.0; }, sm);
  double* colb = v.col[0];
  for (int i = threadIdx.x; i < R; i += blockDim.x) {
    v.rhs[i] = TAT(T, ld, i, C - 1);
    if (e >= 0) colb[i] = TAT(T, ld, i, e);
  }
  if (threadIdx.x == 0) {
    st->enter = e;
    st->cur = 0;
    st->do_sweep = 0;
  }
}

// Multi-CTA select: CTA b owns columns [256 b, 256 b + 256).  Every CTA repeats the (cheap,
// L2-resident) ratio test so no grid-wide barrier is needed before the pivot row is normalised;
// per-CTA entering candidates are combined by the last CTA to finish (atomic ticket), which is
// also the only writer of the state block -- by then every other CTA has read it.
__global__ void __launch_bounds__(256) k_primal_select_fused(TabView v, MinIdx* cand, unsigned* ticket) {
  __shared__ MinIdx sm[32];
  __shared__ int sh_last;
  __shared__ double sh_piv;
  TabState* st = v.st;
  const int tid = threadIdx.x;
  const int R = v.R, C = v.C, ld = v.ld;
  const double* T = v.T;
  const int j = blockIdx.x * blockDim.x + tid;
  pdl_wait_then_release();
  // independent of the state: this CTA's slice of the objective row
  const double t0j = (j < C - 1) ? T[j] : 0.0;
  const int status = st->status;
  const int did = st->do_sweep;
  long long npiv = st->npiv;
  const long long maxp = st->max_piv;
  int cur = st->cur;
  int e = st->enter;
  const int nxt = st->next_enter;
  if (status != LPR_RUNNING) {
    if (blockIdx.x == 0 && tid == 0 && did) st->do_sweep = 0;  // idempotent: no reader needs the old value
    return;
  }
  if (did) {  // commit pivot t: the sweep has produced col[cur^1] for `nxt`
    npiv++;
    cur ^= 1;
    e = nxt;
  }
  int term = LPR_RUNNING, p = -1;
  const double* colb = cur ? v.col[1] : v.col[0];
  double f0 = 0.0, piv = 0.0;
  if (e < 0) {
    term = LPR_OPTIMAL;
  } else {
    // FindLeavingVariable PrimalSimplexSolver.cs:169-191 on the staged (contiguous) column + RHS.
    // All loads of a pass are issued before any use: one memory round trip per 4096 rows.
    f0 = colb[0];
    MinIdx best = minidx_identity();
    double best_a = 0.0;
    constexpr int UR = 16;
    for (int base = 1; base < R; base += 256 * UR) {
      double a[UR], b[UR];
#pragma unroll
      for (int q = 0; q < UR; q++) {
        const int i = base + q * 256 + tid;
        a[q] = (i < R) ? colb[i] : 0.0;
        b[q] = (i < R) ? v.rhs[i] : 0.0;
      }
#pragma unroll
      for (int q = 0; q < UR; q++) {
        const int i = base + q * 256 + tid;
        if (i < R && a[q] > 1e-9) {
          const double val = ddiv(b[q], a[q]);
          if (val >= 0.0 && val < DBL_MAX) {
            MinIdx cnd{val, i - 1};
            MinIdx nb = minidx_combine(best, cnd);
            if (nb.i != best.i) best_a = a[q];
            best = nb;
          }
        }
      }
    }
    const MinIdx mine = best;
    best = block_minidx(best, sm);
    const int k = (best.i == INT_MAX) ? -1 : best.i;
    if (k >= 0 && mine.i == k) sh_piv = best_a;  // exactly one thread owns row k
    __syncthreads();
    if (k < 0) term = LPR_UNBOUNDED;
    else if (maxp >= 0 && npiv >= maxp) term = LPR_ITER_LIMIT;
    p = k + 1;
    if (k >= 0) piv = sh_piv;
  }
  MinIdx m = minidx_identity();
  if (term == LPR_RUNNING) {
    // normalise the pivot row (:197-199) and, fused, evaluate the updated objective row
    // T[0,j] - f0*prow[j] (:206-208) to pick the next entering column (:152-167).
    if (j < ld) {
      double pr = 0.0;
      if (j < C) {
        pr = ddiv(TAT(T, ld, p, j), piv);
        if (j < C - 1) {
          double z = __dsub_rn(t0j, __dmul_rn(f0, pr));
          if (z < 0.0) m = MinIdx{z, j};
        }
      }
      v.prow[j] = pr;
    }
    m = block_minidx(m, sm);
  }
  if (tid == 0) {
    cand[blockIdx.x] = m;
    __threadfence();
    unsigned t = atomicAdd(ticket, 1u);
    sh_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!sh_last) return;
  __threadfence();
  if (term == LPR_RUNNING) {
    MinIdx r = minidx_identity();
    for (int b = tid; b < (int)gridDim.x; b += blockDim.x) r = minidx_combine(r, cand[b]);
    r = block_minidx(r, sm);
    m = r;
  }
  if (tid == 0) {
    *ticket = 0;
    st->npiv = npiv;
    st->cur = cur;
    st->enter = e;
    if (term != LPR_RUNNING) {
      st->status = term;
      st->do_sweep = 0;
    } else {
      st->leave = p;
      st->next_enter = (m.i == INT_MAX) ? -1 : m.i;
      st->do_sweep = 1;
      st->pivot = piv;
      if (v.log && npiv < v.log_cap) {
        v.log[2 * npiv] = p;
        v.log[2 * npiv + 1] = e;
      }
      if (v.basis) v.basis[p - 1] = e;  // :142
    }
  }
}

// =============================================================================================
// k_sweep: the rank-1 row-elimination update (Pivot: PrimalSimplexSolver.cs:201-210 and its
// five siblings).  Every element of the tableau is read once and written once per pivot:
//   row p      : T[p,j] = prow[j]                          (normalised pivot row)
//   row i != p : T[i,j] = T[i,j] - (f_i * prow[j])          (separate mul and sub roundings)
// Mapping (picked with tools/sweep_bench.cu on a B200: contiguous spans beat column-fixed
// threads by ~10%): the R x ld tableau is one flat array of 16-byte double2 chunks; a persistent
// grid of sms*k CTAs gets equal contiguous spans (no tail wave); a CTA walks its span in tiles of
// 256*UNROLL chunks with UNROLL independent 128-bit loads in flight per thread.  prow (<= 100 KB)
// and the factor column are read through the L1 read-only path.  `reverse` walks the span
// backwards: alternating the direction between pivots lets the next sweep start on the chunks
// the previous one left in L2.
//   SKIP: 0 none | 1 skip rows with |f| <= eps | 2 skip rows with |f| < eps
//   OOP : out-of-place (B&B pivots, BranchBoundSimplexSolver.cs:161-192) with -0.0 -> 0.0
//   EMIT: fused primal path, write next factor column / RHS side buffers
// =============================================================================================
template <int SKIP, bool OOP, bool EMIT, int UNROLL>
__global__ void __launch_bounds__(kSweepThreads) k_sweep(TabView v, double eps, int reverse) {
  pdl_wait_then_release();
  const TabState* st = v.st;
  if (!st->do_sweep) return;
  const int cur = st->cur;
  const double* f = cur ? v.col[1] : v.col[0];
  double* cn = cur ? v.col[0] : v.col[1];
  const double2* src = reinterpret_cast<const double2*>(OOP ? (st->src ? v.T2 : v.T) : v.T);
  double2* dst = reinterpret_cast<double2*>(OOP ? (st->src ? v.T : v.T2) : v.T);
  sweep_body<SKIP, OOP, EMIT, UNROLL>(src, dst, f, reinterpret_cast<const double2*>(v.prow), cn, v.rhs, v.R, v.C,
                                      v.ld, st->leave, EMIT ? st->next_enter : -1, eps, reverse);
}

// state reset before a solve
__global__ void k_state_reset(TabState* st, long long max_piv, int src) {
  st->status = LPR_RUNNING;
  st->enter = -1;
  st->leave = -1;
  st->next_enter = -1;
  st->do_sweep = 0;
  st->cur = 0;
  st->src = src;
  st->phase = 0;
  st->have_prev = 0;
  st->dropped = 0;
  st->npiv = 0;
  st->max_piv = max_piv;
  st->pivot = 0.0;
}

// ---- construction kernels ---------------------------------------------------------------------
__global__ void k_fill_zero(double* T, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t st = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += st) T[i] = 0.0;
}
// PrimalSimplexSolver..ctor :56-83 for the synthetic dense LP of SURVEY 8(d) (all rows <=, max)
__global__ void k_build_dense_lp(double* T, int ld, int m, int n, uint64_t seed, int* basis) {
  const int C = n + m + 1;
  const int row = blockIdx.y;  // 0..m
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < ld; j += gridDim.x * blockDim.x) {
    double val = 0.0;
    if (row == 0) {
      if (j < n) val = -(1.0 + u01(seed, 2, (uint64_t)j));  // :61-62 row 0 = -c
    } else {
      const int i = row - 1;
      if (j < n)
        val = 0.1 + u01(seed, 0, (uint64_t)i * (uint64_t)n + (uint64_t)j);
      else if (j == n + i)
        val = 1.0;  // :75-76
      else if (j == C - 1)
        val = ((double)n / 4.0) * (1.0 + u01(seed, 1, (uint64_t)i));  // :82
      if (j == 0) basis[i] = n + i;                                    // :78
    }
    TAT(T, ld, row, j) = val;
  }
}
// PrimalSimplexSolver..ctor :27-87 from a host-provided model already copied to the device
__global__ void k_build_primal(double* T, int ld, int m, int n, const double* obj, const double* coef,
                               int coef_stride, const int* coef_count, const int* relation,
                               const double* rhs, int is_max, int* basis) {
  const int C = n + m + 1;
  const int row = blockIdx.y;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < ld; j += gridDim.x * blockDim.x) {
    double val = 0.0;
    if (row == 0) {
      if (j < n) val = is_max ? -obj[j] : obj[j];
    } else {
      const int i = row - 1;
      const bool ge = relation && relation[i] == LPR_REL_GE;  // :36-41
      const int cnt = coef_count ? coef_count[i] : n;
      if (j < n) {
        if (j < cnt) {
          double a = coef[(size_t)i * coef_stride + j];
          val = ge ? -a : a;
        }
      } else if (j == n + i) {
        val = 1.0;
      } else if (j == C - 1) {
        val = ge ? -rhs[i] : rhs[i];
      }
      if (j == 0) basis[i] = n + i;
    }
    TAT(T, ld, row, j) = val;
  }
}

// same constructor, but the coefficient block was copied straight from the host into T[1.., 0..n) (no
// staging buffer): negate ">=" rows in place and fill everything around the block
__global__ void k_build_primal_inplace(double* T, int ld, int m, int n, const double* obj, const int* relation,
                                       const double* rhs, int is_max, int* basis) {
  const int C = n + m + 1;
  const int row = blockIdx.y;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < ld; j += gridDim.x * blockDim.x) {
    if (row == 0) {
      TAT(T, ld, 0, j) = (j < n) ? (is_max ? -obj[j] : obj[j]) : 0.0;
      continue;
    }
    const int i = row - 1;
    const bool ge = relation && relation[i] == LPR_REL_GE;
    if (j < n) {
      if (ge) TAT(T, ld, row, j) = -TAT(T, ld, row, j);
    } else {
      double val = 0.0;
      if (j == n + i) val = 1.0;
      else if (j == C - 1) val = ge ? -rhs[i] : rhs[i];
      TAT(T, ld, row, j) = val;
    }
    if (j == 0) basis[i] = n + i;
  }
}

// ExtractSolution PrimalSimplexSolver.cs:213-252: one warp per decision column
__global__ void k_extract_solution(TabView v, int n, double* x) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int nw = (gridDim.x * blockDim.x) >> 5;
  const int R = v.R, C = v.C, ld = v.ld;
  for (int j = warp; j < n; j += nw) {
    // sequential semantics: scan rows 1..R-1, stop at the first violation.  A column is basic iff
    // there is exactly one "1" and it is not preceded/followed by a violation before the scan
    // breaks; emulate by finding the first violating row and counting ones before it.
    int firstBad = INT_MAX, ones = 0, firstOne = INT_MAX, secondOne = INT_MAX;
    for (int i = 1 + lane; i < R; i += 32) {
      double t = TAT(v.T, ld, i, j);
      if (fabs(t - 1.0) < 1e-9) {
        ones++;
        if (i < firstOne) { secondOne = firstOne; firstOne = i; }
        else if (i < secondOne) secondOne = i;
      } else if (fabs(t) > 1e-9) {
        if (i < firstBad) firstBad = i;
      }
    }
    for (int o = 16; o > 0; o >>= 1) {
      int ob = __shfl_xor_sync(0xffffffffu, firstBad, o);
      int o1 = __shfl_xor_sync(0xffffffffu, firstOne, o);
      int o2 = __shfl_xor_sync(0xffffffffu, secondOne, o);
      firstBad = min(firstBad, ob);
      // merge two sorted pairs
      int a1 = min(firstOne, o1);
      int a2 = min(max(firstOne, o1), min(secondOne, o2));
      firstOne = a1;
      secondOne = a2;
    }
    (void)ones;
    // a second "1" or a non-zero entry anywhere makes the column non-basic (the scan would reach
    // it: breaks only happen at such rows), so position does not matter.
    bool basic = (firstOne != INT_MAX) && (secondOne == INT_MAX) && (firstBad == INT_MAX);
    if (lane == 0) x[j] = basic ? TAT(v.T, ld, firstOne, C - 1) : 0.0;
  }
}

// ---- Gomory cut (CuttingPlaneSolver.cs:76-107) ------------------------------------------------
__global__ void __launch_bounds__(kSelThreads) k_gomory_cut(TabView v, double* cut, int* chosen_out) {
  __shared__ MinIdx sm[32];
  const int R = v.R, C = v.C, ld = v.ld;
  // row whose RHS fractional part is closest to 0.5; first minimum (List.Sort on <= 16 entries is
  // an insertion sort; exact ties beyond that are unpinned in the reference)
  int k = block_first_min(R - 1, [&](int q, double& val) {
    double fr = net_frac(TAT(v.T, ld, q + 1, C - 1));
    if (!(fr > 1e-9)) return false;
    val = fabs(__dsub_rn(fr, 0.5));
    return true;
  }, sm);
  if (threadIdx.x == 0) *chosen_out = k;
  if (k < 0) return;
  for (int j = threadIdx.x; j < C; j += blockDim.x) cut[j] = -net_frac(TAT(v.T, ld, k + 1, j));
}
// append `row` (C doubles on device) as tableau row R
__global__ void k_append_row(double* T, int ld, int R, int C, const double* row) {
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < ld; j += gridDim.x * blockDim.x)
    TAT(T, ld, R, j) = (j < C) ? row[j] : 0.0;
}
// pivot column on the cut row (CuttingPlaneSolver.cs:113-132) + staging, then generic sweep
__global__ void __launch_bounds__(kSelThreads) k_cut_select(TabView v, int cut_row) {
  __shared__ MinIdx sm[32];
  __shared__ int smi[32];
  const double EPS = 1e-9;
  TabState* st = v.st;
  const int R = v.R, C = v.C, ld = v.ld;
  double* T = v.T;
  int e = block_hyst_min(C - 1, [&](int j, double& val) {
    double a = TAT(T, ld, cut_row, j);
    if (!(a < -EPS)) return false;
    double num = T[j];
    if (!(fabs(num) > EPS)) return false;
    val = fabs(ddiv(num, a));
    return true;
  }, kPosInf, EPS, sm, smi);
  if (e < 0) {
    if (threadIdx.x == 0) { st->status = LPR_NO_PIVOT_COL; st->do_sweep = 0; st->enter = -1; }
    return;
  }
  const double piv = TAT(T, ld, cut_row, e);
  if (fabs(piv) <= EPS) {
    if (threadIdx.x == 0) { st->status = LPR_PIVOT_TOO_SMALL; st->do_sweep = 0; }
    return;
  }
  double* colb = st->cur ? v.col[1] : v.col[0];
  for (int j = threadIdx.x; j < ld; j += blockDim.x) v.prow[j] = (j < C) ? ddiv(TAT(T, ld, cut_row, j), piv) : 0.0;
  for (int i = threadIdx.x; i < R; i += blockDim.x) colb[i] = TAT(T, ld, i, e);
  if (threadIdx.x == 0) {
    st->enter = e;
    st->leave = cut_row;
    st->next_enter = -1;
    st->do_sweep = 1;
    st->pivot = piv;
  }
}
// stage a caller-chosen pivot (lpr_tab_pivot_at)
__global__ void __launch_bounds__(kSelThreads) k_stage_pivot(TabView v, int p, int e) {
  TabState* st = v.st;
  const int R = v.R, C = v.C, ld = v.ld;
  const double piv = TAT(v.T, ld, p, e);
  double* colb = st->cur ? v.col[1] : v.col[0];
  for (int j = threadIdx.x; j < ld; j += blockDim.x) v.prow[j] = (j < C) ? ddiv(TAT(v.T, ld, p, j), piv) : 0.0;
  for (int i = threadIdx.x; i < R; i += blockDim.x) colb[i] = TAT(v.T, ld, i, e);
  if (threadIdx.x == 0) {
    st->enter = e;
    st->leave = p;
    st->next_enter = -1;
    st->do_sweep = 1;
    st->pivot = piv;
  }
}
// flags used by the cutting-plane driver (CuttingPlaneSolver.cs:19-45): out[0] any RHS < -eps,
// out[1] objective not optimal, out[2] any fractional RHS
__global__ void __launch_bounds__(kSelThreads) k_cut_flags(TabView v, int* out) {
  __shared__ int smi[32];
  const int R = v.R, C = v.C, ld = v.ld;
  int neg = 0, nopt = 0, fr = 0;
  for (int i = 1 + threadIdx.x; i < R; i += blockDim.x) {
    double rhs = TAT(v.T, ld, i, C - 1);
    if (rhs < -1e-9) neg++;
    if (net_frac(rhs) > 1e-9) fr++;
  }
  for (int j = threadIdx.x; j < C - 1; j += blockDim.x)
    if (v.T[j] < -1e-9) nopt++;
  neg = block_sum_int(neg, smi);
  nopt = block_sum_int(nopt, smi);
  fr = block_sum_int(fr, smi);
  if (threadIdx.x == 0) {
    out[0] = neg > 0;
    out[1] = nopt > 0;
    out[2] = fr > 0;
  }
}

// =============================================================================================
// host side
// =============================================================================================
// Small per-process cache of tableau-sized device buffers: solver objects are created and destroyed per
// Solve() by the host API, and cudaMalloc/cudaFree of 400 MB cost milliseconds to tens of milliseconds.
struct CachedBuf {
  int device;
  size_t bytes;
  double* ptr;
};
static std::vector<CachedBuf> g_buf_cache;
static std::mutex g_buf_mu;
static size_t cache_limit_bytes() {
  static const size_t lim = (size_t)std::max(0, env_int("LPR_CACHE_MB", 2048)) << 20;
  return lim;
}
static double* cache_take(int device, size_t bytes) {
  std::lock_guard<std::mutex> lk(g_buf_mu);
  for (size_t i = 0; i < g_buf_cache.size(); i++)
    if (g_buf_cache[i].device == device && g_buf_cache[i].bytes == bytes) {
      double* p = g_buf_cache[i].ptr;
      g_buf_cache.erase(g_buf_cache.begin() + i);
      return p;
    }
  return nullptr;
}
static bool cache_give(int device, size_t bytes, double* ptr) {
  if (!ptr || bytes < (8u << 20)) return false;  // only worth it for big buffers
  std::lock_guard<std::mutex> lk(g_buf_mu);
  size_t held = 0;
  for (auto& c : g_buf_cache) held += c.bytes;
  if (held + bytes > cache_limit_bytes()) return false;
  g_buf_cache.push_back(CachedBuf{device, bytes, ptr});
  return true;
}

// Whole-handle cache: a destroyed big handle keeps its tableau buffers, scratch, streams, events and pinned
// state mirror and is handed to the next tab_alloc of the same shape.  The host API creates one solver object
// per Solve(); cudaFreeHost / cudaMallocHost / a dozen cudaFree alone cost 8 ms and sometimes hundreds.
static std::vector<lpr_tab*> g_tab_cache;
static size_t tab_bytes(const lpr_tab* h) {
  return (size_t)h->Rcap * h->ld * sizeof(double) * (h->T2 ? 2 : 1);
}
static lpr_tab* tab_cache_take(int device, int rows, int cols, int rcap, int ccap) {
  std::lock_guard<std::mutex> lk(g_buf_mu);
  for (size_t i = 0; i < g_tab_cache.size(); i++) {
    lpr_tab* c = g_tab_cache[i];
    if (c->device == device && c->R == rows && c->C == cols && c->Rcap == rcap && c->Ccap == ccap) {
      g_tab_cache.erase(g_tab_cache.begin() + i);
      return c;
    }
  }
  return nullptr;
}
static bool tab_cache_give(lpr_tab* h) {
  if (tab_bytes(h) < (8u << 20)) return false;
  std::lock_guard<std::mutex> lk(g_buf_mu);
  size_t held = 0;
  for (auto& c : g_buf_cache) held += c.bytes;
  for (auto* c : g_tab_cache) held += tab_bytes(c);
  if (held + tab_bytes(h) > cache_limit_bytes()) return false;
  g_tab_cache.push_back(h);
  return true;
}
int tab_destroy_now(lpr_tab* h);
// out of device memory: give everything cached back to the driver (the caller retries once)
static void cache_flush(int device) {
  std::vector<lpr_tab*> tabs;
  std::vector<CachedBuf> bufs;
  {
    std::lock_guard<std::mutex> lk(g_buf_mu);
    for (size_t i = 0; i < g_tab_cache.size();)
      if (g_tab_cache[i]->device == device) { tabs.push_back(g_tab_cache[i]); g_tab_cache.erase(g_tab_cache.begin() + i); } else i++;
    for (size_t i = 0; i < g_buf_cache.size();)
      if (g_buf_cache[i].device == device) { bufs.push_back(g_buf_cache[i]); g_buf_cache.erase(g_buf_cache.begin() + i); } else i++;
  }
  for (auto& b : bufs) cudaFree(b.ptr);
  for (auto* t : tabs) {
    double* T = t->T; double* T2 = t->T2;
    t->T = t->T2 = nullptr;  // not back into the buffer cache
    cudaFree(T);
    cudaFree(T2);
    tab_destroy_now(t);
  }
}

static int tab_alloc_fresh(int device, int rows, int cols, int row_cap, int col_cap, lpr_tab** out);
int tab_alloc(int device, int rows, int cols, int row_cap, int col_cap, lpr_tab** out) {
  if (!out) return fail(LPR_E_BADARG, "out is null");
  *out = nullptr;
  if (rows < 1 || cols < 2) return fail(LPR_E_BADARG, "tableau needs rows >= 1 and cols >= 2 (got %d x %d)", rows, cols);
  int rc = select_device(device);
  if (rc) return rc;
  if (lpr_tab* c = tab_cache_take(device, rows, cols, std::max(rows, row_cap), std::max(cols, col_cap))) {
    // same shape as a destroyed handle: reuse everything, reset what a fresh handle guarantees
    cudaError_t e = cudaMemsetAsync(c->basis, 0xff, sizeof(int) * c->Rcap, c->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(c->st, 0, sizeof(TabState), c->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(c->ticket, 0, sizeof(unsigned), c->stream);
    if (e != cudaSuccess) {
      tab_destroy_now(c);
      return fail(LPR_E_CUDA, "handle reuse failed: %s", cudaGetErrorString(e));
    }
    c->last_ms = 0.f;
    c->last_sweep_us = 0.f;
    *out = c;
    return LPR_OK;
  }
  rc = tab_alloc_fresh(device, rows, cols, row_cap, col_cap, out);
  if (rc == LPR_E_NOMEM) {
    cache_flush(device);
    rc = tab_alloc_fresh(device, rows, cols, row_cap, col_cap, out);
  }
  return rc;
}

static int tab_alloc_fresh(int device, int rows, int cols, int row_cap, int col_cap, lpr_tab** out) {
  *out = nullptr;
  lpr_tab* h = new (std::nothrow) lpr_tab();
  if (!h) return fail(LPR_E_NOMEM, "host allocation failed");
  h->device = device;
  h->R = rows;
  h->C = cols;
  h->Rcap = std::max(rows, row_cap);
  h->Ccap = std::max(cols, col_cap);
  h->ld = round_up(h->Ccap, 16);
  h->sms = sm_count(device);
  const size_t bytes = (size_t)h->Rcap * h->ld * sizeof(double);
  cudaError_t e;
#define TRY(x)                                                                                   \
  if ((e = (x)) != cudaSuccess) {                                                                \
    tab_destroy_now(h);                                                                          \
    return fail(e == cudaErrorMemoryAllocation ? LPR_E_NOMEM : LPR_E_CUDA, "%s failed: %s", #x, \
                cudaGetErrorString(e));                                                          \
  }
  TRY(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  h->T = cache_take(device, bytes);
  if (!h->T) TRY(cudaMalloc(&h->T, bytes));
  TRY(cudaMalloc(&h->col[0], sizeof(double) * h->Rcap));
  TRY(cudaMalloc(&h->col[1], sizeof(double) * h->Rcap));
  TRY(cudaMalloc(&h->rhs, sizeof(double) * h->Rcap));
  TRY(cudaMalloc(&h->prow, sizeof(double) * h->ld));
  TRY(cudaMalloc(&h->basis, sizeof(int) * h->Rcap));
  TRY(cudaMalloc(&h->st, sizeof(TabState)));
  TRY(cudaMalloc(&h->selcand, sizeof(MinIdx) * (h->ld / 256 + 2)));
  TRY(cudaMalloc(&h->ticket, sizeof(unsigned)));
  TRY(cudaMemsetAsync(h->ticket, 0, sizeof(unsigned), h->stream));
  TRY(cudaMallocHost(&h->st_host, sizeof(TabState) * 2));
  TRY(cudaEventCreate(&h->ev0));
  TRY(cudaEventCreate(&h->ev1));
  TRY(cudaEventCreateWithFlags(&h->evb[0], cudaEventDisableTiming));
  TRY(cudaEventCreateWithFlags(&h->evb[1], cudaEventDisableTiming));
  TRY(cudaMemsetAsync(h->basis, 0xff, sizeof(int) * h->Rcap, h->stream));
  TRY(cudaMemsetAsync(h->st, 0, sizeof(TabState), h->stream));
#undef TRY
  *out = h;
  return LPR_OK;
}

int tab_ensure_log(lpr_tab* h, long long cap) {
  if (cap <= h->log_cap) return LPR_OK;
  if (h->log) cudaFree(h->log);
  h->log = nullptr;
  h->log_cap = 0;
  LPR_CUDA(cudaMalloc(&h->log, sizeof(int) * 2 * (size_t)cap));
  h->log_cap = cap;
  return LPR_OK;
}
int tab_ensure_T2(lpr_tab* h) {
  if (h->T2) return LPR_OK;
  const size_t bytes = (size_t)h->Rcap * h->ld * sizeof(double);
  h->T2 = cache_take(h->device, bytes);
  if (!h->T2) LPR_CUDA(cudaMalloc(&h->T2, bytes));
  return LPR_OK;
}

static int sweep_grid(const lpr_tab* h) {
  // waves of resident CTAs: grid = sms * (resident CTAs per SM) * LPR_SWEEP_WAVES, capped by the tile count
  static int resident = 0;
  if (!resident) {
    int r = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&r, k_sweep<0, false, true, 8>, kSweepThreads, 0);
    resident = std::max(1, r);
  }
  // measured on B200 (tools/tab_bench.py): one tile per CTA with 4 loads in flight per thread is the
  // fastest shape (123 us/pivot); LPR_SWEEP_WAVES / LPR_SWEEP_CTAS_PER_SM select a persistent grid instead
  static const int waves = std::max(0, env_int("LPR_SWEEP_WAVES", 0));
  static const int per_sm_env = env_int("LPR_SWEEP_CTAS_PER_SM", 0);
  const int per_sm = per_sm_env > 0 ? per_sm_env : (waves > 0 ? resident * waves : (1 << 20));
  static const int unroll_env = env_int("LPR_SWEEP_UNROLL", 4);
  const int unroll = (unroll_env == 4 || unroll_env == 16) ? unroll_env : 8;
  const long long chunks = (long long)h->R * (h->ld / 2);
  const long long tiles = (chunks + kSweepThreads * unroll - 1) / (kSweepThreads * unroll);
  long long g = (long long)h->sms * per_sm;
  g = std::max<long long>(1, std::min(g, tiles));
  return (int)g;
}

// kernel launch with the programmatic-stream-serialization attribute (PDL)
template <class... KArgs, class... Args>
static cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, cudaStream_t stream, Args... args) {
  static const int pdl = env_int("LPR_PDL", 1);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = 0;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// launch the sweep that applies the staged pivot; mode selects the instantiation
static int launch_sweep(lpr_tab* h, int skip, double eps, bool emit, int reverse) {
  const int g = sweep_grid(h);
  TabView v = h->view();
  static const int unroll = env_int("LPR_SWEEP_UNROLL", 4);
  if (emit && unroll == 4)
    launch_pdl(k_sweep<0, false, true, 4>, g, kSweepThreads, h->stream, v, eps, reverse);
  else if (emit && unroll == 16)
    launch_pdl(k_sweep<0, false, true, 16>, g, kSweepThreads, h->stream, v, eps, reverse);
  else if (emit)
    launch_pdl(k_sweep<0, false, true, 8>, g, kSweepThreads, h->stream, v, eps, reverse);
  else if (skip == 0)
    launch_pdl(k_sweep<0, false, false, 8>, g, kSweepThreads, h->stream, v, eps, reverse);
  else if (skip == 1)
    launch_pdl(k_sweep<1, false, false, 8>, g, kSweepThreads, h->stream, v, eps, reverse);
  else
    launch_pdl(k_sweep<2, false, false, 8>, g, kSweepThreads, h->stream, v, eps, reverse);
  LPR_LAUNCH_CHECK();
  return LPR_OK;
}
int launch_sweep_oop(lpr_tab* h) {
  const int g = sweep_grid(h);
  k_sweep<0, true, false, 8><<<g, kSweepThreads, 0, h->stream>>>(h->view(), 0.0, 0);
  LPR_LAUNCH_CHECK();
  return LPR_OK;
}

static int launch_select(lpr_tab* h, int rule, int flags, int* mask) {
  TabView v = h->view();
  switch (rule) {
    case LPR_RULE_PRIMAL:
      if (flags & F_FUSED)
        launch_pdl(k_primal_select_fused, (h->ld + 255) / 256, 256, h->stream, v, h->selcand, h->ticket);
      else
        launch_pdl(k_select<LPR_RULE_PRIMAL>, 1, kSelThreads, h->stream, v, flags, mask);
      break;
    case LPR_RULE_PRIMAL2: launch_pdl(k_select<LPR_RULE_PRIMAL2>, 1, kSelThreads, h->stream, v, flags, mask); break;
    case LPR_RULE_DUAL: launch_pdl(k_select<LPR_RULE_DUAL>, 1, kSelThreads, h->stream, v, flags, mask); break;
    case LPR_RULE_SENS: launch_pdl(k_select<LPR_RULE_SENS>, 1, kSelThreads, h->stream, v, flags, mask); break;
    default: return fail(LPR_E_BADARG, "unknown rule %d", rule);
  }
  LPR_LAUNCH_CHECK();
  return LPR_OK;
}

int tab_solve_blocked(lpr_tab* h, int K, int64_t max_pivots, int* status, int64_t* n_pivots, int* pivot_log,
                      int64_t log_cap, bool time_sweeps);
bool tab_pipe_applicable(const lpr_tab* h);
int tab_solve_pipelined(lpr_tab* h, int K, int64_t max_pivots, int* status, int64_t* n_pivots, int* pivot_log,
                        int64_t log_cap, bool time_sweeps);

bool tab_persist_applicable(const lpr_tab* h);
int tab_run_persistent(lpr_tab* h, int program, int64_t max_pivots, int print_steps, int max_cuts, int* status,
                       int64_t* n_pivots, int* pivot_log, int64_t log_cap, int* n_cuts, int* cut_log, int cut_log_cap);

int tab_solve_internal(lpr_tab* h, int rule, int64_t max_pivots, int flags, int* status,
                       int64_t* n_pivots, int* pivot_log, int64_t log_cap) {
  if (!h) return fail(LPR_E_BADARG, "null handle");
  if (rule < 0 || rule > LPR_RULE_SENS) return fail(LPR_E_BADARG, "unknown rule %d", rule);
  int rc = select_device(h->device);
  if (rc) return rc;
  {
    // delayed-update path: K pivots per tableau sweep (tableau_blocked.cu); flags bit4 or
    // LPR_TAB_BLOCK<=1 keep one sweep per pivot
    static const int blk = env_int("LPR_TAB_BLOCK", 16);
    static const int fused_on = env_int("LPR_TAB_FUSED", 1);
    if (rule == LPR_RULE_PRIMAL && fused_on && blk > 1 && !(flags & (4 | 16)) && (max_pivots < 0 || max_pivots > 1)) {
      // flags bit5 keeps selection and sweep on one stream (tableau_blocked.cu); default: overlapped
      if (!(flags & 32) && tab_pipe_applicable(h))
        return tab_solve_pipelined(h, blk, max_pivots, status, n_pivots, pivot_log, log_cap, (flags & 8) != 0);
      return tab_solve_blocked(h, blk, max_pivots, status, n_pivots, pivot_log, log_cap, (flags & 8) != 0);
    }
  }
  // PrimalSimplexSolver2 / DualSimplexSolver loops on an L2-resident tableau: one cooperative launch for the whole
  // loop (tableau_persistent.cu); flags bit2 keeps the two-kernel path
  if ((rule == LPR_RULE_DUAL || rule == LPR_RULE_PRIMAL2) && !(flags & 4) && tab_persist_applicable(h))
    return tab_run_persistent(h, rule == LPR_RULE_DUAL ? 0 : 1, max_pivots, flags & F_PRINT, 0, status, n_pivots,
                              pivot_log, log_cap, nullptr, nullptr, 0);
  static const int fused_default = env_int("LPR_TAB_FUSED", 1);
  static const int batch = std::max(1, env_int("LPR_TAB_BATCH", 32));
  static const int serp = env_int("LPR_TAB_SERPENTINE", 1);
  const bool fused = (rule == LPR_RULE_PRIMAL) && fused_default && !(flags & 4);
  int kflags = (flags & F_PRINT) | (fused ? F_FUSED : 0);
  if (pivot_log && log_cap > 0) {
    long long want = log_cap;
    if (max_pivots >= 0) want = std::min<long long>(want, max_pivots + 1);
    want = std::min<long long>(want, 1LL << 24);
    rc = tab_ensure_log(h, want);
    if (rc) return rc;
  }
  int* mask = nullptr;
  if (rule == LPR_RULE_SENS) LPR_CUDA(cudaMalloc(&mask, sizeof(int) * h->C));
  TabView v = h->view();
  if (!(pivot_log && log_cap > 0)) { v.log = nullptr; v.log_cap = 0; }
  lpr_tab view_holder = *h;  // shallow copy so launch helpers see the (possibly) disabled log
  view_holder.log = v.log;
  view_holder.log_cap = v.log_cap;
  lpr_tab* hv = &view_holder;

  int skip = 0;
  double eps = 0.0;
  if (rule == LPR_RULE_PRIMAL2) { skip = 1; eps = 1e-10; }
  if (rule == LPR_RULE_DUAL) { skip = 1; eps = 1e-9; }
  if (rule == LPR_RULE_SENS) { skip = 2; eps = 1e-9; }

  // flags bit 3: bracket every sweep launch with a CUDA event pair (roofline measurement pass)
  const bool time_sweeps = (flags & 8) != 0;
  std::vector<cudaEvent_t> sw_ev;
  const int kMaxTimed = 256;
  h->last_sweep_us = 0.f;
  LPR_CUDA(cudaEventRecord(h->ev0, h->stream));
  k_state_reset<<<1, 1, 0, h->stream>>>(h->st, (long long)max_pivots, 0);
  LPR_LAUNCH_CHECK();
  if (fused) {
    k_primal_init<<<1, kSelThreads, 0, h->stream>>>(v);
    LPR_LAUNCH_CHECK();
  }
  // Launch batches of (select, sweep) pairs ahead of the device; the status word of batch b is
  // inspected while batch b+1 is already queued, so the GPU never waits for the host.
  int final_status = LPR_RUNNING;
  long long pivot_parity = 0;
  int pending = 0;  // batches whose status copy has not been inspected yet
  int slot = 0;
  int bsize = std::min(batch, 4);  // small LPs finish in a few pivots: start small, grow
  while (true) {
    for (int b = 0; b < bsize; b++) {
      rc = launch_select(hv, rule, kflags, mask);
      if (rc) return rc;
      const bool timed = time_sweeps && (int)sw_ev.size() < 2 * kMaxTimed;
      if (timed) {
        cudaEvent_t a, b2;
        LPR_CUDA(cudaEventCreate(&a));
        LPR_CUDA(cudaEventCreate(&b2));
        sw_ev.push_back(a);
        sw_ev.push_back(b2);
        LPR_CUDA(cudaEventRecord(a, h->stream));
      }
      rc = launch_sweep(hv, skip, eps, fused, serp ? (int)(pivot_parity & 1) : 0);
      if (rc) return rc;
      if (timed) LPR_CUDA(cudaEventRecord(sw_ev.back(), h->stream));
      pivot_parity++;
    }
    // a trailing select commits the last sweep of the batch so the status word is current
    rc = launch_select(hv, rule, kflags, mask);
    if (rc) return rc;
    LPR_CUDA(cudaMemcpyAsync(&h->st_host[slot], h->st, sizeof(TabState), cudaMemcpyDeviceToHost, h->stream));
    LPR_CUDA(cudaEventRecord(h->evb[slot], h->stream));
    // the select above may have staged the next pivot: apply it before the next batch's select
    rc = launch_sweep(hv, skip, eps, fused, serp ? (int)(pivot_parity & 1) : 0);
    if (rc) return rc;
    pivot_parity++;
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
  LPR_CUDA(cudaEventRecord(h->ev1, h->stream));
  LPR_CUDA(cudaMemcpyAsync(&h->st_host[0], h->st, sizeof(TabState), cudaMemcpyDeviceToHost, h->stream));
  LPR_CUDA(cudaStreamSynchronize(h->stream));
  LPR_CUDA(cudaEventElapsedTime(&h->last_ms, h->ev0, h->ev1));
  final_status = h->st_host[0].status;
  const long long npiv = h->st_host[0].npiv;
  if (time_sweeps) {
    // only the first npiv sweeps did work (later launches are no-ops); skip the first few (warm-up)
    double sum = 0.0;
    int cnt = 0;
    const int real = (int)std::min<long long>(npiv, (long long)sw_ev.size() / 2);
    for (int k = std::min(4, real / 2); k < real; k++) {
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, sw_ev[2 * k], sw_ev[2 * k + 1]) == cudaSuccess) {
        sum += ms;
        cnt++;
      }
    }
    if (cnt) h->last_sweep_us = (float)(sum * 1e3 / cnt);
    for (auto ev : sw_ev) cudaEventDestroy(ev);
  }
  if (mask) cudaFree(mask);
  if (status) *status = final_status;
  if (n_pivots) *n_pivots = npiv;
  if (pivot_log && log_cap > 0 && npiv > 0) {
    long long cnt = std::min<long long>(std::min<long long>(npiv, log_cap), h->log_cap);
    LPR_CUDA(cudaMemcpy(pivot_log, h->log, sizeof(int) * 2 * (size_t)cnt, cudaMemcpyDeviceToHost));
  }
  return LPR_OK;
}

int tab_destroy_now(lpr_tab* h) {
  if (!h) return LPR_OK;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  if (!cache_give(h->device, (size_t)h->Rcap * h->ld * sizeof(double), h->T)) cudaFree(h->T);
  if (!cache_give(h->device, (size_t)h->Rcap * h->ld * sizeof(double), h->T2)) cudaFree(h->T2);
  tab_pipe_free(h);
  cudaFree(h->col[0]);
  cudaFree(h->col[1]);
  cudaFree(h->rhs);
  cudaFree(h->prow);
  cudaFree(h->basis);
  cudaFree(h->st);
  cudaFree(h->selcand);
  cudaFree(h->ticket);
  cudaFree(h->blk_pr);
  cudaFree(h->blk_f);
  cudaFree(h->blk_row0);
  cudaFree(h->blk_rhs[0]);
  cudaFree(h->blk_rhs[1]);
  cudaFree(h->blk_p);
  cudaFree(h->log);
  if (h->st_host) cudaFreeHost(h->st_host);
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  if (h->evb[0]) cudaEventDestroy(h->evb[0]);
  if (h->evb[1]) cudaEventDestroy(h->evb[1]);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return LPR_OK;
}

}  // namespace lpr

using namespace lpr;

// =============================================================================================
// C ABI
// =============================================================================================
extern "C" {

int lpr_version(void) { return LPR_VERSION; }
const char* lpr_last_error(void) { return last_error().c_str(); }
int lpr_device_count(int* count) {
  if (!count) return fail(LPR_E_BADARG, "count is null");
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    *count = 0;
    return fail(LPR_E_CUDA, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
  }
  *count = n;
  return LPR_OK;
}
int64_t lpr_launch_count(void) { return g_launches.load(); }

int lpr_tab_destroy(lpr_tab* h) {
  if (!h) return LPR_OK;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  if (h->T && h->stream && h->st_host && tab_cache_give(h)) return LPR_OK;
  return tab_destroy_now(h);
}


int lpr_tab_upload(lpr_tab* h, const double* host) {
  if (!h || !host) return fail(LPR_E_BADARG, "null argument");
  int rc = select_device(h->device);
  if (rc) return rc;
  const size_t n = (size_t)h->Rcap * h->ld;
  k_fill_zero<<<h->sms * 4, 256, 0, h->stream>>>(h->T, n);
  LPR_LAUNCH_CHECK();
  LPR_CUDA(cudaMemcpy2DAsync(h->T, sizeof(double) * h->ld, host, sizeof(double) * h->C, sizeof(double) * h->C,
                             h->R, cudaMemcpyHostToDevice, h->stream));
  LPR_CUDA(cudaStreamSynchronize(h->stream));
  return LPR_OK;
}

int lpr_tab_create(int device, int rows, int cols, int row_cap, int col_cap, const double* host, lpr_tab** out) {
  lpr_tab* h = nullptr;
  int rc = tab_alloc(device, rows, cols, row_cap, col_cap, &h);
  if (rc) return rc;
  if (host) {
    rc = lpr_tab_upload(h, host);
  } else {
    k_fill_zero<<<h->sms * 4, 256, 0, h->stream>>>(h->T, (size_t)h->Rcap * h->ld);
    count_launch();
    rc = cudaStreamSynchronize(h->stream) == cudaSuccess ? LPR_OK : fail(LPR_E_CUDA, "zero fill failed");
  }
  if (rc) {
    lpr_tab_destroy(h);
    return rc;
  }
  // default basis = trailing identity block (slack columns), as PrimalSimplexSolver.cs:78
  std::vector<int> b(std::max(1, rows - 1));
  for (int i = 0; i < rows - 1; i++) b[i] = cols - rows + i;
  if (rows > 1) cudaMemcpy(h->basis, b.data(), sizeof(int) * (rows - 1), cudaMemcpyHostToDevice);
  *out = h;
  return LPR_OK;
}

int lpr_tab_create_primal(int device, int n, int m, const double* objective, const double* coef, int coef_stride,
                          const int* coef_count, const int* relation, const double* rhs, int is_max,
                          lpr_tab** out) {
  if (n < 1 || m < 1 || !objective || !coef || !rhs || coef_stride < 1)
    return fail(LPR_E_BADARG, "bad model (n=%d m=%d)", n, m);
  lpr_tab* h = nullptr;
  int rc = tab_alloc(device, m + 1, n + m + 1, 0, 0, &h);
  if (rc) return rc;
  double *d_obj = nullptr, *d_coef = nullptr, *d_rhs = nullptr;
  int *d_cnt = nullptr, *d_rel = nullptr;
  auto cleanup = [&]() {
    cudaFree(d_obj); cudaFree(d_coef); cudaFree(d_rhs); cudaFree(d_cnt); cudaFree(d_rel);
  };
#define TRY(x)                                                    \
  if ((x) != cudaSuccess) {                                       \
    cleanup();                                                    \
    lpr_tab_destroy(h);                                           \
    return fail(LPR_E_CUDA, "%s failed: %s", #x, cudaGetErrorString(cudaGetLastError())); \
  }
  TRY(cudaMalloc(&d_obj, sizeof(double) * n));
  TRY(cudaMalloc(&d_rhs, sizeof(double) * m));
  TRY(cudaMemcpyAsync(d_obj, objective, sizeof(double) * n, cudaMemcpyHostToDevice, h->stream));
  TRY(cudaMemcpyAsync(d_rhs, rhs, sizeof(double) * m, cudaMemcpyHostToDevice, h->stream));
  if (relation) {
    TRY(cudaMalloc(&d_rel, sizeof(int) * m));
    TRY(cudaMemcpyAsync(d_rel, relation, sizeof(int) * m, cudaMemcpyHostToDevice, h->stream));
  }
  bool full_rows = coef_stride >= n;
  if (coef_count)
    for (int i = 0; i < m && full_rows; i++) full_rows = coef_count[i] >= n;
  dim3 grid(std::max(1, std::min(64, (h->ld + 255) / 256)), m + 1);
  if (full_rows) {
    // every row provides its n coefficients: DMA them straight into the tableau (H2D once, no staging)
    TRY(cudaMemcpy2DAsync(h->T + h->ld, sizeof(double) * h->ld, coef, sizeof(double) * coef_stride, sizeof(double) * n, m,
                          cudaMemcpyHostToDevice, h->stream));
    k_build_primal_inplace<<<grid, 256, 0, h->stream>>>(h->T, h->ld, m, n, d_obj, d_rel, d_rhs, is_max, h->basis);
  } else {
    TRY(cudaMalloc(&d_coef, sizeof(double) * (size_t)m * coef_stride));
    TRY(cudaMemcpyAsync(d_coef, coef, sizeof(double) * (size_t)m * coef_stride, cudaMemcpyHostToDevice, h->stream));
    if (coef_count) {
      TRY(cudaMalloc(&d_cnt, sizeof(int) * m));
      TRY(cudaMemcpyAsync(d_cnt, coef_count, sizeof(int) * m, cudaMemcpyHostToDevice, h->stream));
    }
    k_build_primal<<<grid, 256, 0, h->stream>>>(h->T, h->ld, m, n, d_obj, d_coef, coef_stride, d_cnt, d_rel, d_rhs,
                                                is_max, h->basis);
  }
  count_launch();
  TRY(cudaGetLastError());
  TRY(cudaStreamSynchronize(h->stream));
#undef TRY
  cleanup();
  *out = h;
  return LPR_OK;
}

int lpr_tab_create_dense_lp(int device, uint64_t seed, int m, int n, lpr_tab** out) {
  if (n < 1 || m < 1) return fail(LPR_E_BADARG, "bad shape (m=%d n=%d)", m, n);
  lpr_tab* h = nullptr;
  int rc = tab_alloc(device, m + 1, n + m + 1, 0, 0, &h);
  if (rc) return rc;
  dim3 grid(std::max(1, std::min(64, (h->ld + 255) / 256)), m + 1);
  k_build_dense_lp<<<grid, 256, 0, h->stream>>>(h->T, h->ld, m, n, seed, h->basis);
  count_launch();
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  if (e != cudaSuccess) {
    lpr_tab_destroy(h);
    return fail(LPR_E_CUDA, "dense LP build failed: %s", cudaGetErrorString(e));
  }
  *out = h;
  return LPR_OK;
}

int lpr_tab_dims(const lpr_tab* h, int* rows, int* cols, int* ld) {
  if (!h) return fail(LPR_E_BADARG, "null handle");
  if (rows) *rows = h->R;
  if (cols) *cols = h->C;
  if (ld) *ld = h->ld;
  return LPR_OK;
}

int lpr_tab_read(lpr_tab* h, double* host) {
  if (!h || !host) return fail(LPR_E_BADARG, "null argument");
  int rc = select_device(h->device);
  if (rc) return rc;
  LPR_CUDA(cudaMemcpy2DAsync(host, sizeof(double) * h->C, h->T, sizeof(double) * h->ld, sizeof(double) * h->C, h->R,
                             cudaMemcpyDeviceToHost, h->stream));
  LPR_CUDA(cudaStreamSynchronize(h->stream));
  return LPR_OK;
}
int lpr_tab_read_row(lpr_tab* h, int row, double* host) {
  if (!h || !host || row < 0 || row >= h->R) return fail(LPR_E_BADARG, "bad row %d", row);
  int rc = select_device(h->device);
  if (rc) return rc;
  LPR_CUDA(cudaMemcpyAsync(host, h->T + (size_t)row * h->ld, sizeof(double) * h->C, cudaMemcpyDeviceToHost, h->stream));
  LPR_CUDA(cudaStreamSynchronize(h->stream));
  return LPR_OK;
}
int lpr_tab_read_col(lpr_tab* h, int col, double* host) {
  if (!h || !host || col < 0 || col >= h->C) return fail(LPR_E_BADARG, "bad col %d", col);
  int rc = select_device(h->device);
  if (rc) return rc;
  LPR_CUDA(cudaMemcpy2DAsync(host, sizeof(double), h->T + col, sizeof(double) * h->ld, sizeof(double), h->R,
                             cudaMemcpyDeviceToHost, h->stream));
  LPR_CUDA(cudaStreamSynchronize(h->stream));
  return LPR_OK;
}
int lpr_tab_get_basis(lpr_tab* h, int* basis) {
  if (!h || !basis) return fail(LPR_E_BADARG, "null argument");
  int rc = select_device(h->device);
  if (rc) return rc;
  if (h->R > 1) {
    LPR_CUDA(cudaMemcpyAsync(basis, h->basis, sizeof(int) * (h->R - 1), cudaMemcpyDeviceToHost, h->stream));
    LPR_CUDA(cudaStreamSynchronize(h->stream));
  }
  return LPR_OK;
}
int lpr_tab_set_basis(lpr_tab* h, const int* basis) {
  if (!h || !basis) return fail(LPR_E_BADARG, "null argument");
  int rc = select_device(h->device);
  if (rc) return rc;
  if (h->R > 1) {
    LPR_CUDA(cudaMemcpyAsync(h->basis, basis, sizeof(int) * (h->R - 1), cudaMemcpyHostToDevice, h->stream));
    LPR_CUDA(cudaStreamSynchronize(h->stream));
  }
  return LPR_OK;
}

int lpr_tab_solve(lpr_tab* h, int rule, int64_t max_pivots, int flags, int* status, int64_t* n_pivots,
                  int* pivot_log, int64_t log_cap) {
  return tab_solve_internal(h, rule, max_pivots, flags, status, n_pivots, pivot_log, log_cap);
}

int lpr_tab_step(lpr_tab* h, int rule, int* enter_col, int* leave_row, int* status) {
  int st = 0;
  int64_t np = 0;
  int log[2] = {-1, -1};
  // PRIMAL2 / DUAL test their counter after the pivot: a cap of 1 with the print flag stops
  // after exactly one pivot; the other rules test before the pivot.
  int flags = (rule == LPR_RULE_PRIMAL2 || rule == LPR_RULE_DUAL) ? 1 : 0;
  int rc = tab_solve_internal(h, rule, 1, flags, &st, &np, log, 1);
  if (rc) return rc;
  if (np >= 1) {
    if (leave_row) *leave_row = log[0];
    if (enter_col) *enter_col = log[1];
    st = LPR_RUNNING;  // a pivot was applied; a terminal state is reported by the next call
  } else {
    if (leave_row) *leave_row = -1;
    if (enter_col) *enter_col = -1;
  }
  if (status) *status = st;
  return LPR_OK;
}

int lpr_tab_pivot_at(lpr_tab* h, int row, int col, double skip_eps, int skip_mode) {
  if (!h || row < 0 || row >= h->R || col < 0 || col >= h->C) return fail(LPR_E_BADARG, "bad pivot (%d,%d)", row, col);
  if (skip_mode < 0 || skip_mode > 2) return fail(LPR_E_BADARG, "bad skip_mode %d", skip_mode);
  int rc = select_device(h->device);
  if (rc) return rc;
  k_state_reset<<<1, 1, 0, h->stream>>>(h->st, -1, 0);
  LPR_LAUNCH_CHECK();
  k_stage_pivot<<<1, kSelThreads, 0, h->stream>>>(h->view(), row, col);
  LPR_LAUNCH_CHECK();
  rc = launch_sweep(h, skip_mode, skip_eps, false, 0);
  if (rc) return rc;
  LPR_CUDA(cudaStreamSynchronize(h->stream));
  return LPR_OK;
}

int lpr_tab_extract_solution(lpr_tab* h, int n, double* x) {
  if (!h || !x || n < 0 || n > h->C - 1) return fail(LPR_E_BADARG, "bad n %d", n);
  int rc = select_device(h->device);
  if (rc) return rc;
  if (n == 0) return LPR_OK;
  double* dx = nullptr;
  LPR_CUDA(cudaMalloc(&dx, sizeof(double) * n));
  const int blocks = std::max(1, std::min(h->sms * 8, (n + 7) / 8));
  k_extract_solution<<<blocks, 256, 0, h->stream>>>(h->view(), n, dx);
  count_launch();
  cudaError_t e = cudaMemcpyAsync(x, dx, sizeof(double) * n, cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  cudaFree(dx);
  if (e != cudaSuccess) return fail(LPR_E_CUDA, "extract_solution: %s", cudaGetErrorString(e));
  return LPR_OK;
}

int lpr_tab_objective(lpr_tab* h, double* z) {
  if (!h || !z) return fail(LPR_E_BADARG, "null argument");
  int rc = select_device(h->device);
  if (rc) return rc;
  LPR_CUDA(cudaMemcpyAsync(z, h->T + (h->C - 1), sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  LPR_CUDA(cudaStreamSynchronize(h->stream));
  return LPR_OK;
}

int lpr_tab_last_solve_ms(const lpr_tab* h, float* ms) {
  if (!h || !ms) return fail(LPR_E_BADARG, "null argument");
  *ms = h->last_ms;
  return LPR_OK;
}

int lpr_tab_last_sweep_us(const lpr_tab* h, float* us) {
  if (!h || !us) return fail(LPR_E_BADARG, "null argument");
  *us = h->last_sweep_us;
  return LPR_OK;
}

int lpr_tab_append_row(lpr_tab* h, const double* row) {
  if (!h || !row) return fail(LPR_E_BADARG, "null argument");
  if (h->R + 1 > h->Rcap) return fail(LPR_E_CAPACITY, "no row headroom (rows=%d cap=%d)", h->R, h->Rcap);
  int rc = select_device(h->device);
  if (rc) return rc;
  double* d = nullptr;
  LPR_CUDA(cudaMalloc(&d, sizeof(double) * h->C));
  cudaError_t e = cudaMemcpyAsync(d, row, sizeof(double) * h->C, cudaMemcpyHostToDevice, h->stream);
  if (e == cudaSuccess) {
    k_append_row<<<std::max(1, (h->ld + 255) / 256), 256, 0, h->stream>>>(h->T, h->ld, h->R, h->C, d);
    count_launch();
    e = cudaStreamSynchronize(h->stream);
  }
  cudaFree(d);
  if (e != cudaSuccess) return fail(LPR_E_CUDA, "append_row: %s", cudaGetErrorString(e));
  h->R += 1;
  return LPR_OK;
}

int lpr_tab_gomory_cut(lpr_tab* h, int* chosen_row, double* cut_host, int append) {
  if (!h) return fail(LPR_E_BADARG, "null handle");
  int rc = select_device(h->device);
  if (rc) return rc;
  if (append && h->R + 1 > h->Rcap) return fail(LPR_E_CAPACITY, "no row headroom (rows=%d cap=%d)", h->R, h->Rcap);
  double* dcut = nullptr;
  int* dch = nullptr;
  LPR_CUDA(cudaMalloc(&dcut, sizeof(double) * h->C));
  cudaError_t e = cudaMalloc(&dch, sizeof(int));
  int chosen = -1;
  if (e == cudaSuccess) {
    k_gomory_cut<<<1, kSelThreads, 0, h->stream>>>(h->view(), dcut, dch);
    count_launch();
    e = cudaMemcpyAsync(&chosen, dch, sizeof(int), cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  }
  if (e == cudaSuccess && chosen >= 0) {
    if (cut_host) e = cudaMemcpy(cut_host, dcut, sizeof(double) * h->C, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && append) {
      k_append_row<<<std::max(1, (h->ld + 255) / 256), 256, 0, h->stream>>>(h->T, h->ld, h->R, h->C, dcut);
      count_launch();
      e = cudaStreamSynchronize(h->stream);
      if (e == cudaSuccess) h->R += 1;
    }
  }
  cudaFree(dcut);
  cudaFree(dch);
  if (e != cudaSuccess) return fail(LPR_E_CUDA, "gomory_cut: %s", cudaGetErrorString(e));
  if (chosen_row) *chosen_row = chosen;
  return LPR_OK;
}

// CuttingPlaneSolver.CuttingPlaneSolution (CuttingPlaneSolver.cs:64-229); the recursion of :220
// is a loop.  Every tableau operation runs on the device; the host only sequences the steps.
int lpr_tab_cutting_plane(lpr_tab* h, int max_cuts, int* status, int* n_cuts, int* cut_log, int cut_log_cap) {
  if (!h) return fail(LPR_E_BADARG, "null handle");
  int rc = select_device(h->device);
  if (rc) return rc;
  if (tab_persist_applicable(h)) {  // the whole recursion in one cooperative launch (tableau_persistent.cu)
    int64_t piv = 0;
    return tab_run_persistent(h, 2, 10000, 1, max_cuts, status, &piv, nullptr, 0, n_cuts, cut_log, cut_log_cap);
  }
  int* dflags = nullptr;
  LPR_CUDA(cudaMalloc(&dflags, sizeof(int) * 4));
  int st = LPR_RUNNING, cuts = 0;
  auto read_flags = [&](int* f3) -> int {
    k_cut_flags<<<1, kSelThreads, 0, h->stream>>>(h->view(), dflags);
    count_launch();
    cudaError_t e = cudaMemcpyAsync(f3, dflags, sizeof(int) * 3, cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    return e == cudaSuccess ? LPR_OK : fail(LPR_E_CUDA, "cut flags: %s", cudaGetErrorString(e));
  };
  while (true) {
    if (max_cuts >= 0 && cuts >= max_cuts) { st = LPR_ITER_LIMIT; break; }
    if (h->R + 1 > h->Rcap) { st = LPR_ITER_LIMIT; break; }
    int chosen = -1;
    rc = lpr_tab_gomory_cut(h, &chosen, nullptr, 1);  // steps 1-5 (:76-111)
    if (rc) break;
    if (chosen < 0) { st = LPR_NO_CUT_NEEDED; break; }
    const int cut_row = h->R - 1;
    int nd = 0, np = 0, pcol = -1;
    auto log_cut = [&]() {
      if (cut_log && cuts < cut_log_cap) {
        cut_log[4 * cuts + 0] = chosen;
        cut_log[4 * cuts + 1] = pcol;
        cut_log[4 * cuts + 2] = nd;
        cut_log[4 * cuts + 3] = np;
      }
      cuts++;
    };
    // step 6-7 (:113-176): pivot on the cut row
    k_state_reset<<<1, 1, 0, h->stream>>>(h->st, -1, 0);
    count_launch();
    k_cut_select<<<1, kSelThreads, 0, h->stream>>>(h->view(), cut_row);
    count_launch();
    rc = launch_sweep(h, 1, 1e-9, false, 0);
    if (rc) break;
    TabState hs;
    if (cudaMemcpyAsync(&hs, h->st, sizeof hs, cudaMemcpyDeviceToHost, h->stream) != cudaSuccess ||
        cudaStreamSynchronize(h->stream) != cudaSuccess) {
      rc = fail(LPR_E_CUDA, "cutting plane state read failed");
      break;
    }
    pcol = hs.enter;
    if (hs.status == LPR_NO_PIVOT_COL || hs.status == LPR_PIVOT_TOO_SMALL) {
      log_cut();
      st = hs.status;
      break;
    }
    int fl[3];
    if ((rc = read_flags(fl))) break;
    bool needDual = fl[0], needPrimal = fl[1];
    if (needDual) {  // :186-194, printSteps: true
      int dst = 0;
      int64_t dn = 0;
      rc = tab_solve_internal(h, LPR_RULE_DUAL, 10000, 1, &dst, &dn, nullptr, 0);
      if (rc) break;
      nd = (int)dn;
      if (dst != LPR_OPTIMAL) {
        log_cut();
        st = dst == LPR_PIVOT_TOO_SMALL ? LPR_PIVOT_TOO_SMALL : LPR_INFEASIBLE;
        break;
      }
      if ((rc = read_flags(fl))) break;
      needPrimal = fl[1];
    }
    if (needPrimal) {  // :196-212, printSteps: true; the bool result is ignored by the reference
      int pst = 0;
      int64_t pn = 0;
      rc = tab_solve_internal(h, LPR_RULE_PRIMAL2, 10000, 1, &pst, &pn, nullptr, 0);
      if (rc) break;
      np = (int)pn;
      if (pst == LPR_PIVOT_TOO_SMALL) {
        log_cut();
        st = LPR_PIVOT_TOO_SMALL;
        break;
      }
    }
    log_cut();
    if ((rc = read_flags(fl))) break;
    if (!fl[1] && !fl[0]) {  // :215-226
      if (fl[2]) continue;
      st = LPR_OPTIMAL;
      break;
    }
    st = LPR_CUT_STEP_DONE;  // :228
    break;
  }
  cudaFree(dflags);
  if (rc) return rc;
  if (status) *status = st;
  if (n_cuts) *n_cuts = cuts;
  return LPR_OK;
}

}  // extern "C"
