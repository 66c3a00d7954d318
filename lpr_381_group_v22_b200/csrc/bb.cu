// bb.cu -- BranchBoundSimplexSolver (IntegerProgramming/BranchBoundSimplexSolver.cs) on the device.
//
// A node of the tree IS its final (4-d.p. rounded) tableau, exactly as in the reference, which warm
// starts every child from the parent's final tableau (:1105-1108).  Per batch of open nodes:
//   k_bb_eval_x/pick GetObjective / ExtractSolution / CheckIntegerBasicVar        (:805-857,:892-921)
//   k_bb_addc_build  AddConstraint: round, basic-column detection, bound row      (:642-747)
//   k_bb_addc_elim   ... ordering and elimination with 4-d.p. rounding per step   (:664-685,:752-803)
//   k_bb_select +    DoDualSimplex with tableauOverride: dual phase, primal phase, (:115-279,:289-468)
//   k_bb_sweep       out-of-place Gauss-Jordan pivots, -0.0 -> 0.0, "drop last tableau" quirk
//   k_bb_round       RoundAllTableaux on the children                               (:1124,:1187)
// Every kernel is batched over the node LPs of the round (blockIdx.y / one CTA per LP), which is
// what turns a launch-latency bound tree walk into an HBM/L2 bound one.  The host keeps the
// open-node stack, DFS keys and the incumbent.  With batch = 1 the visit order, 20-node cap and
// strict-improvement incumbent of ExecuteBranchAndBound (:1006-1233) are reproduced exactly.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <functional>
#include <vector>

#include "sweep.cuh"
#include "tableau.cuh"

namespace lpr {

struct BBLp {
  double* buf[2];  // ping-pong tableaux (out-of-place pivots)
  double* col;     // factor column scratch
  double* prow;    // pivot row scratch (ld doubles)
  int* log;        // optional (row, col) pairs
  int log_cap;
  int R, C, ld;
  // state
  int status, src, phase, do_sweep, leave, enter, dropped, entered_primal;
  long long npiv, max_piv;
  int is_min;  // DoDualSimplex's isMinimization (only the initial solve of RunBranchAndBound :1271 passes true)
};

struct BBEval {  // per node outputs of k_bb_eval
  double z;
  double branch_val;
  int branch_var;
  int all_integer;
};

struct BBAddc {  // one AddConstraint job
  const double* parent;
  double* child;
  int* key;     // C ints, preset to kAddcNoRow: first-"1" row of basic columns, -1 for non-basic
  int* order;   // C ints: basic columns in elimination order
  double* ksum; // C doubles, zeroed: exact integer column sums of the 1e4-scaled entries
  int* big;     // set to 1 when the job leaves the division-free / order-free range (see k_bb_addc_build)
  int R, C;     // parent dims
  int ldp, ldc;
  int n_vars, var, type;
  int negzero;  // 1: write the child with -0.0 -> 0.0 already applied (the first step of DoDualSimplex :307-313)
  double bound;
};

// Control words of a batched DoDualSimplex run (device ints): the LPs that staged a pivot in round r are appended
// to list[r & 1] by k_bb_select and swept by k_bb_sweep(r); count[r & 3] is their number (k_bb_sweep(r) clears
// count[(r + 2) & 3] for the select after next).
constexpr int kCtlRunning = 4, kCtlBar = 5, kCtlLists = 8;  // [kCtlBar]: grid-barrier counter of k_bb_chains

constexpr int kBBT = 1024;

// ---- elementwise helpers --------------------------------------------------------------------
// FormulateTableau :28-113 on the device: row 0 = -objective; constraint row i (ragged: [coefficients..., rhs,
// type flag], len[i] entries) is negated entirely when its flag == 1 (-1 * x, so zeros become -0.0 as in the
// reference), every entry but rhs/flag goes to columns 0.., the rhs to the last column, and row i gets a 1 at
// column i+n-1 whatever its type (:103-109).
__global__ void k_bb_formulate(double* T, int ld, int n, int m, const double* obj, const double* cons, int stride,
                               const int* len) {
  const int W = n + m + 1;
  const size_t total = (size_t)(m + 1) * ld;
  for (size_t q = blockIdx.x * (size_t)blockDim.x + threadIdx.x; q < total; q += (size_t)gridDim.x * blockDim.x) {
    const int i = (int)(q / ld), j = (int)(q - (size_t)i * ld);
    double v = 0.0;
    if (i == 0) {
      if (j < n) v = -obj[j];
    } else if (j < W) {
      const double* row = cons + (size_t)(i - 1) * stride;
      const int L = len[i - 1];
      const bool neg = row[L - 1] == 1.0;
      if (j < L - 2) v = neg ? __dmul_rn(-1.0, row[j]) : row[j];
      if (j == W - 1) v = neg ? __dmul_rn(-1.0, row[L - 2]) : row[L - 2];
      if (j == i + n - 1 && j < W - 1) v = 1.0;
    }
    T[q] = v;
  }
}

__global__ void k_bb_negzero(BBLp* lps) {
  BBLp& lp = lps[blockIdx.y];
  double* T = lp.buf[lp.src];
  const size_t n = (size_t)lp.R * lp.ld;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    if (T[i] == 0.0) T[i] = 0.0;  // :307-313
}
// `dirty` (optional, one int per LP, written by k_bb_addc_*): 0 = the LP's start tableau was 4-d.p. rounded, free of
// -0.0 and inside net_round4's idempotent range, so an LP that ended without a pivot is already its own rounding
__global__ void k_bb_round(BBLp* lps, const int* dirty) {
  BBLp& lp = lps[blockIdx.y];
  if (lp.status != LPR_OPTIMAL) return;
  if (dirty && !dirty[blockIdx.y] && lp.npiv == 0) return;
  double* T = lp.buf[lp.src];
  const size_t n = (size_t)lp.R * lp.ld;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    T[i] = net_round4(T[i]);  // RoundTableau :552-567
}
__global__ void k_round4(double* T, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    T[i] = net_round4(T[i]);
}

// ---- DoDualSimplex state machine (:289-468), one CTA per LP --------------------------------------
__device__ __forceinline__ void bb_select_body(BBLp* lps, int lp_index, int* ctl, int round, int cap) {
  __shared__ MinIdx sm[32];
  __shared__ int smi[32];
  BBLp& lp = lps[lp_index];
  const int tid = threadIdx.x;
  const int status = lp.status;
  const int did = lp.do_sweep;
  __syncthreads();
  if (status != LPR_RUNNING) return;
  int src = lp.src, phase = lp.phase;
  long long npiv = lp.npiv;
  if (did) {
    src ^= 1;
    npiv++;
  }
  const int R = lp.R, C = lp.C, ld = lp.ld;
  const double* T = lp.buf[src];
  auto finish = [&](int s, int new_src, long long np, int dropped) {
    if (tid == 0) {
      lp.status = s;
      lp.do_sweep = 0;
      lp.src = new_src;
      lp.npiv = np;
      lp.phase = phase;
      lp.dropped = dropped;
    }
  };
  int r = -1, c = -1;
  bool go_final = false;
  if (phase == 0) {
    // all RHS (objective row included) >= -1e-9 ?  (:315-320)
    int bad = 0;
    for (int i = tid; i < R; i += blockDim.x)
      if (!(TAT(T, ld, i, C - 1) >= -1e-9)) bad++;
    bad = block_sum_int(bad, smi);
    if (bad) {
      if (lp.max_piv >= 0 && npiv >= lp.max_piv) { finish(LPR_ITER_LIMIT, src, npiv, 0); return; }
      // PerformDualPivot :115-201: row = first index of the most negative RHS (exact < 0)
      r = block_first_min(R, [&](int i, double& val) { val = TAT(T, ld, i, C - 1); return val < 0.0; }, sm);
      if (r >= 0) {
        int other = 0;
        for (int j = tid; j < C - 1; j += blockDim.x) {
          double a = TAT(T, ld, r, j);
          double th = (a < 0.0) ? fabs(ddiv(T[j], a)) : kPosInf;
          if (!(th == 0.0 || th == kPosInf)) other++;
        }
        other = block_sum_int(other, smi);
        if (other == 0) {  // all thetas in {0, inf}: IndexOf(0)
          c = block_first_min(C - 1, [&](int j, double& val) {
            double a = TAT(T, ld, r, j);
            if (!(a < 0.0)) return false;
            val = fabs(ddiv(T[j], a));
            return val == 0.0;
          }, sm);
        } else {  // min over theta > 0 (inf included), IndexOf => first
          c = block_first_min(C - 1, [&](int j, double& val) {
            double a = TAT(T, ld, r, j);
            val = (a < 0.0) ? fabs(ddiv(T[j], a)) : kPosInf;
            return val > 0.0;
          }, sm);
        }
      }
      if (r < 0 || c < 0) { finish(LPR_INFEASIBLE, src, npiv, 0); return; }  // :324-331
    } else {
      int notopt = 0;  // :345-348
      for (int j = tid; j < C - 1; j += blockDim.x)
        if (lp.is_min ? !(T[j] <= 0.0) : !(T[j] >= 0.0)) notopt++;
      notopt = block_sum_int(notopt, smi);
      if (!notopt) { finish(LPR_OPTIMAL, src, npiv, 0); return; }
      phase = 1;
    }
  }
  if (phase == 1 && r < 0) {
    int notopt = 0;  // :367-373
    for (int j = tid; j < C - 1; j += blockDim.x)
      if (lp.is_min ? !(T[j] <= 0.0) : !(T[j] >= 0.0)) notopt++;
    notopt = block_sum_int(notopt, smi);
    if (!notopt) {
      go_final = true;
    } else if (lp.max_piv >= 0 && npiv >= lp.max_piv) {
      finish(LPR_ITER_LIMIT, src, npiv, 0);
      return;
    } else {
      // PerformPrimalPivot :203-279: Min() of the negative entries (max) / of the positive entries (min)
      const int is_min = lp.is_min;
      c = block_first_min(C - 1, [&](int j, double& val) { val = T[j]; return is_min ? val > 0.0 : val < 0.0; }, sm);
      if (c < 0 || R <= 1) {
        go_final = true;
      } else {
        int notneg = 0, posfin = 0, zero = 0;
        for (int i = 1 + tid; i < R; i += blockDim.x) {
          double a = TAT(T, ld, i, c);
          double th = (a != 0.0) ? ddiv(TAT(T, ld, i, C - 1), a) : kPosInf;
          if (!(th < 0.0)) notneg++;
          if (th > 0.0 && th != kPosInf) posfin++;
          if (th == 0.0) zero++;
        }
        notneg = block_sum_int(notneg, smi);
        posfin = block_sum_int(posfin, smi);
        zero = block_sum_int(zero, smi);
        if (notneg == 0) {
          go_final = true;  // all thetas negative :228-231
        } else if (posfin == 0) {
          if (!zero) {
            go_final = true;
          } else {
            int k = block_first_min(R - 1, [&](int q, double& val) {
              double a = TAT(T, ld, q + 1, c);
              val = (a != 0.0) ? ddiv(TAT(T, ld, q + 1, C - 1), a) : kPosInf;
              return val == 0.0;
            }, sm);
            r = k + 1;
          }
        } else {
          int k = block_first_min(R - 1, [&](int q, double& val) {
            double a = TAT(T, ld, q + 1, c);
            val = (a != 0.0) ? ddiv(TAT(T, ld, q + 1, C - 1), a) : kPosInf;
            return val > 0.0 && val != kPosInf;
          }, sm);
          r = k + 1;
        }
        if (!go_final && (r < 1 || TAT(T, ld, r, c) == 0.0)) go_final = true;
      }
    }
  }
  if (go_final) {
    // :392-400: any negative RHS after the primal phase => drop the last tableau
    int neg = 0;
    for (int i = tid; i < R; i += blockDim.x)
      if (!(TAT(T, ld, i, C - 1) >= 0.0)) neg++;
    neg = block_sum_int(neg, smi);
    if (neg) {
      if (npiv == 0)
        finish(LPR_INFEASIBLE, src, npiv, 0);  // pivotColumns.RemoveAt(-1) throws => branch "failed"
      else
        finish(LPR_OPTIMAL, src ^ 1, npiv - 1, 1);
    } else {
      finish(LPR_OPTIMAL, src, npiv, 0);
    }
    return;
  }
  // stage the pivot: normalised row with -0.0 -> 0.0 (:174-178), pre-update column
  const double piv = TAT(T, ld, r, c);
  for (int j = tid; j < ld; j += blockDim.x) {
    double v = 0.0;
    if (j < C) {
      v = ddiv(TAT(T, ld, r, j), piv);
      if (v == 0.0) v = 0.0;
    }
    lp.prow[j] = v;
  }
  for (int i = tid; i < R; i += blockDim.x) lp.col[i] = TAT(T, ld, i, c);
  if (tid == 0) {
    if (lp.log && npiv < lp.log_cap) {
      lp.log[2 * npiv] = r;
      lp.log[2 * npiv + 1] = c;
    }
    lp.src = src;
    lp.npiv = npiv;
    lp.phase = phase;
    lp.leave = r;
    lp.enter = c;
    lp.do_sweep = 1;
    ctl[kCtlLists + (round & 1) * cap + atomicAdd(&ctl[round & 3], 1)] = lp_index;
  }
}
__global__ void __launch_bounds__(kBBT) k_bb_select(BBLp* lps, int* ctl, int round, int cap) {
  bb_select_body(lps, blockIdx.x, ctl, round, cap);
}

// Sweeps of the LPs that staged a pivot this round.  The (LP, tile) pairs are dealt round-robin to a grid sized for
// the machine, not for the batch: late rounds, in which one or two long chains are still pivoting, get every SM.
__global__ void __launch_bounds__(kSweepThreads) k_bb_sweep(const BBLp* lps, int* ctl, int round, int cap,
                                                            int tiles_max) {
  const int na = ctl[round & 3];
  if (blockIdx.x == 0 && threadIdx.x == 0) ctl[(round + 2) & 3] = 0;
  const int* list = ctl + kCtlLists + (round & 1) * cap;
  constexpr unsigned TILE = kSweepThreads * 8;
  const long long total = (long long)na * tiles_max;
  for (long long g = blockIdx.x; g < total; g += gridDim.x) {
    const BBLp& lp = lps[list[g / tiles_max]];
    const unsigned long long t = (unsigned long long)(g % tiles_max);
    const unsigned ldv = (unsigned)(lp.ld >> 1);
    const unsigned long long n = (unsigned long long)lp.R * ldv;
    if (t * TILE >= n) continue;
    sweep_tile<0, true, false, 8>(reinterpret_cast<const double2*>(lp.buf[lp.src]),
                                  reinterpret_cast<double2*>(lp.buf[lp.src ^ 1]), lp.col,
                                  reinterpret_cast<const double2*>(lp.prow), nullptr, nullptr, n, ldv, lp.leave, t,
                                  0.0, 0u, 0, 0xffffffffu, 0);
  }
}
// number of LPs of the batch still running (host polls it)
__global__ void k_bb_count_running(const BBLp* lps, int n, int* ctl) {
  int c = 0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) c += (lps[i].status == LPR_RUNNING);
  __shared__ int smi[32];
  c = block_sum_int(c, smi);
  if (threadIdx.x == 0) ctl[kCtlRunning] = c;
}

// ---- all rounds of a batch's DoDualSimplex chains in ONE cooperative launch -----------------------------------------
// Round r = select (one CTA per LP: the state machine above) -> grid barrier -> sweeps of the LPs that staged a pivot
// (the (LP, tile) pairs dealt to every CTA, four 256-thread tiles per 1024-thread CTA) -> grid barrier; the loop ends
// in the round in which no LP staged a pivot (then every LP is terminal).  Replaces per round two launches and, every
// few rounds, a running-count kernel with a host synchronisation: for the ~0.3 pivots per node of cfg5 a batch's
// chains were ~13 short launches and two round trips.
__device__ __forceinline__ void bb_grid_barrier(unsigned* bar, unsigned& generation) {
  __syncthreads();
  generation++;
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(bar, 1u);
    const unsigned target = generation * gridDim.x;
    unsigned v;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
    } while (v < target);
    __threadfence();
  }
  __syncthreads();
}
__global__ void __launch_bounds__(kBBT) k_bb_chains(BBLp* lps, int nlp, int* ctl, int cap, int tiles_max, unsigned* bar) {
  unsigned generation = 0;
  constexpr unsigned TILE = kSweepThreads * 8;
  constexpr int kSub = kBBT / kSweepThreads;  // 256-thread tiles a CTA works on side by side
  const int sub = threadIdx.x / kSweepThreads;
  const unsigned tix = threadIdx.x % kSweepThreads;
  for (int round = 0;; round++) {
    for (int lp = blockIdx.x; lp < nlp; lp += gridDim.x) {
      bb_select_body(lps, lp, ctl, round, cap);
      __syncthreads();  // the body's shared scratch is reused by the next LP of this CTA
    }
    bb_grid_barrier(bar, generation);
    const int na = *(volatile int*)&ctl[round & 3];
    if (na == 0) break;  // uniform: every CTA reads the same counter after the barrier
    const int* list = ctl + kCtlLists + (round & 1) * cap;
    const long long total = (long long)na * tiles_max;
    for (long long g = (long long)blockIdx.x * kSub + sub; g < total; g += (long long)gridDim.x * kSub) {
      // the descriptor was updated by the LP's own CTA this round: read it from L2, not from a stale L1 line
      const BBLp* lp = lps + __ldcg(&list[g / tiles_max]);
      const int src = __ldcg(&lp->src), leave = __ldcg(&lp->leave);
      const unsigned long long t = (unsigned long long)(g % tiles_max);
      const unsigned ldv = (unsigned)(__ldcg(&lp->ld) >> 1);
      const unsigned long long n = (unsigned long long)__ldcg(&lp->R) * ldv;
      if (t * TILE >= n) continue;
      const double* b0 = lp->buf[0];  // the buffer / scratch pointers never change during a launch
      const double* b1 = lp->buf[1];
      sweep_tile<0, true, false, 8, true>(reinterpret_cast<const double2*>(src ? b1 : b0),
                                          reinterpret_cast<double2*>(const_cast<double*>(src ? b0 : b1)), lp->col,
                                          reinterpret_cast<const double2*>(lp->prow), nullptr, nullptr, n, ldv, leave, t,
                                          0.0, 0u, 0, 0xffffffffu, 0, tix);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) ctl[(round + 2) & 3] = 0;
    bb_grid_barrier(bar, generation);
  }
}

// the children a batch has just solved, as (tableau, dims) entries for the evaluation kernels: the final tableau of
// LP j sits in buf[src]; a child that did not end OPTIMAL gets R = 0 (its evaluation is never read)
__global__ void k_bb_child_tabs(const BBLp* __restrict__ lps, int n, const double** tabs, int* dims) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const BBLp& lp = lps[j];
  tabs[j] = lp.buf[lp.src];
  dims[3 * j] = lp.status == LPR_OPTIMAL ? lp.R : 0;
  dims[3 * j + 1] = lp.C;
  dims[3 * j + 2] = lp.ld;
}

// ---- node evaluation (:805-857, :892-921) ------------------------------------------------------
// k_bb_eval_x: x_i = Round(RHS of the first row, objective row included, holding a rounded 1 in column i).  Grid
// (column blocks, nodes), CTA = kEvalCols columns x kEvalSegs row segments: a column's rows are scanned by kEvalSegs
// threads (16 independent loads in flight each) and the first hit is the minimum over the segments, so a node costs a
// handful of memory round trips instead of R / 8.
constexpr int kEvalCols = 64, kEvalSegs = 8;
__global__ void __launch_bounds__(kEvalCols* kEvalSegs) k_bb_eval_x(const double* const* tabs, const int* dims,
                                                                    int n_vars, double* xout) {
  __shared__ int first[kEvalSegs][kEvalCols];
  const int node = blockIdx.y;
  const double* T = tabs[node];
  const int R = dims[3 * node], C = dims[3 * node + 1], ld = dims[3 * node + 2];
  const int lc = threadIdx.x % kEvalCols, seg = threadIdx.x / kEvalCols;
  const int i = blockIdx.x * kEvalCols + lc;
  const int per = (R + kEvalSegs - 1) / kEvalSegs;
  const int r0 = seg * per, r1 = min(R, r0 + per);
  int hit = INT_MAX;
  if (i < n_vars) {
    for (int j0 = r0; j0 < r1 && hit == INT_MAX; j0 += 16) {
      double t[16];
      const double* col = T + (size_t)j0 * ld + i;
      if (j0 + 16 <= r1) {  // full block: 16 unpredicated loads issued back to back
#pragma unroll
        for (int q = 0; q < 16; q++) t[q] = __ldg(col + (size_t)q * ld);
      } else {
#pragma unroll
        for (int q = 0; q < 16; q++) t[q] = (j0 + q < r1) ? col[(size_t)q * ld] : 0.0;
      }
#pragma unroll
      for (int q = 15; q >= 0; q--)
        if (j0 + q < r1 && net_round4_is_one(t[q])) hit = j0 + q;
    }
  }
  first[seg][lc] = hit;
  __syncthreads();
  if (seg == 0 && i < n_vars) {
#pragma unroll
    for (int q = 1; q < kEvalSegs; q++) hit = min(hit, first[q][lc]);
    xout[(size_t)node * n_vars + i] = (hit == INT_MAX) ? 0.0 : net_round4(TAT(T, ld, hit, C - 1));
  }
}
// k_bb_eval_pick: integrality count, branching variable (fractional part closest to 0.5, first index), objective
__global__ void __launch_bounds__(kBBT) k_bb_eval_pick(const double* const* tabs, const int* dims, int n_vars,
                                                       BBEval* out, const double* xout) {
  __shared__ MinIdx sm[32];
  __shared__ int smi[32];
  const int node = blockIdx.x;
  const double* T = tabs[node];
  const int C = dims[3 * node + 1], ld = dims[3 * node + 2];
  const double* x = xout + (size_t)node * n_vars;
  int nonint = 0;
  for (int i = threadIdx.x; i < n_vars; i += blockDim.x) {
    double rr = net_round4(x[i]);  // IsInteger :595-599
    if (!(fabs(rr - rint(rr)) <= 1e-6)) nonint++;
  }
  nonint = block_sum_int(nonint, smi);
  int var = block_first_min(n_vars, [&](int i, double& val) {
    double xi = x[i];
    double rr = net_round4(xi);
    if (fabs(rr - rint(rr)) <= 1e-6) return false;
    double fp = __dsub_rn(xi, floor(xi));
    val = fabs(__dsub_rn(fp, 0.5));
    return true;
  }, sm);
  if (threadIdx.x == 0) {
    out[node].z = net_round4(TAT(T, ld, 0, C - 1));  // GetObjective :892-897
    out[node].branch_var = var;
    out[node].branch_val = var >= 0 ? x[var] : 0.0;
    out[node].all_integer = (nonint == 0);
  }
}

// ---- AddConstraint (:694-803) -------------------------------------------------------------------
// k_bb_addc_build: one pass over the parent does both halves of the preparation.  Grid (column blocks, row
// segments, jobs); a thread owns two adjacent columns of one row segment (8 x 128-bit loads in flight):
//   * child rows 0..R-1 = Round(Round(parent)) with a zero column inserted before the RHS, row R = the bound row;
//   * IdentifyBasicVariables :642-662: the reference adds the rounded entries v_i = k_i / 1e4 of a column (ALL rows,
//     objective row and RHS column included) in row order, rounds the sum and tests |Round(sum, 4) - 1| <= 1e-6.
//     With every |entry| < 2e5 and R <= kAddcExactRows the floating-point sum is within (R+1) 2^-53 sum|v_i| < 2.4e-5
//     of K / 1e4, K = sum k_i, so Round recovers K whatever the order of the additions and the test is K == 10000.
//     K is a sum of integers below 2^53: exact and order independent, so the segments combine with atomicAdd (and
//     atomicMin for the first row holding an exact 1).  A job outside those bounds raises its `big` flag and
//     k_bb_addc_elim recomputes its keys with the literal row-order sums (addc_keys_literal).
constexpr int kAddcThreads = 64, kAddcSegs = 4, kAddcExactRows = 1024;
constexpr int kAddcNoRow = 0x7f7f7f7f;  // cudaMemset(0x7f) pattern of the key array: no row holds an exact 1
__device__ __forceinline__ double addc_bound_row(const BBAddc& jb, int j) {  // :727-747
  const int C = jb.C, C2 = C + 1;
  double v = 0.0;
  if (j < C2) {
    if (j < jb.n_vars && j == jb.var) v = net_round4(1.0);
    if (j == C2 - 1) v = net_round4(jb.bound);
    if (j == C - 1) v = (jb.type == 1) ? -1.0 : 1.0;  // overrides a coefficient there
    v = net_round4(v);
  }
  return v;
}
// Round(Round(x, 4), 4) and the integer k = x * 1e4 rounded (0 outside the division-free range)
__device__ __forceinline__ double addc_round(double x, double& k, bool& big) {
  if (fabs(x) < 2e5) {
    k = rint(__dmul_rn(x, 1e4));
    if (k == 0.0) return k;
    const double q0 = __dmul_rn(k, 1e-4);
    return __fma_rn(__fma_rn(-1e4, q0, k), 1e-4, q0);
  }
  big = true;
  k = 0.0;
  return net_round4_div(net_round4_div(x));
}
__global__ void __launch_bounds__(kAddcThreads) k_bb_addc_build(const BBAddc* jobs) {
  const BBAddc& jb = jobs[blockIdx.z];
  const int R = jb.R, C = jb.C, ldc = jb.ldc;
  const int c0 = 2 * (blockIdx.x * kAddcThreads + threadIdx.x);
  if (c0 >= ldc) return;
  const int per = (R + kAddcSegs - 1) / kAddcSegs;
  const int r0 = blockIdx.y * per, r1 = min(R, r0 + per);
  const bool nz = jb.negzero != 0;
  const bool in0 = c0 < C, in1 = c0 + 1 < C;
  const bool interior = c0 + 1 < C - 1;  // both columns keep their place in the child
  const double2* src = reinterpret_cast<const double2*>(jb.parent + c0);
  const size_t sp = (size_t)(jb.ldp >> 1);
  double s0 = 0.0, s1 = 0.0;
  int f0 = INT_MAX, f1 = INT_MAX;
  bool big = R > kAddcExactRows;
  for (int i0 = r0; i0 < r1; i0 += 8) {
    double2 t[8];
#pragma unroll
    for (int q = 0; q < 8; q++)
      if (i0 + q < r1 && in0) t[q] = ld_stream(src + (size_t)(i0 + q) * sp);
#pragma unroll
    for (int q = 0; q < 8; q++) {
      const int i = i0 + q;
      if (i >= r1) continue;
      double v0 = 0.0, v1 = 0.0, k0 = 0.0, k1 = 0.0;
      if (in0) {
        v0 = addc_round(t[q].x, k0, big);
        s0 += k0;
        if (f0 == INT_MAX && v0 == 1.0) f0 = i;
      }
      if (in1) {
        v1 = addc_round(t[q].y, k1, big);
        s1 += k1;
        if (f1 == INT_MAX && v1 == 1.0) f1 = i;
      }
      if (nz) {
        if (v0 == 0.0) v0 = 0.0;
        if (v1 == 0.0) v1 = 0.0;
      }
      double* crow = jb.child + (size_t)i * ldc;
      if (interior) {
        *reinterpret_cast<double2*>(crow + c0) = make_double2(v0, v1);
      } else {  // the pair straddles the inserted column or lies in the padding
#pragma unroll
        for (int h = 0; h < 2; h++) {
          const int c = c0 + h;
          const double v = h ? v1 : v0;
          if (c < C - 1) {
            crow[c] = v;
          } else if (c == C - 1) {
            crow[C - 1] = 0.0;
            crow[C] = v;
          } else if (c > C && c < ldc) {
            crow[c] = 0.0;
          }
        }
      }
    }
  }
  if (blockIdx.y == 0) {
    double* crow = jb.child + (size_t)R * ldc;
    crow[c0] = addc_bound_row(jb, c0);
    if (c0 + 1 < ldc) crow[c0 + 1] = addc_bound_row(jb, c0 + 1);
    if (c0 <= C && C <= c0 + 1) big |= !(fabs(jb.bound) < 2e5);
  }
  if (in0) {
    if (s0 != 0.0) atomicAdd(jb.ksum + c0, s0);
    if (f0 != INT_MAX) atomicMin(jb.key + c0, f0);
  }
  if (in1) {
    if (s1 != 0.0) atomicAdd(jb.ksum + c0 + 1, s1);
    if (f1 != INT_MAX) atomicMin(jb.key + c0 + 1, f1);
  }
  if (big) *jb.big = 1;
}
// the literal IdentifyBasicVariables scan of one column (row-order floating-point sum), for jobs flagged `big`
__device__ int addc_key_literal(const BBAddc& jb, int k) {
  double sum = 0.0;
  int first1 = jb.R;
  for (int i = 0; i < jb.R; i++) {
    const double v = net_round4_div(net_round4_div(jb.parent[(size_t)i * jb.ldp + k]));
    sum = __dadd_rn(sum, v);
    if (first1 == jb.R && v == 1.0) first1 = i;
  }
  sum = net_round4_div(sum);
  return (fabs(sum - 1.0) <= 1e-6) ? first1 : -1;
}
// ordering (:664-685, stable by first-"1" row) and the elimination loop (:752-797); one CTA per job
__global__ void __launch_bounds__(kBBT) k_bb_addc_elim(const BBAddc* jobs) {
  __shared__ MinIdx sm[32];
  __shared__ int sh_nb;
  const BBAddc& jb = jobs[blockIdx.x];
  const int R = jb.R, C = jb.C, C2 = C + 1, ldc = jb.ldc;
  const int tid = threadIdx.x;
  double* child = jb.child;
  double* v = child + (size_t)R * ldc;  // the new constraint row
  // 1) compact the basic columns (ascending column) into order[C..2C) scratch-free: two passes over the
  //    key array with a block-wide running offset; 2) rank them by (key, column) among themselves only
  //    (nb ~ R entries, not C), position = number of basic columns ordered before.
  __shared__ int sh_cnt[kBBT / 32];
  __shared__ int sh_base;
  int* list = jb.order + C;  // second half of the 2*ld scratch row: (col) compacted, keys re-read
  if (tid == 0) {
    sh_nb = 0;
    sh_base = 0;
  }
  // keys: first-"1" row of the columns whose rounded sum is 1 (k_bb_addc_build left K and the first row there)
  {
    const bool literal = *jb.big != 0;
    for (int k = tid; k < C; k += blockDim.x) {
      int key;
      if (literal) {
        key = addc_key_literal(jb, k);
      } else {
        const int f = jb.key[k];
        key = (jb.ksum[k] == 1e4) ? (f == kAddcNoRow ? R : f) : -1;
      }
      jb.key[k] = key;
    }
  }
  __syncthreads();
  for (int k0 = 0; k0 < C; k0 += blockDim.x) {
    const int k = k0 + tid;
    const bool isb = k < C && jb.key[k] >= 0;
    const unsigned bal = __ballot_sync(0xffffffffu, isb);
    const int lane = tid & 31, w = tid >> 5;
    if (lane == 0) sh_cnt[w] = __popc(bal);
    __syncthreads();
    int off = sh_base;
    for (int q = 0; q < w; q++) off += sh_cnt[q];
    if (isb) list[off + __popc(bal & ((1u << lane) - 1u))] = k;
    __syncthreads();
    if (tid == 0) {
      int tot = 0;
      for (int q = 0; q < (int)(blockDim.x >> 5); q++) tot += sh_cnt[q];
      sh_base += tot;
    }
    __syncthreads();
  }
  if (tid == 0) sh_nb = sh_base;
  __syncthreads();
  {
    const int nbl = sh_nb;
    for (int a = tid; a < nbl; a += blockDim.x) {
      const int k = list[a], kk = jb.key[k];
      int pos = 0;
      for (int b = 0; b < nbl; b++) {
        const int q = list[b], kq = jb.key[q];
        if (kq < kk || (kq == kk && q < k)) pos++;
      }
      jb.order[pos] = k;
    }
  }
  __syncthreads();
  const int nb = sh_nb;
  int pos = 0;
  while (true) {
    // next basic column (in order) whose current coefficient in the new row is non-zero
    int q = block_first_min(nb - pos, [&](int t, double& val) {
      const int col = jb.order[pos + t];
      val = (double)t;
      return fabs(net_round4(v[col])) > 1e-6;
    }, sm);
    if (q < 0) break;
    q += pos;
    const int col = jb.order[q];
    const double coef = net_round4(v[col]);
    const int prw = block_first_min(R, [&](int rr, double& val) {
      val = (double)rr;
      return fabs(net_round4(child[(size_t)rr * ldc + col]) - 1.0) <= 1e-6;
    }, sm);
    __syncthreads();
    if (prw >= 0) {
      for (int cc = tid; cc < C2; cc += blockDim.x) {
        const double pv = net_round4(child[(size_t)prw * ldc + cc]);
        const double cv = net_round4(v[cc]);
        const double nv = (jb.type == 1) ? __dsub_rn(pv, __dmul_rn(coef, cv)) : __dsub_rn(cv, __dmul_rn(coef, pv));
        v[cc] = net_round4(nv);
      }
    }
    __syncthreads();
    pos = q + 1;
  }
  bool big = false;
  for (int cc = tid; cc < C2; cc += blockDim.x) {
    double x = net_round4(v[cc]);  // :799
    if (jb.negzero && x == 0.0) x = 0.0;
    big |= !(fabs(x) < 2e5);
    v[cc] = x;
  }
  if (big) *jb.big = 1;
}

}  // namespace lpr

using namespace lpr;

// =============================================================================================
// host side: slab pool, batched node processing
// =============================================================================================
struct BBNode {
  double* slab = nullptr;
  int R = 0, C = 0, depth = 0;
  std::vector<uint8_t> key;  // DFS path: 0 = lower (<= floor) child, 1 = upper (>= ceil) child
  // evaluation (objective, integrality, branching variable), computed on the device right after the node's LP was
  // solved -- in the batch that created it -- so that popping it later needs no kernel and no round trip; roots and
  // imported nodes come without one and are evaluated when popped
  bool has_eval = false;
  BBEval ev = {};
  std::vector<double> x;  // only kept for all-integer nodes (incumbent candidates)
};

static int key_cmp(const std::vector<uint8_t>& a, const std::vector<uint8_t>& b) {
  const size_t n = std::min(a.size(), b.size());
  for (size_t i = 0; i < n; i++)
    if (a[i] != b[i]) return a[i] < b[i] ? -1 : 1;
  if (a.size() == b.size()) return 0;
  return a.size() < b.size() ? -1 : 1;  // an ancestor precedes its descendants (pre-order)
}

struct lpr_bb {
  int device = 0, sms = 148;
  cudaStream_t stream = nullptr;
  int n_vars = 0, R0 = 0, C0 = 0, max_depth = 0, Rmax = 0, ldmax = 0;
  size_t slab_doubles = 0;
  bool prune = false;
  std::vector<double*> free_slabs;
  std::vector<double*> chunks;  // cudaMalloc'ed blocks the slabs are carved from
  std::vector<BBNode> open;     // stack: back() is the next node in DFS order
  bool has_inc = false;
  double inc_z = -INFINITY;
  std::vector<double> inc_x;
  std::vector<uint8_t> inc_key;
  int64_t processed = 0, pivots = 0, depth_overflow = 0;
  double t_eval = 0, t_host = 0, t_addc = 0, t_solve = 0, t_push = 0;  // LPR_BB_PROFILE=1 prints these
  int64_t n_batches = 0, n_children = 0;
  // batch scratch
  int cap = 0;  // max nodes per batch
  BBLp* d_lps = nullptr;
  BBLp* h_lps = nullptr;
  BBAddc* d_jobs = nullptr;
  BBAddc* h_jobs = nullptr;
  BBEval* d_eval = nullptr;
  BBEval* h_eval = nullptr;
  double* d_x = nullptr;
  double* h_x = nullptr;
  const double** d_tabs = nullptr;
  const double** h_tabs = nullptr;
  int* d_dims = nullptr;
  int* h_dims = nullptr;
  double* d_col = nullptr;   // 2*cap x Rmax
  double* d_prow = nullptr;  // 2*cap x ldmax
  int* d_key = nullptr;      // 2*cap x ldmax
  double* d_ksum = nullptr;  // 2*cap x ldmax
  int* d_order = nullptr;    // 2*cap x ldmax
  int* d_ctl = nullptr;    // kCtlLists + 2 * (2*cap) control words of the batched solve
  int* d_dirty = nullptr;  // 2*cap
  int* h_running = nullptr;
  // optional node log (sequential mode)
  int* node_log = nullptr;
  double* node_z = nullptr;
  int64_t node_log_cap = 0;
};

static int bb_take_slab(lpr_bb* h, double** out) {
  if (h->free_slabs.empty()) {
    // grow by chunks of 64 slabs (bounded by what cudaMalloc can give)
    const int per = 64;
    double* blk = nullptr;
    cudaError_t e = cudaMalloc(&blk, sizeof(double) * h->slab_doubles * per);
    int got = per;
    if (e != cudaSuccess) {
      cudaGetLastError();
      got = 4;
      e = cudaMalloc(&blk, sizeof(double) * h->slab_doubles * got);
      if (e != cudaSuccess) return fail(LPR_E_NOMEM, "B&B node pool: out of device memory (%zu open nodes)", h->open.size());
    }
    h->chunks.push_back(blk);
    for (int i = got - 1; i >= 0; i--) h->free_slabs.push_back(blk + (size_t)i * h->slab_doubles);
  }
  *out = h->free_slabs.back();
  h->free_slabs.pop_back();
  return LPR_OK;
}
static void bb_give_slab(lpr_bb* h, double* s) {
  if (s) h->free_slabs.push_back(s);
}

static int bb_alloc_scratch(lpr_bb* h, int cap) {
  h->cap = cap;
  const int nlp = 2 * cap;
#define A_DEV(ptr, type, count) LPR_CUDA(cudaMalloc(&h->ptr, sizeof(type) * (size_t)(count)))
#define A_HOST(ptr, type, count) LPR_CUDA(cudaMallocHost(&h->ptr, sizeof(type) * (size_t)(count)))
  A_DEV(d_lps, BBLp, nlp);
  A_HOST(h_lps, BBLp, nlp);
  A_DEV(d_jobs, BBAddc, nlp);
  A_HOST(h_jobs, BBAddc, nlp);
  A_DEV(d_eval, BBEval, nlp);
  A_HOST(h_eval, BBEval, nlp);
  A_DEV(d_x, double, (size_t)nlp * h->n_vars);
  A_HOST(h_x, double, (size_t)nlp * h->n_vars);
  A_DEV(d_tabs, const double*, nlp);
  A_HOST(h_tabs, const double*, nlp);
  A_DEV(d_dims, int, 3 * nlp);
  A_HOST(h_dims, int, 3 * nlp);
  A_DEV(d_col, double, (size_t)nlp * h->Rmax);
  A_DEV(d_prow, double, (size_t)nlp * h->ldmax);
  A_DEV(d_key, int, (size_t)nlp * h->ldmax);
  A_DEV(d_ksum, double, (size_t)nlp * h->ldmax);
  A_DEV(d_order, int, (size_t)nlp * 2 * h->ldmax);
  A_DEV(d_ctl, int, kCtlLists + 2 * nlp);
  A_DEV(d_dirty, int, nlp);
  A_HOST(h_running, int, 1);
#undef A_DEV
#undef A_HOST
  return LPR_OK;
}

// run the DoDualSimplex loop for nlp LPs whose descriptors are in h_lps[0..nlp).  d_ctl: kCtlLists + 2 * cap device
// ints (cap >= nlp).  skip_negzero: the start tableaux already have -0.0 -> 0.0 applied (BBAddc::negzero);
// round_result: RoundAllTableaux on the finished LPs, skipping those d_dirty (optional) marks as already rounded.
static int bb_solve_batch(cudaStream_t stream, int sms, BBLp* h_lps, BBLp* d_lps, int nlp, int* d_ctl, int cap,
                          int* h_running, size_t max_elems, bool skip_negzero, bool round_result, const int* d_dirty,
                          const std::function<int()>& before_readback = {}) {
  LPR_CUDA(cudaMemcpyAsync(d_lps, h_lps, sizeof(BBLp) * nlp, cudaMemcpyHostToDevice, stream));
  LPR_CUDA(cudaMemsetAsync(d_ctl, 0, sizeof(int) * kCtlLists, stream));
  const int tiles_max = (int)((max_elems / 2 + kSweepThreads * 8 - 1) / (kSweepThreads * 8));
  const int gs = (int)std::max<long long>(1, std::min<long long>((long long)tiles_max * nlp, (long long)sms * 8));
  dim3 ge(std::max(1, std::min(sms, (int)((max_elems + 8191) / 8192))), nlp);
  if (!skip_negzero) {
    k_bb_negzero<<<ge, 256, 0, stream>>>(d_lps);
    LPR_LAUNCH_CHECK();
  }
  // k_bb_chains (all rounds in one cooperative launch) measured the same as the launch-per-round driver on cfg5
  // (74 k against 77 k nodes/s: the chains are bound by their own dependent scans and sweeps, not by launches), so
  // the simpler driver stays the default; LPR_BB_PERSIST=1 selects the cooperative kernel.
  static const int persist = getenv("LPR_BB_PERSIST") ? atoi(getenv("LPR_BB_PERSIST")) : 0;
  if (persist) {
    // one cooperative launch for every round of the batch (k_bb_chains); the barrier counter sits behind the lists
    unsigned* bar = reinterpret_cast<unsigned*>(d_ctl + kCtlBar);
    int nlp_arg = nlp, cap_arg = cap, tiles_arg = tiles_max;
    int* ctl_arg = d_ctl;
    BBLp* lps_arg = d_lps;
    void* args[] = {&lps_arg, &nlp_arg, &ctl_arg, &cap_arg, &tiles_arg, &bar};
    const int grid = std::max(1, std::min(sms, std::max(nlp, 32)));
    LPR_CUDA(cudaLaunchCooperativeKernel((void*)k_bb_chains, dim3(grid), dim3(kBBT), args, 0, stream));
    count_launch();
  } else {
  int round = 0, chunk = 2;
  auto pivot_round = [&](bool count) -> int {
    k_bb_select<<<nlp, kBBT, 0, stream>>>(d_lps, d_ctl, round, cap);
    LPR_LAUNCH_CHECK();
    if (count) {  // the select commits the previous sweep, so the count is exact; it may have staged the next pivot
      k_bb_count_running<<<1, 256, 0, stream>>>(d_lps, nlp, d_ctl);
      LPR_LAUNCH_CHECK();
      LPR_CUDA(cudaMemcpyAsync(h_running, d_ctl + kCtlRunning, sizeof(int), cudaMemcpyDeviceToHost, stream));
    }
    k_bb_sweep<<<gs, kSweepThreads, 0, stream>>>(d_lps, d_ctl, round, cap, tiles_max);
    LPR_LAUNCH_CHECK();
    round++;
    return LPR_OK;
  };
  while (true) {
    int rc;
    for (int q = 0; q < chunk; q++)
      if ((rc = pivot_round(false))) return rc;
    if ((rc = pivot_round(true))) return rc;
    LPR_CUDA(cudaStreamSynchronize(stream));
    if (*h_running == 0) break;
    chunk = std::min(16, chunk * 2);
  }
  }
  if (round_result) {
    k_bb_round<<<ge, 256, 0, stream>>>(d_lps, d_dirty);  // children are stored rounded (:1124, :1187)
    LPR_LAUNCH_CHECK();
  }
  if (before_readback) {  // more work on the finished LPs, enqueued behind them and waited for with the same sync
    const int rc = before_readback();
    if (rc) return rc;
  }
  LPR_CUDA(cudaMemcpyAsync(h_lps, d_lps, sizeof(BBLp) * nlp, cudaMemcpyDeviceToHost, stream));
  LPR_CUDA(cudaStreamSynchronize(stream));
  return LPR_OK;
}

static int bb_run_addc(cudaStream_t stream, BBAddc* h_jobs, BBAddc* d_jobs, int njobs, int max_ldc) {
  LPR_CUDA(cudaMemcpyAsync(d_jobs, h_jobs, sizeof(BBAddc) * njobs, cudaMemcpyHostToDevice, stream));
  dim3 gb((max_ldc / 2 + kAddcThreads - 1) / kAddcThreads, kAddcSegs, njobs);
  k_bb_addc_build<<<gb, kAddcThreads, 0, stream>>>(d_jobs);
  LPR_LAUNCH_CHECK();
  k_bb_addc_elim<<<njobs, kBBT, 0, stream>>>(d_jobs);
  LPR_LAUNCH_CHECK();
  return LPR_OK;
}

static int bb_run_eval(cudaStream_t stream, const double** d_tabs, const int* d_dims, int nb, int n_vars, BBEval* d_eval,
                       double* d_x) {
  dim3 gx((n_vars + kEvalCols - 1) / kEvalCols, nb);
  k_bb_eval_x<<<gx, kEvalCols * kEvalSegs, 0, stream>>>(d_tabs, d_dims, n_vars, d_x);
  LPR_LAUNCH_CHECK();
  k_bb_eval_pick<<<nb, kBBT, 0, stream>>>(d_tabs, d_dims, n_vars, d_eval, d_x);
  LPR_LAUNCH_CHECK();
  return LPR_OK;
}

static inline double now_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// pools created by the multi-GPU driver share a device: each carves 1/n of the slab budget (multi_gpu.cu)
static std::atomic<int> g_prealloc_share{1};
namespace lpr {
void bb_set_prealloc_share(int pools_per_device) { g_prealloc_share.store(std::max(1, pools_per_device)); }
}

static int bb_process(lpr_bb* h, int64_t max_nodes, int batch, int64_t* processed_out, int64_t* pivots_out,
                      bool* hit_limit, double max_seconds = 0.0) {
  int rc = select_device(h->device);
  if (rc) return rc;
  int64_t done = 0, piv = 0;
  if (hit_limit) *hit_limit = false;
  const double t_start = now_s();
  while (!h->open.empty()) {
    if (max_nodes >= 0 && done >= max_nodes) {
      if (hit_limit) *hit_limit = true;
      break;
    }
    if (max_seconds > 0.0 && done > 0 && now_s() - t_start >= max_seconds) break;
    int nb = (int)std::min<int64_t>(std::min<int64_t>(batch, h->cap), (int64_t)h->open.size());
    if (max_nodes >= 0) nb = (int)std::min<int64_t>(nb, max_nodes - done);
    double tp0 = now_s();
    h->n_batches++;
    std::vector<BBNode> cur;
    for (int i = 0; i < nb; i++) {  // cur[0] is the DFS-next node
      cur.push_back(std::move(h->open.back()));
      h->open.pop_back();
    }
    // nodes created by this pool carry their evaluation; roots and imported nodes are evaluated here
    int ne = 0;
    std::vector<int> need;
    for (int i = 0; i < nb; i++)
      if (!cur[i].has_eval) need.push_back(i);
    ne = (int)need.size();
    if (ne > 0) {
      for (int q = 0; q < ne; q++) {
        const BBNode& nd = cur[need[q]];
        h->h_tabs[q] = nd.slab;
        h->h_dims[3 * q] = nd.R;
        h->h_dims[3 * q + 1] = nd.C;
        h->h_dims[3 * q + 2] = h->ldmax;
      }
      LPR_CUDA(cudaMemcpyAsync(h->d_tabs, h->h_tabs, sizeof(double*) * ne, cudaMemcpyHostToDevice, h->stream));
      LPR_CUDA(cudaMemcpyAsync(h->d_dims, h->h_dims, sizeof(int) * 3 * ne, cudaMemcpyHostToDevice, h->stream));
      if ((rc = bb_run_eval(h->stream, h->d_tabs, h->d_dims, ne, h->n_vars, h->d_eval, h->d_x))) return rc;
      LPR_CUDA(cudaMemcpyAsync(h->h_eval, h->d_eval, sizeof(BBEval) * ne, cudaMemcpyDeviceToHost, h->stream));
      LPR_CUDA(cudaStreamSynchronize(h->stream));
      for (int q = 0; q < ne; q++) {
        BBNode& nd = cur[need[q]];
        nd.ev = h->h_eval[q];
        nd.has_eval = true;
        if (nd.ev.all_integer) {  // x is only needed for incumbent candidates: fetched on demand
          nd.x.resize(h->n_vars);
          LPR_CUDA(cudaMemcpy(nd.x.data(), h->d_x + (size_t)q * h->n_vars, sizeof(double) * h->n_vars, cudaMemcpyDeviceToHost));
        }
      }
    }

    double tp1 = now_s();
    h->t_eval += tp1 - tp0;
    // host bookkeeping in DFS (pop) order: prune, incumbent, branch (:1060-1076)
    struct Job { int node; int side; };
    std::vector<Job> jobs;
    for (int i = 0; i < nb; i++) {
      const BBEval& ev = cur[i].ev;
      const int64_t q = h->processed++;
      done++;
      bool pruned = false;
      if (h->prune && h->has_inc) {  // ShouldPrunebranch :985-1004 (+ DFS-key tie rule)
        if (ev.z < h->inc_z || (ev.z == h->inc_z && key_cmp(cur[i].key, h->inc_key) > 0)) pruned = true;
      }
      if (!pruned && ev.all_integer) {  // UpdateOptimalSolution :935-983
        if (!h->has_inc || ev.z > h->inc_z || (ev.z == h->inc_z && key_cmp(cur[i].key, h->inc_key) < 0)) {
          h->has_inc = true;
          h->inc_z = ev.z;
          h->inc_x = cur[i].x;
          h->inc_key = cur[i].key;
        }
      }
      if (h->node_log && q < h->node_log_cap) {
        h->node_log[4 * q + 0] = cur[i].depth;
        h->node_log[4 * q + 1] = pruned ? -1 : ev.branch_var;
        h->node_log[4 * q + 2] = pruned ? 0 : ev.all_integer;
        h->node_log[4 * q + 3] = pruned ? 1 : 0;
        if (h->node_z) h->node_z[q] = ev.z;
      }
      if (pruned || ev.branch_var < 0) continue;
      if (cur[i].depth + 1 > h->max_depth) {
        h->depth_overflow++;
        continue;
      }
      jobs.push_back({i, 0});
      jobs.push_back({i, 1});
    }
    // children: AddConstraint + DoDualSimplex, batched
    const int nj = (int)jobs.size();
    bool evaluated_children = false;
    std::vector<double*> slabA(nj, nullptr), slabB(nj, nullptr);
    if (nj > 0) {
      size_t max_elems = 0;
      for (int j = 0; j < nj; j++) {
        if ((rc = bb_take_slab(h, &slabA[j]))) return rc;
        if ((rc = bb_take_slab(h, &slabB[j]))) return rc;
        const BBNode& nd = cur[jobs[j].node];
        const BBEval& ev = nd.ev;
        BBAddc& jb = h->h_jobs[j];
        jb.parent = nd.slab;
        jb.child = slabA[j];
        jb.key = h->d_key + (size_t)j * h->ldmax;
        jb.order = h->d_order + (size_t)j * 2 * h->ldmax;
        jb.ksum = h->d_ksum + (size_t)j * h->ldmax;
        jb.big = h->d_dirty + j;
        jb.negzero = 1;
        jb.R = nd.R;
        jb.C = nd.C;
        jb.ldp = h->ldmax;
        jb.ldc = h->ldmax;
        jb.n_vars = h->n_vars;
        jb.var = ev.branch_var;
        jb.type = jobs[j].side;
        // (int)Math.Floor / (int)Math.Ceiling :870-871
        jb.bound = jobs[j].side == 0 ? (double)(int)std::floor(ev.branch_val) : (double)(int)std::ceil(ev.branch_val);
        max_elems = std::max(max_elems, (size_t)(nd.R + 1) * h->ldmax);
        BBLp& lp = h->h_lps[j];
        memset(&lp, 0, sizeof lp);
        lp.buf[0] = slabA[j];
        lp.buf[1] = slabB[j];
        lp.col = h->d_col + (size_t)j * h->Rmax;
        lp.prow = h->d_prow + (size_t)j * h->ldmax;
        lp.log = nullptr;
        lp.log_cap = 0;
        lp.R = nd.R + 1;
        lp.C = nd.C + 1;
        lp.ld = h->ldmax;
        lp.status = LPR_RUNNING;
        lp.max_piv = -1;
      }
      double tp2 = now_s();
      h->t_host += tp2 - tp1;
      LPR_CUDA(cudaMemsetAsync(h->d_dirty, 0, sizeof(int) * nj, h->stream));
      LPR_CUDA(cudaMemsetAsync(h->d_ksum, 0, sizeof(double) * (size_t)nj * h->ldmax, h->stream));
      LPR_CUDA(cudaMemsetAsync(h->d_key, 0x7f, sizeof(int) * (size_t)nj * h->ldmax, h->stream));
      if ((rc = bb_run_addc(h->stream, h->h_jobs, h->d_jobs, nj, h->ldmax))) return rc;
      if (getenv("LPR_BB_PROFILE")) cudaStreamSynchronize(h->stream);
      double tp3 = now_s();
      h->t_addc += tp3 - tp2;
      auto eval_children = [&]() -> int {
        k_bb_child_tabs<<<(nj + 127) / 128, 128, 0, h->stream>>>(h->d_lps, nj, h->d_tabs, h->d_dims);
        LPR_LAUNCH_CHECK();
        int rc2 = bb_run_eval(h->stream, h->d_tabs, h->d_dims, nj, h->n_vars, h->d_eval, h->d_x);
        if (rc2) return rc2;
        LPR_CUDA(cudaMemcpyAsync(h->h_eval, h->d_eval, sizeof(BBEval) * nj, cudaMemcpyDeviceToHost, h->stream));
        return LPR_OK;
      };
      // Evaluating children here saves the evaluation round trip of the batch that pops them, but a child that is
      // still open when the search stops was evaluated for nothing: on trees that close it is a gain, on the
      // time-sliced cfg5 run (which leaves as many nodes open as it processes) it doubles the evaluation work.
      // Default: evaluate at pop; LPR_BB_EVAL_AT_CREATE=1 switches.
      static const bool eval_at_create = getenv("LPR_BB_EVAL_AT_CREATE") && atoi(getenv("LPR_BB_EVAL_AT_CREATE")) != 0;
      if ((rc = bb_solve_batch(h->stream, h->sms, h->h_lps, h->d_lps, nj, h->d_ctl, 2 * h->cap, h->h_running, max_elems,
                               true, true, h->d_dirty, eval_at_create ? std::function<int()>(eval_children)
                                                                      : std::function<int()>())))
        return rc;
      evaluated_children = eval_at_create;
      h->t_solve += now_s() - tp3;
      h->n_children += nj;
    }
    double tp4 = now_s();
    // push feasible children so that the DFS order is preserved: children of cur[0] end on top,
    // lower branch above upper branch (:1210-1213)
    for (int i = nb - 1; i >= 0; i--) {
      for (int side = 1; side >= 0; side--) {
        for (int j = 0; j < nj; j++) {
          if (jobs[j].node != i || jobs[j].side != side) continue;
          const BBLp& lp = h->h_lps[j];
          piv += lp.npiv;
          if (lp.status == LPR_OPTIMAL) {
            BBNode ch;
            ch.slab = lp.src ? slabB[j] : slabA[j];
            bb_give_slab(h, lp.src ? slabA[j] : slabB[j]);
            ch.R = lp.R;
            ch.C = lp.C;
            ch.depth = cur[i].depth + 1;
            ch.key = cur[i].key;
            ch.key.push_back((uint8_t)side);
            if (evaluated_children) {
              ch.ev = h->h_eval[j];
              ch.has_eval = true;
            }
            if (evaluated_children && ch.ev.all_integer) {  // an incumbent candidate: keep its x (rare, on demand)
              ch.x.resize(h->n_vars);
              LPR_CUDA(cudaMemcpy(ch.x.data(), h->d_x + (size_t)j * h->n_vars, sizeof(double) * h->n_vars,
                                  cudaMemcpyDeviceToHost));
            }
            h->open.push_back(std::move(ch));
          } else {
            bb_give_slab(h, slabA[j]);
            bb_give_slab(h, slabB[j]);
          }
        }
      }
    }
    for (int i = 0; i < nb; i++) bb_give_slab(h, cur[i].slab);
    h->t_push += now_s() - tp4;
  }
  h->pivots += piv;
  if (processed_out) *processed_out = done;
  if (pivots_out) *pivots_out = piv;
  return LPR_OK;
}

extern "C" {

int lpr_bb_destroy(lpr_bb* h) {
  if (!h) return LPR_OK;
  cudaSetDevice(h->device);
  if (getenv("LPR_BB_PROFILE"))
    fprintf(stderr, "[lpr_bb] nodes=%lld batches=%lld children=%lld pivots=%lld | eval %.3fs host %.3fs addc %.3fs solve %.3fs push %.3fs\n",
            (long long)h->processed, (long long)h->n_batches, (long long)h->n_children, (long long)h->pivots, h->t_eval,
            h->t_host, h->t_addc, h->t_solve, h->t_push);
  if (h->stream) cudaStreamSynchronize(h->stream);
  for (double* c : h->chunks) cudaFree(c);
  cudaFree(h->d_lps); cudaFree(h->d_jobs); cudaFree(h->d_eval); cudaFree(h->d_x); cudaFree(h->d_tabs);
  cudaFree(h->d_dims); cudaFree(h->d_col); cudaFree(h->d_prow); cudaFree(h->d_key); cudaFree(h->d_order);
  cudaFree(h->d_ctl); cudaFree(h->d_dirty); cudaFree(h->d_ksum);
  if (h->h_lps) cudaFreeHost(h->h_lps);
  if (h->h_jobs) cudaFreeHost(h->h_jobs);
  if (h->h_eval) cudaFreeHost(h->h_eval);
  if (h->h_x) cudaFreeHost(h->h_x);
  if (h->h_tabs) cudaFreeHost(h->h_tabs);
  if (h->h_dims) cudaFreeHost(h->h_dims);
  if (h->h_running) cudaFreeHost(h->h_running);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return LPR_OK;
}

static int bb_create_empty(int device, int rows, int cols, int n_vars, int enable_pruning, lpr_bb** out) {
  if (!out) return fail(LPR_E_BADARG, "out is null");
  *out = nullptr;
  if (rows < 1 || cols < 2 || n_vars < 1 || n_vars > cols - 1)
    return fail(LPR_E_BADARG, "bad B&B shape rows=%d cols=%d n_vars=%d", rows, cols, n_vars);
  int rc = select_device(device);
  if (rc) return rc;
  lpr_bb* h = new (std::nothrow) lpr_bb();
  if (!h) return fail(LPR_E_NOMEM, "host allocation failed");
  h->device = device;
  h->sms = sm_count(device);
  h->n_vars = n_vars;
  h->R0 = rows;
  h->C0 = cols;
  const char* md = getenv("LPR_BB_MAX_DEPTH");
  h->max_depth = md ? std::max(1, atoi(md)) : 128;
  h->Rmax = rows + h->max_depth;
  h->ldmax = round_up(cols + h->max_depth, 16);
  h->slab_doubles = (size_t)h->Rmax * h->ldmax;
  h->prune = enable_pruning != 0;
  if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) {
    delete h;
    return fail(LPR_E_CUDA, "stream creation failed");
  }
  {  // pre-carve slabs so that steady-state node processing never calls cudaMalloc
    const char* pm = getenv("LPR_BB_PREALLOC_MB");
    const int share = std::max(1, g_prealloc_share.load());
    size_t want = ((size_t)(pm ? std::max(0, atoi(pm)) : 1024) << 20) / share;
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) want = std::min(want, (free_b - free_b / 4) / share);
    size_t nsl = want / (sizeof(double) * h->slab_doubles);
    while (nsl > 0) {
      const size_t take = std::min<size_t>(nsl, 512);
      double* blk = nullptr;
      if (cudaMalloc(&blk, sizeof(double) * h->slab_doubles * take) != cudaSuccess) {
        cudaGetLastError();
        break;
      }
      h->chunks.push_back(blk);
      for (size_t i = take; i-- > 0;) h->free_slabs.push_back(blk + i * h->slab_doubles);
      nsl -= take;
    }
  }
  const char* bc = getenv("LPR_BB_BATCH");
  rc = bb_alloc_scratch(h, bc ? std::max(1, atoi(bc)) : 64);
  if (rc) {
    lpr_bb_destroy(h);
    return rc;
  }
  *out = h;
  return LPR_OK;
}

static int bb_push_host_node(lpr_bb* h, int R, int C, int depth, const uint8_t* key, int key_len, const double* dense,
                             bool round) {
  if (R > h->Rmax || C > h->ldmax) return fail(LPR_E_CAPACITY, "node %dx%d exceeds the slab size", R, C);
  BBNode nd;
  int rc = bb_take_slab(h, &nd.slab);
  if (rc) return rc;
  // `dense` may live on the host or on the device (cudaMemcpyDefault); padding columns are zeroed
  if (h->ldmax > C)
    LPR_CUDA(cudaMemset2DAsync(nd.slab + C, sizeof(double) * h->ldmax, 0, sizeof(double) * (h->ldmax - C), R, h->stream));
  LPR_CUDA(cudaMemcpy2DAsync(nd.slab, sizeof(double) * h->ldmax, dense, sizeof(double) * C, sizeof(double) * C, R,
                             cudaMemcpyDefault, h->stream));
  if (round) {
    k_round4<<<h->sms * 2, 256, 0, h->stream>>>(nd.slab, (size_t)R * h->ldmax);  // :1021
    LPR_LAUNCH_CHECK();
  }
  nd.R = R;
  nd.C = C;
  nd.depth = depth;
  nd.key.assign(key, key + key_len);
  h->open.push_back(std::move(nd));
  return LPR_OK;
}

int lpr_bb_create(int device, int rows, int cols, const double* root_tableau, int n_vars, int enable_pruning,
                  lpr_bb** out) {
  lpr_bb* h = nullptr;
  int rc = bb_create_empty(device, rows, cols, n_vars, enable_pruning, &h);
  if (rc) return rc;
  if (root_tableau) {
    rc = bb_push_host_node(h, rows, cols, 0, nullptr, 0, root_tableau, true);
    if (rc == LPR_OK && cudaStreamSynchronize(h->stream) != cudaSuccess) rc = fail(LPR_E_CUDA, "root upload failed");
    if (rc) {
      lpr_bb_destroy(h);
      return rc;
    }
  }
  *out = h;
  return LPR_OK;
}

int lpr_bb_open_count(lpr_bb* h, int64_t* n) {
  if (!h || !n) return fail(LPR_E_BADARG, "null argument");
  *n = (int64_t)h->open.size();
  return LPR_OK;
}

int lpr_bb_stats(lpr_bb* h, int64_t* processed, int64_t* pivots, int64_t* depth_overflow, int* max_depth) {
  if (!h) return fail(LPR_E_BADARG, "null handle");
  if (processed) *processed = h->processed;
  if (pivots) *pivots = h->pivots;
  if (depth_overflow) *depth_overflow = h->depth_overflow;
  if (max_depth) *max_depth = h->max_depth;
  return LPR_OK;
}

int lpr_bb_run(lpr_bb* h, int64_t max_nodes, int64_t* processed, int64_t* pivots) {
  if (!h) return fail(LPR_E_BADARG, "null handle");
  return bb_process(h, max_nodes, h->cap, processed, pivots, nullptr);
}
int lpr_bb_keep_stride(lpr_bb* h, int offset, int stride) {
  if (!h || stride < 1 || offset < 0 || offset >= stride) return fail(LPR_E_BADARG, "bad keep_stride arguments");
  std::vector<BBNode> kept;
  for (size_t i = 0; i < h->open.size(); i++) {
    if ((int)(i % (size_t)stride) == offset)
      kept.push_back(std::move(h->open[i]));
    else
      bb_give_slab(h, h->open[i].slab);
  }
  h->open.swap(kept);
  return LPR_OK;
}
int lpr_bb_run_timed(lpr_bb* h, int64_t max_nodes, double max_seconds, int64_t* processed, int64_t* pivots) {
  if (!h) return fail(LPR_E_BADARG, "null handle");
  return bb_process(h, max_nodes, h->cap, processed, pivots, nullptr, max_seconds);
}

int lpr_bb_get_incumbent(lpr_bb* h, int* has, double* z, double* x, int* key, int* key_len) {
  if (!h || !has) return fail(LPR_E_BADARG, "null argument");
  *has = h->has_inc ? 1 : 0;
  if (z) *z = h->inc_z;
  if (x && h->has_inc) std::copy(h->inc_x.begin(), h->inc_x.end(), x);
  if (key_len) {
    const int capk = *key_len;
    *key_len = (int)h->inc_key.size();
    if (key)
      for (int i = 0; i < std::min<int>(capk, (int)h->inc_key.size()); i++) key[i] = h->inc_key[i];
  }
  return LPR_OK;
}

int lpr_bb_set_incumbent(lpr_bb* h, double z, const double* x, const int* key, int key_len) {
  if (!h || !x || key_len < 0 || (key_len > 0 && !key)) return fail(LPR_E_BADARG, "bad incumbent");
  std::vector<uint8_t> k(key_len);
  for (int i = 0; i < key_len; i++) k[i] = (uint8_t)key[i];
  if (!h->has_inc || z > h->inc_z || (z == h->inc_z && key_cmp(k, h->inc_key) < 0)) {
    h->has_inc = true;
    h->inc_z = z;
    h->inc_x.assign(x, x + h->n_vars);
    h->inc_key = k;
  }
  return LPR_OK;
}

// node record: int32 R, C, depth, key_len; key bytes padded to 8; R*C doubles
int lpr_bb_export_nodes(lpr_bb* h, int max_nodes, void* buf, int64_t buf_cap, int64_t* bytes, int* n_exported) {
  if (!h || !buf || !bytes || !n_exported) return fail(LPR_E_BADARG, "null argument");
  int rc = select_device(h->device);
  if (rc) return rc;
  char* p = (char*)buf;
  int64_t used = 0;
  int n = 0;
  // shallowest nodes sit at the bottom of the stack: give those away (largest subtrees); the copies go through the
  // handle's stream and are waited for once, at the end (the caller hands `buf` to NCCL on another stream)
  size_t first = 0;
  std::vector<std::vector<char>> heads;
  while (n < max_nodes && first < h->open.size()) {
    BBNode& nd = h->open[first];
    const int64_t kl = (int64_t)nd.key.size(), kpad = (kl + 7) / 8 * 8;
    const int64_t need = 16 + kpad + (int64_t)sizeof(double) * nd.R * nd.C;
    if (used + need > buf_cap) break;
    heads.emplace_back(16 + kpad, 0);
    std::vector<char>& head = heads.back();
    int32_t hdr[4] = {nd.R, nd.C, nd.depth, (int32_t)kl};
    memcpy(head.data(), hdr, 16);
    if (kl) memcpy(head.data() + 16, nd.key.data(), kl);
    // cudaMemcpyDefault: buf may be pageable/pinned host memory or device memory (NCCL staging buffer)
    LPR_CUDA(cudaMemcpyAsync(p + used, head.data(), head.size(), cudaMemcpyDefault, h->stream));
    LPR_CUDA(cudaMemcpy2DAsync(p + used + 16 + kpad, sizeof(double) * nd.C, nd.slab, sizeof(double) * h->ldmax,
                               sizeof(double) * nd.C, nd.R, cudaMemcpyDefault, h->stream));
    used += need;
    first++;
    n++;
  }
  LPR_CUDA(cudaStreamSynchronize(h->stream));
  for (size_t i = 0; i < first; i++) bb_give_slab(h, h->open[i].slab);
  h->open.erase(h->open.begin(), h->open.begin() + (ptrdiff_t)first);  // one erase, not one per node
  *bytes = used;
  *n_exported = n;
  return LPR_OK;
}

int lpr_bb_import_nodes(lpr_bb* h, const void* buf, int64_t bytes) {
  if (!h || (!buf && bytes > 0)) return fail(LPR_E_BADARG, "null argument");
  int rc = select_device(h->device);
  if (rc) return rc;
  const char* p = (const char*)buf;
  int64_t off = 0;
  std::vector<uint8_t> keybuf;
  while (off + 16 <= bytes) {
    int32_t hdr[4];
    LPR_CUDA(cudaMemcpy(hdr, p + off, 16, cudaMemcpyDefault));
    const int64_t kl = hdr[3], kpad = (kl + 7) / 8 * 8;
    const int64_t need = 16 + kpad + (int64_t)sizeof(double) * hdr[0] * hdr[1];
    if (kl < 0 || off + need > bytes) return fail(LPR_E_BADARG, "truncated node record");
    keybuf.resize((size_t)kpad + 8);
    if (kpad) LPR_CUDA(cudaMemcpy(keybuf.data(), p + off + 16, kpad, cudaMemcpyDefault));
    rc = bb_push_host_node(h, hdr[0], hdr[1], hdr[2], keybuf.data(), (int)kl, (const double*)(p + off + 16 + kpad), false);
    if (rc) return rc;
    off += need;
  }
  LPR_CUDA(cudaStreamSynchronize(h->stream));
  // keep the stack sorted so that back() is the DFS-first open node
  std::stable_sort(h->open.begin(), h->open.end(),
                   [](const BBNode& a, const BBNode& b) { return key_cmp(a.key, b.key) > 0; });
  return LPR_OK;
}

int lpr_bb_solve(int device, int rows, int cols, const double* final_tableau, int n_vars, int enable_pruning,
                 int64_t max_nodes, double* x, double* z, int* has_solution, int64_t* nodes, int64_t* pivots,
                 int* node_log, double* node_z, int64_t node_log_cap, int* status) {
  if (!final_tableau) return fail(LPR_E_BADARG, "null tableau");
  lpr_bb* h = nullptr;
  int rc = lpr_bb_create(device, rows, cols, final_tableau, n_vars, enable_pruning, &h);
  if (rc) return rc;
  h->node_log = node_log;
  h->node_z = node_z;
  h->node_log_cap = node_log ? node_log_cap : 0;
  bool limit = false;
  int64_t done = 0, piv = 0;
  // sequential reference order: one node per round
  rc = bb_process(h, max_nodes, 1, &done, &piv, &limit);
  if (rc == LPR_OK) {
    if (has_solution) *has_solution = h->has_inc ? 1 : 0;
    if (z) *z = h->has_inc ? h->inc_z : -INFINITY;
    if (x)
      for (int i = 0; i < n_vars; i++) x[i] = h->has_inc ? h->inc_x[i] : 0.0;
    if (nodes) *nodes = done;
    if (pivots) *pivots = piv;
    // a subtree cut at the slab depth headroom means the incumbent is not proven optimal: say so (LPR_DEPTH_LIMIT)
    if (status) *status = h->depth_overflow > 0 ? LPR_DEPTH_LIMIT : (limit ? LPR_NODE_LIMIT : LPR_OPTIMAL);
  }
  lpr_bb_destroy(h);
  return rc;
}

// ---- building blocks on a single lpr_tab (parity tests, C# shim of DualSimplexSolverBB) --------------
int lpr_tab_round4(lpr_tab* h) {
  if (!h) return fail(LPR_E_BADARG, "null handle");
  int rc = select_device(h->device);
  if (rc) return rc;
  k_round4<<<h->sms * 2, 256, 0, h->stream>>>(h->T, (size_t)h->R * h->ld);
  LPR_LAUNCH_CHECK();
  LPR_CUDA(cudaStreamSynchronize(h->stream));
  return LPR_OK;
}

int lpr_tab_create_bb(int device, int n, int m, const double* objective, const double* cons, int stride,
                      const int* len, int row_cap, int col_cap, lpr_tab** out) {
  if (!out) return fail(LPR_E_BADARG, "out is null");
  *out = nullptr;
  if (n < 1 || m < 0 || !objective || (m > 0 && (!cons || !len)) || stride < 2)
    return fail(LPR_E_BADARG, "bad model (n=%d m=%d)", n, m);
  for (int i = 0; i < m; i++)
    if (len[i] < 2 || len[i] > stride || len[i] - 2 > n + m)
      return fail(LPR_E_BADARG, "constraint row %d has %d entries (need coefficients, rhs, type; at most %d)", i, len[i],
                  std::min(stride, n + m + 2));
  lpr_tab* h = nullptr;
  int rc = tab_alloc(device, m + 1, n + m + 1, row_cap, col_cap, &h);
  if (rc) return rc;
  double *d_obj = nullptr, *d_cons = nullptr;
  int* d_len = nullptr;
  cudaError_t e = cudaMalloc(&d_obj, sizeof(double) * n);
  if (e == cudaSuccess) e = cudaMalloc(&d_cons, sizeof(double) * std::max<size_t>(1, (size_t)m * stride));
  if (e == cudaSuccess) e = cudaMalloc(&d_len, sizeof(int) * std::max(1, m));
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_obj, objective, sizeof(double) * n, cudaMemcpyHostToDevice, h->stream);
  if (e == cudaSuccess && m > 0)
    e = cudaMemcpyAsync(d_cons, cons, sizeof(double) * (size_t)m * stride, cudaMemcpyHostToDevice, h->stream);
  if (e == cudaSuccess && m > 0) e = cudaMemcpyAsync(d_len, len, sizeof(int) * m, cudaMemcpyHostToDevice, h->stream);
  if (e == cudaSuccess) {
    const size_t total = (size_t)(m + 1) * h->ld;
    k_bb_formulate<<<(int)std::max<size_t>(1, std::min<size_t>((size_t)h->sms * 8, (total + 255) / 256)), 256, 0,
                     h->stream>>>(h->T, h->ld, n, m, d_obj, d_cons, stride, d_len);
    count_launch();
    e = cudaStreamSynchronize(h->stream);
  }
  cudaFree(d_obj);
  cudaFree(d_cons);
  cudaFree(d_len);
  if (e != cudaSuccess) {
    lpr_tab_destroy(h);
    return fail(LPR_E_CUDA, "create_bb: %s", cudaGetErrorString(e));
  }
  *out = h;
  return LPR_OK;
}

int lpr_tab_bb_node_solve(lpr_tab* h, int64_t max_pivots, int* status, int64_t* n_pivots, int* pivot_log,
                          int64_t log_cap) {
  return lpr_tab_bb_node_solve_ex(h, 0, max_pivots, status, n_pivots, pivot_log, log_cap);
}

int lpr_tab_bb_node_solve_ex(lpr_tab* h, int is_minimization, int64_t max_pivots, int* status, int64_t* n_pivots,
                             int* pivot_log, int64_t log_cap) {
  if (!h) return fail(LPR_E_BADARG, "null handle");
  int rc = select_device(h->device);
  if (rc) return rc;
  if ((rc = tab_ensure_T2(h))) return rc;
  if (pivot_log && log_cap > 0 && (rc = tab_ensure_log(h, std::min<long long>(log_cap, 1 << 22)))) return rc;
  BBLp* d_lp = nullptr;
  BBLp* h_lp = nullptr;
  int *d_ctl = nullptr, *h_run = nullptr;
  LPR_CUDA(cudaMalloc(&d_lp, sizeof(BBLp)));
  LPR_CUDA(cudaMallocHost(&h_lp, sizeof(BBLp)));
  LPR_CUDA(cudaMalloc(&d_ctl, sizeof(int) * (kCtlLists + 2)));
  LPR_CUDA(cudaMallocHost(&h_run, sizeof(int)));
  memset(h_lp, 0, sizeof(BBLp));
  h_lp->buf[0] = h->T;
  h_lp->buf[1] = h->T2;
  h_lp->col = h->col[0];
  h_lp->prow = h->prow;
  h_lp->log = (pivot_log && log_cap > 0) ? h->log : nullptr;
  h_lp->log_cap = (pivot_log && log_cap > 0) ? (int)std::min<long long>(h->log_cap, log_cap) : 0;
  h_lp->R = h->R;
  h_lp->C = h->C;
  h_lp->ld = h->ld;
  h_lp->status = LPR_RUNNING;
  h_lp->max_piv = max_pivots;
  h_lp->is_min = is_minimization ? 1 : 0;
  // note: the building block returns the un-rounded final tableau (rounding is the caller's step :1124)
  rc = bb_solve_batch(h->stream, h->sms, h_lp, d_lp, 1, d_ctl, 1, h_run, (size_t)h->R * h->ld, false, false, nullptr);
  if (rc) {
    cudaFree(d_lp); cudaFreeHost(h_lp); cudaFree(d_ctl); cudaFreeHost(h_run);
    return rc;
  }
  if (h_lp->src == 1) std::swap(h->T, h->T2);  // the result lives in the other ping-pong buffer
  if (status) *status = h_lp->status;
  if (n_pivots) *n_pivots = h_lp->npiv;
  if (pivot_log && log_cap > 0 && h_lp->npiv > 0) {
    // the log keeps the dropped pivot too (the reference removes it from pivotRows :398-399)
    long long cnt = std::min<long long>(std::min<long long>(h_lp->npiv, log_cap), h->log_cap);
    LPR_CUDA(cudaMemcpy(pivot_log, h->log, sizeof(int) * 2 * (size_t)cnt, cudaMemcpyDeviceToHost));
  }
  cudaFree(d_lp);
  cudaFreeHost(h_lp);
  cudaFree(d_ctl);
  cudaFreeHost(h_run);
  return LPR_OK;
}

int lpr_tab_bb_add_constraint(lpr_tab* parent, int n_vars, int var, double bound, int type, lpr_tab** child) {
  if (!parent || !child || n_vars < 0 || n_vars > parent->C - 1 || var < 0 || var >= std::max(1, n_vars))
    return fail(LPR_E_BADARG, "bad AddConstraint arguments");
  int rc = select_device(parent->device);
  if (rc) return rc;
  lpr_tab* ch = nullptr;
  rc = tab_alloc(parent->device, parent->R + 1, parent->C + 1, 0, 0, &ch);
  if (rc) return rc;
  BBAddc jb;
  // scratch in one block: [ksum: C doubles | big: 2 ints | key: C ints | order: 2C ints | job descriptor]
  const size_t Cp = (size_t)round_up(parent->C, 2);
  const size_t off_big = sizeof(double) * Cp, off_key = off_big + 2 * sizeof(int), off_order = off_key + sizeof(int) * Cp,
               off_job = (off_order + sizeof(int) * 2 * Cp + 15) / 16 * 16;
  char* d_scr = nullptr;
  if (cudaMalloc(&d_scr, off_job + sizeof(BBAddc)) != cudaSuccess) {
    cudaGetLastError();
    lpr_tab_destroy(ch);
    return fail(LPR_E_NOMEM, "AddConstraint scratch allocation failed");
  }
  jb.parent = parent->T;
  jb.child = ch->T;
  jb.ksum = reinterpret_cast<double*>(d_scr);
  jb.big = reinterpret_cast<int*>(d_scr + off_big);
  jb.key = reinterpret_cast<int*>(d_scr + off_key);
  jb.order = reinterpret_cast<int*>(d_scr + off_order);
  jb.negzero = 0;  // AddConstraint proper: the -0.0 clean-up belongs to DoDualSimplex
  jb.R = parent->R;
  jb.C = parent->C;
  jb.ldp = parent->ld;
  jb.ldc = ch->ld;
  jb.n_vars = n_vars;
  jb.var = var;
  jb.type = type;
  jb.bound = bound;
  cudaStreamSynchronize(parent->stream);
  cudaError_t e = cudaMemsetAsync(d_scr, 0, off_key, ch->stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(d_scr + off_key, 0x7f, sizeof(int) * Cp, ch->stream);
  rc = e == cudaSuccess ? bb_run_addc(ch->stream, &jb, reinterpret_cast<BBAddc*>(d_scr + off_job), 1, ch->ld)
                        : fail(LPR_E_CUDA, "AddConstraint: %s", cudaGetErrorString(e));
  if (rc == LPR_OK && cudaStreamSynchronize(ch->stream) != cudaSuccess) rc = fail(LPR_E_CUDA, "AddConstraint failed");
  cudaFree(d_scr);
  if (rc) {
    lpr_tab_destroy(ch);
    return rc;
  }
  *child = ch;
  return LPR_OK;
}

int lpr_tab_bb_branch_var(lpr_tab* h, int n_vars, int* var, double* value, double* x) {
  if (!h || n_vars < 1 || n_vars > h->C - 1) return fail(LPR_E_BADARG, "bad arguments");
  int rc = select_device(h->device);
  if (rc) return rc;
  const double** d_tabs = nullptr;
  int* d_dims = nullptr;
  BBEval* d_ev = nullptr;
  double* d_x = nullptr;
  LPR_CUDA(cudaMalloc(&d_tabs, sizeof(double*)));
  LPR_CUDA(cudaMalloc(&d_dims, sizeof(int) * 3));
  LPR_CUDA(cudaMalloc(&d_ev, sizeof(BBEval)));
  LPR_CUDA(cudaMalloc(&d_x, sizeof(double) * n_vars));
  const double* tp = h->T;
  int dims[3] = {h->R, h->C, h->ld};
  BBEval ev;
  cudaError_t e = cudaMemcpyAsync(d_tabs, &tp, sizeof(double*), cudaMemcpyHostToDevice, h->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_dims, dims, sizeof dims, cudaMemcpyHostToDevice, h->stream);
  if (e == cudaSuccess) {
    if (bb_run_eval(h->stream, d_tabs, d_dims, 1, n_vars, d_ev, d_x) != LPR_OK) e = cudaErrorLaunchFailure;
    if (e == cudaSuccess) e = cudaMemcpyAsync(&ev, d_ev, sizeof ev, cudaMemcpyDeviceToHost, h->stream);
  }
  if (e == cudaSuccess && x) e = cudaMemcpyAsync(x, d_x, sizeof(double) * n_vars, cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  cudaFree(d_tabs); cudaFree(d_dims); cudaFree(d_ev); cudaFree(d_x);
  if (e != cudaSuccess) return fail(LPR_E_CUDA, "branch_var: %s", cudaGetErrorString(e));
  if (var) *var = ev.branch_var;
  if (value) *value = ev.branch_val;
  return LPR_OK;
}

}  // extern "C"
