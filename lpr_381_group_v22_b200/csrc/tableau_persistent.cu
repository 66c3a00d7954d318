// tableau_persistent.cu -- a whole pivot loop in ONE cooperative launch, for tableaux that live in L2.
//
// The generic rules (PrimalSimplexSolver2.cs:46-164, DualSimplex.cs:14-178) and the cutting-plane recursion
// (CuttingPlaneSolver.cs:64-229) run on small and mid-size tableaux (a 513 x 1537 tableau is 6.3 MB; the L2 holds
// 126 MB).  Per pivot the two-kernel path (k_select<RULE> + k_sweep<SKIP>, tableau.cu) costs two dependent launches and a
// single-CTA selection, and the cutting-plane driver adds several host round trips per cut: ~26 us per pivot where
// the data movement needs ~2.  Here every CTA of a cooperative grid
//   1. selects the pivot ITSELF, redundantly, from the current tableau (the candidates of a rule are one column and
//      one or two rows: staged into shared memory with L2 loads, then the same block_hyst_min / block_first_min scans
//      as k_select -- all CTAs reach the same decision from the same bits, so no exchange is needed);
//   2. applies the rank-1 update to its share of the rows OUT OF PLACE (buffer A -> buffer B: the tableau it selects
//      from is immutable while others are still reading it), with the rule's skip test on |f|;
//   3. meets the others at ONE grid barrier per pivot, after which the buffers swap roles.
// The cutting-plane program is a state machine over the same loop: CUT (Gomory row from the row whose fractional RHS
// is closest to 0.5, pivot on it -- the cut row only ever materialises as the normalised pivot row of the next
// tableau) -> DUAL while some RHS is negative -> PRIMAL2 while the objective row has a negative entry -> next cut.
// Arithmetic, tolerances and tie breaks are those of the two-kernel path (same helpers), so results are bit-identical;
// tests compare both with the oracle.  Tableau loads are ld.global.cg (L2): a line cached in an SM's L1 two pivots
// ago would be stale.
#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include "sweep.cuh"
#include "tableau.cuh"

namespace lpr {

enum : int { PROG_DUAL = 0, PROG_PRIMAL2 = 1, PROG_CUTTING_PLANE = 2 };
enum : int { PH_CUT = 0, PH_DUAL = 1, PH_PRIMAL2 = 2, PH_AFTER_CUT = 3, PH_END_OF_CUT = 4 };

struct PersistOut {
  int status;       // final status of the program
  int src;          // which buffer holds the final tableau
  int R;            // final row count (cuts append rows)
  int n_cuts;
  long long npiv;   // pivots of the solve programs / total pivots of the cutting plane
};

struct PersistArgs {
  double* buf[2];
  int ld, R, C, Rcap, src;
  int program, print_steps, max_cuts;
  long long max_pivots;
  int* log;  // (row, col) pairs of the solve programs
  long long log_cap;
  int* cut_log;  // (chosen_row, pivot_col, n_dual, n_primal) per cut
  int cut_log_cap;
  int vec_cap;  // doubles per shared staging vector (even)
  int xrows;    // 4: the CTA's first four tableau rows are prefetched into shared memory; 0: no room
  PersistOut* out;
  unsigned* bar;
  unsigned long long* prof;  // LPR_PERSIST_PROF=1: clocks CTA 0 spent per section of the DUAL pivot loop, else null
};

constexpr int kPT = 1024;  // one CTA per SM: a staging loop or a sweep pass over a row is one L2 round trip, not four

__device__ __forceinline__ double ldcg(const double* p) { return __ldcg(p); }
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem) {  // through L1: see the call site
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {  // L2 -> shared memory, not through L1
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}

// grid barrier on a monotonically increasing counter (reset by the host before the launch), in two halves: whatever
// a CTA does between arrive and wait must not touch the tableau the others are still writing
__device__ __forceinline__ void grid_arrive(unsigned* bar, unsigned& generation) {
  __syncthreads();
  generation++;
  if (threadIdx.x == kPT - 1) {  // the LAST warp drains the stores and polls: the scanning warps (the first ones) go on
    __threadfence();
    asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
  }
}
__device__ __forceinline__ void grid_wait(unsigned* bar, unsigned generation) {
  if (threadIdx.x == kPT - 1) {
    const unsigned target = generation * gridDim.x;
    unsigned v;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
    } while (v < target);
  }
  __syncthreads();
}

template <bool kProf>
__global__ void __launch_bounds__(kPT) k_persist(PersistArgs a) {
  extern __shared__ double s_dyn[];
  __shared__ MinIdx sm[32];
  __shared__ int smi[32];
  __shared__ double s_prhs;  // RHS entry of the normalised pivot row
  __shared__ int s_next_k;   // leaving row of the next dual pivot (found by the scanning warps during the grid barrier)
  double* s_a = s_dyn;                  // column-shaped staging (rows)
  double* s_b = s_dyn + a.vec_cap;      // row-shaped staging (columns): pivot row / cut row
  double* s_c = s_dyn + 2 * a.vec_cap;  // row-shaped staging: ratio candidates
  double2* s_x = reinterpret_cast<double2*>(s_dyn + 3 * a.vec_cap);  // a.xrows x ld: this CTA's first tableau rows
  const int tid = threadIdx.x;
  const int ld = a.ld, C = a.C;
  int R = a.R, src = a.src;
  unsigned generation = 0;
  const double kNaN = __longlong_as_double(0x7ff8000000000000LL);
  // program state (identical in every CTA)
  int phase = a.program == PROG_CUTTING_PLANE ? PH_CUT : (a.program == PROG_DUAL ? PH_DUAL : PH_PRIMAL2);
  int npiv = 0, total_piv = 0;  // pivots of the running DUAL / PRIMAL2 phase; all pivots (a launch ends at 2^31 - 1)
  int cuts = 0, chosen = -1, pcol = -1, nd = 0, np = 0;
  int status = LPR_RUNNING;
  const int ldv = ld >> 1;
  int rows_dealt = -1, r0 = 0, r1 = 0;  // this CTA's rows [r0, r1) of a tableau of rows_dealt rows
  bool dual_carry = false;  // s_a / next_k describe the CURRENT tableau (set by the previous dual pivot)
  int next_k = -1;
  const bool leader = blockIdx.x == 0 && tid == 0;
  long long t_last = kProf ? clock64() : 0;
  auto stamp = [&](int slot) {
    if (kProf && leader) {
      const long long t = clock64();
      a.prof[slot] += (unsigned long long)(t - t_last);
      t_last = t;
    }
  };

  auto log_cut = [&]() {
    if (leader && a.cut_log && cuts < a.cut_log_cap) {
      a.cut_log[4 * cuts + 0] = chosen;
      a.cut_log[4 * cuts + 1] = pcol;
      a.cut_log[4 * cuts + 2] = nd;
      a.cut_log[4 * cuts + 3] = np;
    }
    cuts++;
  };

  while (status == LPR_RUNNING) {
    const double* T = a.buf[src];
    double* D = a.buf[src ^ 1];
    int p = -1, e = -1;          // pivot position; p == R means "the virtual cut row"
    bool have_pivot = false;
    double eps_skip = 1e-9;
    // this CTA's rows of the next tableau (a cut pivot appends one); the first four of them are requested right away
    // (cp.async into shared memory, L2 path): the sweep's tableau loads depend on neither p nor e, so their L2 round trip
    // and transfer hide behind the selection.  A thread reads back only what it copied itself.
    const int Rn = phase == PH_CUT ? R + 1 : R;
    if (Rn != rows_dealt) {  // (a division by gridDim.x: not on every pivot)
      rows_dealt = Rn;
      r0 = (int)((unsigned)Rn * blockIdx.x / gridDim.x);
      r1 = (int)((unsigned)Rn * (blockIdx.x + 1) / gridDim.x);
    }
    if (a.xrows && (phase == PH_CUT || phase == PH_DUAL || phase == PH_PRIMAL2)) {
      const double2* S2 = reinterpret_cast<const double2*>(T);
      for (int c = tid; c < ldv; c += kPT)
#pragma unroll
        for (int u = 0; u < 4; u++)
          if (r0 + u < r1 && r0 + u < R) cp_async16(s_x + (size_t)u * ldv + c, S2 + (size_t)(r0 + u) * ldv + c);
      asm volatile("cp.async.commit_group;" ::: "memory");
    }

    // ---- the flags of CuttingPlaneSolver.cs:19-45 on the current tableau (after a cut pivot / at the end of a cut)
    if (phase == PH_AFTER_CUT || phase == PH_END_OF_CUT) {
      int neg = 0, nopt = 0, fr = 0;
      for (int i = 1 + tid; i < R; i += kPT) {
        const double rhs = ldcg(&TAT(T, ld, i, C - 1));
        if (rhs < -1e-9) neg++;
        if (net_frac(rhs) > 1e-9) fr++;
      }
      for (int j = tid; j < C - 1; j += kPT)
        if (ldcg(&T[j]) < -1e-9) nopt++;
      neg = block_sum_int(neg, smi);
      nopt = block_sum_int(nopt, smi);
      fr = block_sum_int(fr, smi);
      if (phase == PH_AFTER_CUT) {  // :186-212
        npiv = 0;
        phase = neg ? PH_DUAL : (nopt ? PH_PRIMAL2 : PH_END_OF_CUT);
        if (phase != PH_END_OF_CUT) continue;
      }
      // :215-228
      log_cut();
      if (!nopt && !neg) {
        if (fr) {
          phase = PH_CUT;
          continue;
        }
        status = LPR_OPTIMAL;
      } else {
        status = LPR_CUT_STEP_DONE;
      }
      break;
    }

    if (phase == PH_CUT) {
      // ---- Gomory row (:76-111) and the pivot column on it (:113-132)
      if (a.max_cuts >= 0 && cuts >= a.max_cuts) {
        status = LPR_ITER_LIMIT;
        break;
      }
      for (int i = tid; i < R; i += kPT) s_a[i] = ldcg(&TAT(T, ld, i, C - 1));
      __syncthreads();
      chosen = block_first_min(R - 1, [&](int q, double& val) {
        const double fr = net_frac(s_a[q + 1]);
        if (!(fr > 1e-9)) return false;
        val = fabs(__dsub_rn(fr, 0.5));
        return true;
      }, sm);
      if (chosen < 0) {
        status = LPR_NO_CUT_NEEDED;
        break;
      }
      if (R + 1 > a.Rcap) {  // no room for another cut row
        status = LPR_ITER_LIMIT;
        break;
      }
      nd = np = 0;
      pcol = -1;
      const double EPS = 1e-9;
      for (int j = tid; j < C; j += kPT) {
        const double cut = -net_frac(ldcg(&TAT(T, ld, chosen + 1, j)));
        s_b[j] = cut;
        double val = kNaN;
        if (j < C - 1 && cut < -EPS) {
          const double num = ldcg(&T[j]);
          if (fabs(num) > EPS) val = fabs(ddiv(num, cut));
        }
        s_c[j] = val;
      }
      __syncthreads();
      e = block_hyst_min_scan(C - 1, [&](int j, double& val) { val = s_c[j]; return val == val; }, kPosInf, EPS);
      pcol = e;
      if (e < 0 || fabs(s_b[e]) <= EPS) {
        // no pivot on the cut row: the reference has already appended it (:110) when it gives up (:134-141, :147-149)
        if (blockIdx.x == 0)
          for (int j = tid; j < ld; j += kPT) TAT(a.buf[src], ld, R, j) = j < C ? s_b[j] : 0.0;
        R = R + 1;
        log_cut();
        status = e < 0 ? LPR_NO_PIVOT_COL : LPR_PIVOT_TOO_SMALL;
        break;
      }
      p = R;  // the cut row: it exists only as s_b until the sweep writes its normalised form as row R
      have_pivot = true;
    } else if (phase == PH_DUAL) {
      const double EPS = 1e-9;
      // DualSimplex.cs:94,108: the counter advances only when printing and is tested after the pivot
      if (npiv > 0 && a.max_pivots >= 0 && (a.print_steps ? npiv : 0) >= a.max_pivots) {
        if (a.program == PROG_CUTTING_PLANE) {  // dual phase did not reach feasibility (:189-193)
          nd = (int)npiv;
          log_cut();
          status = LPR_INFEASIBLE;
        } else {
          status = LPR_ITER_LIMIT;
        }
        break;
      }
      stamp(0);
      int k;
      if (dual_carry) {  // RHS column and leaving row were derived before the last grid barrier completed (see below)
        k = next_k;
      } else {
        for (int i = tid; i < R; i += kPT) s_a[i] = ldcg(&TAT(T, ld, i, C - 1));
        __syncthreads();
        stamp(1);
        // :27-37 most negative RHS among the constraint rows
        k = block_hyst_min_scan(R - 1, [&](int q, double& val) { val = s_a[q + 1]; return true; }, 0.0, EPS);
      }
      dual_carry = false;
      stamp(2);
      if (k < 0) {  // feasible
        if (a.program == PROG_CUTTING_PLANE) {
          nd = (int)npiv;
          npiv = 0;
          // needPrimal is re-read after the dual phase (:196)
          int nopt = 0;
          for (int j = tid; j < C - 1; j += kPT)
            if (ldcg(&T[j]) < -1e-9) nopt++;
          nopt = block_sum_int(nopt, smi);
          phase = nopt ? PH_PRIMAL2 : PH_END_OF_CUT;
          continue;
        }
        status = LPR_OPTIMAL;
        break;
      }
      p = k + 1;
      // :50-70 min |obj_j / a_j| over a_j < -EPS, |obj_j| > EPS
      for (int j = tid; j < C; j += kPT) {
        const double av = ldcg(&TAT(T, ld, p, j));
        s_b[j] = av;
        double val = kNaN;
        if (j < C - 1 && av < -EPS) {
          const double num = ldcg(&T[j]);
          if (fabs(num) > EPS) val = fabs(ddiv(num, av));
        }
        s_c[j] = val;
      }
      __syncthreads();
      stamp(3);
      e = block_hyst_min_scan(C - 1, [&](int j, double& val) { val = s_c[j]; return val == val; }, kPosInf, EPS);
      stamp(4);
      int bad = 0;
      if (e < 0) bad = LPR_INFEASIBLE;
      else if (fabs(s_b[e]) <= EPS) bad = LPR_PIVOT_TOO_SMALL;
      if (bad) {
        if (a.program == PROG_CUTTING_PLANE) {
          nd = (int)npiv;
          log_cut();
        }
        status = bad;
        break;
      }
      have_pivot = true;
    } else {  // PH_PRIMAL2
      const double EPS = 1e-10;
      eps_skip = EPS;
      bool stop = false;
      int how = LPR_OPTIMAL;
      if (npiv > 0 && a.max_pivots >= 0 && (a.print_steps ? npiv : 0) >= a.max_pivots) {  // :75, :90
        stop = true;
        how = LPR_ITER_LIMIT;
      }
      if (!stop) {
        // FindEnteringColumn :102-117
        for (int j = tid; j < C; j += kPT) s_c[j] = ldcg(&T[j]);
        __syncthreads();
        e = block_hyst_min_scan(C - 1, [&](int j, double& val) { val = s_c[j]; return true; }, 0.0, EPS);
        if (e < 0) {
          stop = true;
          how = LPR_OPTIMAL;
        }
      }
      if (!stop) {
        // FindLeavingRow :120-141 (the precedence quirk reduces to ratio > EPS && ratio < best - EPS)
        for (int i = tid; i < R; i += kPT) {
          double val = kNaN;
          if (i >= 1) {
            const double av = ldcg(&TAT(T, ld, i, e));
            if (av > EPS) {
              const double r = ddiv(ldcg(&TAT(T, ld, i, C - 1)), av);
              if (r > EPS) val = r;
            }
          }
          s_a[i] = val;
        }
        __syncthreads();
        const int k = block_hyst_min_scan(R - 1, [&](int q, double& val) { val = s_a[q + 1]; return val == val; }, kPosInf,
                                          EPS);
        if (k < 0) {
          stop = true;
          how = LPR_UNBOUNDED;
        } else {
          p = k + 1;
          for (int j = tid; j < C; j += kPT) s_b[j] = ldcg(&TAT(T, ld, p, j));
          __syncthreads();
          if (fabs(s_b[e]) <= EPS) {
            stop = true;
            how = LPR_PIVOT_TOO_SMALL;
          }
        }
      }
      if (stop) {
        if (a.program == PROG_CUTTING_PLANE) {
          np = (int)npiv;
          npiv = 0;
          if (how == LPR_PIVOT_TOO_SMALL) {  // :148-149 throws
            log_cut();
            status = how;
            break;
          }
          phase = PH_END_OF_CUT;  // the bool result of Solve() is ignored by the reference (:203)
          continue;
        }
        status = how;
        break;
      }
      have_pivot = true;
    }

    if (!have_pivot) break;  // not reached
    if (a.program != PROG_CUTTING_PLANE) {  // the solve programs test the pivot cap before a PRIMAL pivot only in k_select<PRIMAL>
      if (leader && a.log && npiv < a.log_cap) {
        a.log[2 * (size_t)npiv] = p;
        a.log[2 * (size_t)npiv + 1] = e;
      }
    }

    // ---- rank-1 update, out of place: this CTA's rows of every column chunk ---------------------------------------
    {
      const double piv = s_b[e];
      const bool carry = phase == PH_DUAL;  // the next dual pivot's RHS column and leaving row are derived below
      // rows are dealt to the CTAs in contiguous spans; a thread owns column chunks tid, tid + kPT, ...
      const double2* __restrict__ S2 = reinterpret_cast<const double2*>(T);
      double2* __restrict__ D2 = reinterpret_cast<double2*>(D);
      const int rhs_chunk = (C - 1) >> 1;
      // entering column of ALL rows -> s_c (the ratio candidates are not needed any more): the factors of this CTA's rows,
      // and the scanning warps repeat the update of the RHS column from it, so the next leaving row is known without
      // reading the next tableau.  Asynchronous copies (no registers held across the divisions below); through L1 is
      // fine, the acquire of the last grid barrier emptied it (CCTL.IVALL) and T is not written during this pivot.
      // (Requested BEFORE arriving at the barrier: once every CTA has arrived, the fast ones start overwriting T.)
      for (int i = tid; i < R; i += kPT) cp_async8(s_c + i, &TAT(T, ld, i, e));
      asm volatile("cp.async.commit_group;" ::: "memory");
      double2 pr0 = {0.0, 0.0};  // normalised pivot row at this thread's first column chunk
      if (tid < ldv) {
        pr0.x = 2 * tid < C ? ddiv(s_b[2 * tid], piv) : 0.0;
        pr0.y = 2 * tid + 1 < C ? ddiv(s_b[2 * tid + 1], piv) : 0.0;
        if (tid == rhs_chunk % kPT && rhs_chunk < kPT) s_prhs = ((C - 1) & 1) ? pr0.y : pr0.x;
      }
      if (rhs_chunk >= kPT && tid == 0) s_prhs = ddiv(s_b[C - 1], piv);
      asm volatile("cp.async.wait_all;" ::: "memory");  // the tableau rows requested at the top, and column e
      __syncthreads();
      stamp(8);
      for (int c = tid; c < ldv; c += kPT) {
        double2 pr = pr0;
        if (c != tid) {
          pr.x = 2 * c < C ? ddiv(s_b[2 * c], piv) : 0.0;
          pr.y = 2 * c + 1 < C ? ddiv(s_b[2 * c + 1], piv) : 0.0;
        }
        for (int i0 = r0; i0 < r1; i0 += 4) {
          const bool pre = a.xrows && i0 == r0;  // these four were requested at the top of the pivot
          double2 x[4] = {};
#pragma unroll
          for (int u = 0; u < 4; u++) {
            const int i = i0 + u;
            if (i < r1 && i < R) x[u] = pre ? s_x[(size_t)u * ldv + c] : __ldcg(S2 + (size_t)i * ldv + c);
          }
#pragma unroll
          for (int u = 0; u < 4; u++) {
            const int i = i0 + u;
            if (i >= r1) continue;
            const double f = i < R ? s_c[i] : 0.0;
            double2 y;
            if (i == p) {
              y = pr;
            } else if (fabs(f) <= eps_skip) {
              y = x[u];
            } else {
              y.x = __dsub_rn(x[u].x, __dmul_rn(f, pr.x));
              y.y = __dsub_rn(x[u].y, __dmul_rn(f, pr.y));
            }
            D2[(size_t)i * ldv + c] = y;
          }
        }
      }
      stamp(5);
      grid_arrive(a.bar, generation);  // CTA barrier inside: s_prhs and s_c are visible
      if (carry && tid < kScanThreads) {
        // RHS column of the next tableau, the arithmetic of the sweep above on column C-1 (same skip test, same bits),
        // and the leaving row on it; the other warps wait at the barrier below, the last one drains the stores
        const double prhs = s_prhs;
        for (int i = tid; i < R; i += kScanThreads) {
          const double fi = s_c[i], x = s_a[i];
          s_a[i] = i == p ? prhs : (fabs(fi) <= eps_skip ? x : __dsub_rn(x, __dmul_rn(fi, prhs)));
        }
        scan_threads_barrier();
        const int nk = hyst_min_scan_threads(R - 1, [&](int q, double& val) { val = s_a[q + 1]; return true; }, 0.0, 1e-9);
        if (tid == 0) s_next_k = nk;
      }
      dual_carry = carry;
      if (p == R) {
        R = R + 1;
        phase = PH_AFTER_CUT;
      } else {
        npiv++;
      }
      total_piv++;
      src ^= 1;
    }
    stamp(6);
    grid_wait(a.bar, generation);  // CTA barrier inside: s_next_k is visible
    if (dual_carry) next_k = s_next_k;
    stamp(7);
  }

  if (leader) {
    a.out->status = status;
    a.out->src = src;
    a.out->R = R;
    a.out->n_cuts = cuts;
    a.out->npiv = a.program == PROG_CUTTING_PLANE ? total_piv : npiv;
  }
}

// ---- host side ----------------------------------------------------------------------------------------------------------
struct PersistRes {
  PersistOut* d_out = nullptr;
  PersistOut* h_out = nullptr;
  unsigned* d_bar = nullptr;
  int* d_cut_log = nullptr;
  int cut_log_cap = 0;
  int device = -1;
};
static thread_local PersistRes g_pr[16];

static int persist_resources(int device, int cut_log_cap, PersistRes** out) {
  if (device < 0 || device >= 16) return fail(LPR_E_BADARG, "device out of range");
  PersistRes& r = g_pr[device];
  if (r.device != device) {
    LPR_CUDA(cudaMalloc(&r.d_out, sizeof(PersistOut)));
    LPR_CUDA(cudaMallocHost(&r.h_out, sizeof(PersistOut)));
    LPR_CUDA(cudaMalloc(&r.d_bar, sizeof(unsigned)));
    r.device = device;
  }
  if (cut_log_cap > r.cut_log_cap) {
    if (r.d_cut_log) cudaFree(r.d_cut_log);
    r.d_cut_log = nullptr;
    r.cut_log_cap = 0;
    LPR_CUDA(cudaMalloc(&r.d_cut_log, sizeof(int) * 4 * (size_t)cut_log_cap));
    r.cut_log_cap = cut_log_cap;
  }
  *out = &r;
  return LPR_OK;
}

// can this tableau run on the persistent path?  (staging vectors must fit in shared memory; the tableau pair should
// be L2 resident for the path to pay off)
bool tab_persist_applicable(const lpr_tab* h) {
  const char* on = getenv("LPR_TAB_PERSIST");  // read per call: the tests run both paths in one process
  if (on && atoi(on) == 0) return false;
  const size_t vec = (size_t)std::max(h->Rcap, h->C) + 10;
  if (3 * vec * sizeof(double) > 200u * 1024u) return false;
  static const int max_mb = getenv("LPR_TAB_PERSIST_MAX_MB") ? atoi(getenv("LPR_TAB_PERSIST_MAX_MB")) : 48;
  return (size_t)h->Rcap * h->ld * sizeof(double) <= (size_t)max_mb << 20;
}

// run `program` on the handle's tableau; on return h->T holds the final tableau and h->R the final row count
int tab_run_persistent(lpr_tab* h, int program, int64_t max_pivots, int print_steps, int max_cuts, int* status,
                       int64_t* n_pivots, int* pivot_log, int64_t log_cap, int* n_cuts, int* cut_log, int cut_log_cap) {
  int rc = select_device(h->device);
  if (rc) return rc;
  if ((rc = tab_ensure_T2(h))) return rc;
  PersistRes* pr = nullptr;
  if ((rc = persist_resources(h->device, std::max(1, cut_log_cap), &pr))) return rc;
  if (pivot_log && log_cap > 0) {
    long long want = std::min<long long>(log_cap, 1LL << 24);
    if (max_pivots >= 0) want = std::min<long long>(want, max_pivots + 1);
    if ((rc = tab_ensure_log(h, std::max<long long>(1, want)))) return rc;
  }
  PersistArgs a;
  a.buf[0] = h->T;
  a.buf[1] = h->T2;
  a.ld = h->ld;
  a.R = h->R;
  a.C = h->C;
  a.Rcap = h->Rcap;
  a.src = 0;
  a.program = program;
  a.print_steps = print_steps;
  a.max_cuts = max_cuts;
  a.max_pivots = max_pivots;
  a.log = (pivot_log && log_cap > 0) ? h->log : nullptr;
  a.log_cap = (pivot_log && log_cap > 0) ? h->log_cap : 0;
  a.cut_log = pr->d_cut_log;
  a.cut_log_cap = std::max(0, std::min(cut_log_cap, pr->cut_log_cap));
  a.vec_cap = (std::max(h->Rcap, h->C) + 8 + 1) & ~1;
  a.out = pr->d_out;
  a.bar = pr->d_bar;
  a.prof = nullptr;
  static const bool want_prof = getenv("LPR_PERSIST_PROF") && atoi(getenv("LPR_PERSIST_PROF")) != 0;
  if (want_prof) {
    LPR_CUDA(cudaMalloc(&a.prof, 16 * sizeof(unsigned long long)));
    LPR_CUDA(cudaMemset(a.prof, 0, 16 * sizeof(unsigned long long)));
  }
  void (*kfn)(PersistArgs) = want_prof ? k_persist<true> : k_persist<false>;
  size_t smem = 3 * (size_t)a.vec_cap * sizeof(double);
  a.xrows = 0;
  const char* xr = getenv("LPR_PERSIST_XROWS");  // 0: as for tableaux too wide for the row prefetch (read per call: tests)
  if (!(xr && atoi(xr) == 0) && smem + 4 * (size_t)h->ld * sizeof(double) <= 200u * 1024u) {
    a.xrows = 4;
    smem += 4 * (size_t)h->ld * sizeof(double);
  }
  LPR_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));  // per device
  int per_sm = 0;
  LPR_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kfn, kPT, smem));
  if (per_sm < 1) return fail(LPR_E_CAPACITY, "persistent pivot kernel does not fit (%zu bytes of shared memory)", smem);
  // one CTA per SM is enough to saturate L2 for these sizes; never more CTAs than rows
  int grid = std::min(h->sms * std::min(per_sm, 1), std::max(1, h->R));
  LPR_CUDA(cudaEventRecord(h->ev0, h->stream));
  LPR_CUDA(cudaMemsetAsync(pr->d_bar, 0, sizeof(unsigned), h->stream));
  void* args[] = {&a};
  LPR_CUDA(cudaLaunchCooperativeKernel((void*)kfn, dim3(grid), dim3(kPT), args, smem, h->stream));
  count_launch();
  LPR_CUDA(cudaEventRecord(h->ev1, h->stream));
  LPR_CUDA(cudaMemcpyAsync(pr->h_out, pr->d_out, sizeof(PersistOut), cudaMemcpyDeviceToHost, h->stream));
  LPR_CUDA(cudaStreamSynchronize(h->stream));
  LPR_CUDA(cudaEventElapsedTime(&h->last_ms, h->ev0, h->ev1));
  const PersistOut o = *pr->h_out;
  if (a.prof) {  // development aid: where CTA 0 spends a DUAL pivot (clocks per section, summed over the launch)
    unsigned long long hp[16];
    LPR_CUDA(cudaMemcpy(hp, a.prof, sizeof(hp), cudaMemcpyDeviceToHost));
    cudaFree(a.prof);
    fprintf(stderr, "[lpr] k_persist sections (clk, %lld pivots, %.3f ms): other %llu | rhs stage %llu | row scan %llu | "
            "row stage %llu | col scan %llu | sweep: loads issued + pivot row normalised %llu, rest %llu | carried rhs + row scan %llu | barrier wait %llu\n",
            (long long)o.npiv, h->last_ms, hp[0], hp[1], hp[2], hp[3], hp[4], hp[8], hp[5], hp[6], hp[7]);
  }
  if (o.src) std::swap(h->T, h->T2);  // the final tableau is in the other buffer: swap roles, no copy
  h->R = o.R;
  if (status) *status = o.status;
  if (n_pivots) *n_pivots = o.npiv;
  if (n_cuts) *n_cuts = o.n_cuts;
  if (pivot_log && log_cap > 0 && o.npiv > 0) {
    const long long cnt = std::min<long long>(std::min<long long>(o.npiv, log_cap), h->log_cap);
    LPR_CUDA(cudaMemcpy(pivot_log, h->log, sizeof(int) * 2 * (size_t)cnt, cudaMemcpyDeviceToHost));
  }
  if (cut_log && cut_log_cap > 0 && o.n_cuts > 0) {
    const int cnt = std::min(std::min(o.n_cuts, cut_log_cap), pr->cut_log_cap);
    LPR_CUDA(cudaMemcpy(cut_log, pr->d_cut_log, sizeof(int) * 4 * (size_t)cnt, cudaMemcpyDeviceToHost));
  }
  return LPR_OK;
}

}  // namespace lpr
