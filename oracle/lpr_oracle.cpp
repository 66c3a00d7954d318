/*
 * oracle/lpr_oracle.cpp -- CPU oracle (TEST INFRASTRUCTURE ONLY; see lpr_oracle.h).
 *
 * C++17 restatement of the C# solver loops of Storm-Tarran/LPR_381_Group_V22.  Each function
 * cites the reference file:line it follows (relative to /root/reference/LPR_381_Group_V22/).
 * Arithmetic model: IEEE binary64, round-to-nearest-even, multiply and subtract as separate
 * roundings (RyuJIT never contracts to FMA) => build with -ffp-contract=off.
 * Pinning: the reference ships no tests and no .NET toolchain exists in this image, so the reference's
 * own .cs files are EXECUTED by the C# interpreter under oracle/csharp/ (tests/golden/make_reference_run.py)
 * and every function here is compared bit for bit with what the reference's classes returned
 * (tests/golden/reference_run.json, tests/test_reference_run.py), on top of the restated known answers of
 * SURVEY.md Appendix C (tests/test_oracle_golden.py).  The knapsack functions have no reference body to run
 * (missing upstream) and stay pinned to DP equality, the reference's own check (Program.cs:467-470).
 */
#include "lpr_oracle.h"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>
#include <numeric>
#include <vector>

#include <thread>

namespace {
constexpr double kInf = std::numeric_limits<double>::infinity();
inline double& at(double* T, int C, int i, int j) { return T[(size_t)i * C + j]; }
inline double at(const double* T, int C, int i, int j) { return T[(size_t)i * C + j]; }
inline void log_pivot(int* log, int64_t cap, int64_t k, int r, int c) {
  if (log && k < cap) {
    log[2 * k] = r;
    log[2 * k + 1] = c;
  }
}
}  // namespace

extern "C" {

/* ======================= synthetic generator ============================== */
uint64_t orc_splitmix64(uint64_t x) {
  uint64_t z = x + 0x9E3779B97F4A7C15ULL;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}
double orc_u01(uint64_t seed, uint64_t k) {
  return (double)(orc_splitmix64(seed + k) >> 11) * 0x1.0p-53;
}
static inline uint64_t stream(uint64_t s, uint64_t idx) { return (s << 40) + idx; }

void orc_gen_dense_lp(uint64_t seed, int m, int n, double* A, double* b, double* c) {
  for (int i = 0; i < m; i++)
    for (int j = 0; j < n; j++) A[(size_t)i * n + j] = 0.1 + orc_u01(seed, stream(0, (uint64_t)i * n + j));
  for (int i = 0; i < m; i++) b[i] = ((double)n / 4.0) * (1.0 + orc_u01(seed, stream(1, i)));
  for (int j = 0; j < n; j++) c[j] = 1.0 + orc_u01(seed, stream(2, j));
}
void orc_gen_dense_ip(uint64_t seed, int m, int n, double* A, double* b, double* c) {
  for (int i = 0; i < m; i++) {
    double s = 0.0;
    for (int j = 0; j < n; j++) {
      double a = 1.0 + std::floor(20.0 * orc_u01(seed, stream(0, (uint64_t)i * n + j)));
      A[(size_t)i * n + j] = a;
      s += a;
    }
    b[i] = std::floor(s / 4.0);
  }
  for (int j = 0; j < n; j++) c[j] = 1.0 + std::floor(30.0 * orc_u01(seed, stream(2, j)));
}
void orc_gen_knapsack(uint64_t seed, int n, double* w, double* v, double* capacity) {
  double sw = 0.0;
  for (int i = 0; i < n; i++) {
    w[i] = 1.0 + std::floor(1000.0 * orc_u01(seed, stream(0, i)));
    double vv = w[i] + std::floor(200.0 * orc_u01(seed, stream(1, i))) - 100.0;
    v[i] = vv < 1.0 ? 1.0 : vv;
    sw += w[i];
  }
  *capacity = std::floor(sw / 2.0);
}

/* ======================= Math.Round restatements ========================================== */
/* .NET Framework COMDouble::Round (floor(x+0.5) with the even-tie fix and copysign).  It is
 * value-identical to rint() in round-to-nearest mode for every finite double (the one case
 * where x+0.5 rounds up, x = 0.5-2^-54, is caught by the even-tie fix); tests check this. */
double orc_net_round(double x) {
  if (std::fabs(x) < 9.2e18 && x == (double)((int64_t)x)) return x;
  double t = x + 0.5;
  double f = std::floor(t);
  if (f == t && std::fmod(t, 2.0) != 0.0) f -= 1.0;
  return std::copysign(f, x);
}
/* Math.Round(value, 4): InternalRound -- value*1e4, Round, /1e4, only when |value| < 1e16
 * (BranchBoundSimplexSolver.cs:540-550 RoundNumber) */
double orc_net_round4(double x) {
  if (std::fabs(x) < 1e16) {
    x *= 1e4;
    x = orc_net_round(x);
    x /= 1e4;
  }
  return x;
}
/* CuttingPlaneSolver.cs:12-17 */
double orc_frac(double a) {
  double f = a - std::floor(a);
  if (std::fabs(f) < 1e-9 || std::fabs(1 - f) < 1e-9) return 0.0;
  return f;
}

/* ======================= PrimalSimplexSolver ============================================== */
/* Simplex/PrimalSimplexSolver.cs:27-87 */
void orc_primal_build(int n, int m, const double* objective, const double* coef, int coef_stride,
                      const int* coef_count, const int* relation, const double* rhs,
                      int is_maximization, double* T, int* basis) {
  const int R = m + 1, C = n + m + 1;
  std::fill(T, T + (size_t)R * C, 0.0);
  for (int i = 0; i < n; i++) T[i] = is_maximization ? -objective[i] : objective[i];  // :61-62
  for (int i = 0; i < m; i++) {
    const bool ge = relation && relation[i] == 1;  // :36-41 ">=" => negate row and RHS
    const int cnt = coef_count ? coef_count[i] : n;
    for (int j = 0; j < n; j++)
      if (j < cnt) {  // :68-72 only the first n coefficients are read
        double a = coef[(size_t)i * coef_stride + j];
        at(T, C, i + 1, j) = ge ? -a : a;
      }
    at(T, C, i + 1, n + i) = 1.0;  // :75-76
    basis[i] = n + i;              // :78
    at(T, C, i + 1, C - 1) = ge ? -rhs[i] : rhs[i];  // :82
  }
}

/* :152-167 */
int orc_primal_find_entering(int R, int C, const double* T) {
  (void)R;
  int e = -1;
  double most = 0;
  for (int j = 0; j < C - 1; j++)
    if (T[j] < most) {
      most = T[j];
      e = j;
    }
  return e;
}
/* :169-191 */
int orc_primal_find_leaving(int R, int C, const double* T, int col) {
  int leave = -1;
  double minRatio = DBL_MAX;
  for (int i = 1; i < R; i++) {
    double a = at(T, C, i, col);
    if (a > 1e-9) {
      double ratio = at(T, C, i, C - 1) / a;
      if (ratio >= 0 && ratio < minRatio) {
        minRatio = ratio;
        leave = i;
      }
    }
  }
  return leave;
}
/* :193-211.  The row loop is element-wise independent once the pivot row is normalised and
 * the factor column is read before each row, so splitting rows over threads is bit-identical. */
void orc_primal_pivot(int R, int C, double* T, int p, int e, int threads) {
  double piv = at(T, C, p, e);
  double* pr = T + (size_t)p * C;
  for (int j = 0; j < C; j++) pr[j] /= piv;
  auto rows = [=](int lo, int hi) {
    for (int i = lo; i < hi; i++) {
      if (i == p) continue;
      double* row = T + (size_t)i * C;
      double f = row[e];
      for (int j = 0; j < C; j++) row[j] -= f * pr[j];
    }
  };
  if (threads <= 1 || R < 2 * threads) {
    rows(0, R);
    return;
  }
  std::vector<std::thread> pool;
  for (int t = 0; t < threads; t++)
    pool.emplace_back(rows, (int)((int64_t)R * t / threads), (int)((int64_t)R * (t + 1) / threads));
  for (auto& th : pool) th.join();
}
/* :102-150 */
int orc_primal_solve(int R, int C, double* T, int* basis, int64_t max_pivots, int* status,
                     int64_t* n_pivots, int* pivot_log, int64_t log_cap, int threads) {
  int64_t k = 0;
  int st = ORC_RUNNING;
  while (true) {
    int e = orc_primal_find_entering(R, C, T);
    if (e == -1) {
      st = ORC_OPTIMAL;
      break;
    }
    int p = orc_primal_find_leaving(R, C, T, e);
    if (p == -1) {
      st = ORC_UNBOUNDED;
      break;
    }
    if (max_pivots >= 0 && k >= max_pivots) {
      st = ORC_ITER_LIMIT;
      break;
    }
    log_pivot(pivot_log, log_cap, k, p, e);
    orc_primal_pivot(R, C, T, p, e, threads);
    if (basis) basis[p - 1] = e;  // :142
    k++;
  }
  if (status) *status = st;
  if (n_pivots) *n_pivots = k;
  return 0;
}
/* :213-252 */
void orc_primal_extract(int R, int C, int n, const double* T, double* x) {
  for (int j = 0; j < n; j++) {
    x[j] = 0.0;
    int basicRow = -1;
    bool isBasic = true;
    for (int i = 1; i < R; i++) {
      double v = at(T, C, i, j);
      if (std::fabs(v - 1.0) < 1e-9) {
        if (basicRow == -1)
          basicRow = i;
        else {
          isBasic = false;
          break;
        }
      } else if (std::fabs(v) > 1e-9) {
        isBasic = false;
        break;
      }
    }
    if (isBasic && basicRow != -1) x[j] = at(T, C, basicRow, C - 1);
  }
}

/* ======================= PrimalSimplexSolver2 ============================================= */
/* Simplex/PrimalSimplexSolver2.cs:102-117 */
static int p2_entering(int C, const double* T) {
  const double EPS = 1e-10;
  int pc = -1;
  double mostNeg = 0.0;
  for (int j = 0; j < C - 1; j++) {
    double c = T[j];
    if (c < mostNeg - EPS || (std::fabs(c - mostNeg) <= EPS && pc != -1 && j < pc)) {
      mostNeg = c;
      pc = j;
    }
  }
  return pc;
}
/* :120-141 -- C# precedence makes the test ((A && B) || (C && D)) ? true : (i < bestRow) */
static int p2_leaving(int R, int C, const double* T, int pc) {
  const double EPS = 1e-10;
  int bestRow = -1;
  double bestRatio = kInf;
  for (int i = 1; i < R; i++) {
    double a = at(T, C, i, pc);
    if (a > EPS) {
      double ratio = at(T, C, i, C - 1) / a;
      bool cond = ((ratio > EPS && ratio < bestRatio - EPS) ||
                   (std::fabs(ratio - bestRatio) <= EPS && bestRow == -1))
                      ? true
                      : (i < bestRow);
      if (cond) {
        bestRatio = ratio;
        bestRow = i;
      }
    }
  }
  return bestRow;
}
/* generic in-place pivot with |f| skip used by PrimalSimplexSolver2.cs:145-164 (skip |f|<=EPS),
 * DualSimplex.cs:150-178 (skip unless |f|>EPS), SensitivityAnalyzer.cs:98-119 (skip |f|<EPS) */
static void pivot_skip(int R, int C, double* T, int pr, int pc, double eps, bool skip_if_lt) {
  double piv = at(T, C, pr, pc);
  double* prow = T + (size_t)pr * C;
  for (int j = 0; j < C; j++) prow[j] /= piv;
  for (int i = 0; i < R; i++) {
    if (i == pr) continue;
    double* row = T + (size_t)i * C;
    double f = row[pc];
    if (skip_if_lt ? (std::fabs(f) < eps) : (std::fabs(f) <= eps)) continue;
    for (int j = 0; j < C; j++) row[j] -= f * prow[j];
  }
}
/* :46-97 */
int orc_primal2_solve(int R, int C, double* T, int max_iters, int print_steps, int* status,
                      int64_t* n_pivots, int* pivot_log, int64_t log_cap) {
  const double EPS = 1e-10;
  int iter = 0;
  int64_t k = 0;
  int st = ORC_RUNNING;
  while (true) {
    int pc = p2_entering(C, T);
    if (pc == -1) {
      st = ORC_OPTIMAL;
      break;
    }
    int pr = p2_leaving(R, C, T, pc);
    if (pr == -1) {
      st = ORC_UNBOUNDED;
      break;
    }
    if (print_steps) ++iter;  // :75 -- the counter only advances when printing
    if (std::fabs(at(T, C, pr, pc)) <= EPS) {  // :148-149 throws InvalidOperationException
      st = ORC_PIVOT_TOO_SMALL;
      break;
    }
    log_pivot(pivot_log, log_cap, k, pr, pc);
    pivot_skip(R, C, T, pr, pc, EPS, false);
    k++;
    if (iter >= max_iters) {  // :90
      st = ORC_ITER_LIMIT;
      break;
    }
  }
  if (status) *status = st;
  if (n_pivots) *n_pivots = k;
  return 0;
}

/* ======================= DualSimplexSolver ================================================ */
/* Simplex/DualSimplex.cs:14-114; T row 0 = objectiveRow, rows 1.. = constraintRows.
 * pivot_log rows are 1-based tableau rows (constraint index + 1). */
int orc_dual_solve(int R, int C, double* T, int max_iters, int print_steps, int* status,
                   int64_t* n_pivots, int* pivot_log, int64_t log_cap) {
  const double EPS = 1e-9;
  int iter = 0;
  int64_t k = 0;
  int st = ORC_RUNNING;
  while (true) {
    int pivotRow = -1;  // index among constraint rows
    double mostNeg = 0.0;
    for (int r = 0; r < R - 1; r++) {  // :27-37
      double rhs = at(T, C, r + 1, C - 1);
      if (rhs < mostNeg - EPS || (std::fabs(rhs - mostNeg) <= EPS && pivotRow != -1 && r < pivotRow)) {
        mostNeg = rhs;
        pivotRow = r;
      }
    }
    if (pivotRow == -1) {
      st = ORC_OPTIMAL;  // "Dual phase complete" -> returns true
      break;
    }
    int pivotCol = -1;  // :50-70
    double bestRatio = kInf;
    for (int j = 0; j < C - 1; j++) {
      double a = at(T, C, pivotRow + 1, j);
      if (a < -EPS) {
        double num = T[j];
        if (std::fabs(num) > EPS) {
          double ratio = std::fabs(num / a);
          if (ratio < bestRatio - EPS ||
              (std::fabs(ratio - bestRatio) <= EPS && (pivotCol == -1 || j < pivotCol))) {
            bestRatio = ratio;
            pivotCol = j;
          }
        }
      }
    }
    if (pivotCol == -1) {
      st = ORC_INFEASIBLE;  // :72-76 returns false
      break;
    }
    if (print_steps) ++iter;  // :94
    if (std::fabs(at(T, C, pivotRow + 1, pivotCol)) <= EPS) {  // :155-156
      st = ORC_PIVOT_TOO_SMALL;
      break;
    }
    log_pivot(pivot_log, log_cap, k, pivotRow + 1, pivotCol);
    /* :150-178 constraint rows first, objective row last; element-wise so order is moot */
    pivot_skip(R, C, T, pivotRow + 1, pivotCol, EPS, false);
    k++;
    if (iter >= max_iters) {  // :108-112
      st = ORC_ITER_LIMIT;
      break;
    }
  }
  if (status) *status = st;
  if (n_pivots) *n_pivots = k;
  return 0;
}

/* ======================= SensitivityAnalyzer re-solve ===================================== */
/* SensitivityAnalyzer.cs:64-83: GetBasicRow / IsPivotColumn */
static int sens_basic_row(int R, int C, const double* T, int col) {
  const double EPS = 1e-9;
  for (int i = 1; i < R; i++) {
    if (std::fabs(at(T, C, i, col) - 1.0) < EPS) {
      bool pivot_col = true;
      for (int k = 1; k < R; k++)
        if (k != i && std::fabs(at(T, C, k, col)) > EPS) { pivot_col = false; break; }
      if (pivot_col) return i;
    }
  }
  return -1;
}
/* SensitivityAnalyzer.cs:706-723 */
void orc_sens_rebuild_basis(int R, int C, const double* T, int* basis) {
  const double EPS = 1e-9;
  for (int i = 1; i < R; i++) {
    basis[i - 1] = -1;
    for (int j = 0; j < C - 1; j++) {
      if (std::fabs(at(T, C, i, j) - 1.0) < EPS) {
        bool pivot_col = true;
        for (int k = 1; k < R; k++)
          if (k != i && std::fabs(at(T, C, k, j)) > EPS) { pivot_col = false; break; }
        if (pivot_col) { basis[i - 1] = j; break; }
      }
    }
  }
}
/* SensitivityAnalyzer.cs:158-164 */
void orc_sens_solution(int R, int C, const double* T, double* x) {
  for (int j = 0; j < C - 1; j++) {
    int r = sens_basic_row(R, C, T, j);
    x[j] = (r == -1) ? 0.0 : at(T, C, r, C - 1);
  }
}
/* SensitivityAnalyzer.cs:609-659 (the part before ResolveAll) */
void orc_sens_add_constraint(int R, int C, const double* T, int* basis, const double* tech,
                             double rhs_minus_ax, double* Tout) {
  const int C2 = C + 1;
  const int oldM = R - 1, oldNPlusM = C - 1;
  for (int i = 0; i < (R + 1) * C2; i++) Tout[i] = 0.0;
  for (int i = 0; i < R; i++) {
    for (int j = 0; j < C - 1; j++) Tout[i * C2 + j] = at(T, C, i, j);
    Tout[i * C2 + C] = at(T, C, i, C - 1);  // RHS moves one column to the right
  }
  const int newSlackCol = C - 1;
  for (int i = 0; i < R + 1; i++) Tout[i * C2 + newSlackCol] = (i == R) ? 1.0 : 0.0;
  for (int j = 0; j < oldNPlusM; j++) {
    double coeff = -tech[j];
    for (int pos = 0; pos < oldM; pos++) {
      int basicCol = basis[pos];
      coeff += tech[basicCol] * at(T, C, pos + 1, j);
    }
    Tout[R * C2 + j] = coeff;
  }
  Tout[R * C2 + C] = rhs_minus_ax;
  Tout[0 * C2 + newSlackCol] = 0.0;
  basis[R - 1] = newSlackCol;
}

/* SensitivityAnalysis/SensitivityAnalyzer.cs:168-201 (DualSimplexIfNeeded) then :121-166
 * (ReOptimize loop; the solution rebuild is done by the caller). */
int orc_sens_resolve(int R, int C, double* T, int* basis, int max_iter, int* status,
                     int64_t* n_pivots, int* pivot_log, int64_t log_cap) {
  const double EPS = 1e-9;
  int64_t k = 0;
  int st = ORC_RUNNING;
  auto is_basic = [&](int j) {
    for (int i = 0; i < R - 1; i++)
      if (basis[i] == j) return true;
    return false;
  };
  int iter = 0;
  while (st == ORC_RUNNING) {  // dual phase
    int leave = -1;
    double mostNeg = 0.0;
    for (int i = 1; i < R; i++) {
      double bi = at(T, C, i, C - 1);
      if (bi < mostNeg - EPS) {
        mostNeg = bi;
        leave = i;
      }
    }
    if (leave == -1) break;
    if (iter++ > max_iter) {
      st = ORC_ITER_LIMIT;
      break;
    }
    int enter = -1;
    double best = kInf;
    for (int j = 0; j < C - 1; j++) {
      double aij = at(T, C, leave, j);
      if (aij < -EPS) {
        double ratio = T[j] / (-aij);
        if (ratio < best - EPS) {
          best = ratio;
          enter = j;
        }
      }
    }
    if (enter == -1) {
      st = ORC_INFEASIBLE;
      break;
    }
    if (std::fabs(at(T, C, leave, enter)) < EPS) {
      st = ORC_PIVOT_TOO_SMALL;
      break;
    }
    log_pivot(pivot_log, log_cap, k, leave, enter);
    pivot_skip(R, C, T, leave, enter, EPS, true);
    if (leave - 1 < R - 1) basis[leave - 1] = enter;
    k++;
  }
  iter = 0;
  while (st == ORC_RUNNING) {  // primal phase
    bool optimal = true;
    for (int j = 0; j < C - 1; j++) {
      if (is_basic(j)) continue;
      if (T[j] < -EPS) {
        optimal = false;
        break;
      }
    }
    if (optimal) {
      st = ORC_OPTIMAL;
      break;
    }
    if (iter++ > max_iter) {
      st = ORC_ITER_LIMIT;
      break;
    }
    int enter = -1;
    double mostNeg = 0.0;
    for (int j = 0; j < C - 1; j++) {
      if (is_basic(j)) continue;
      double rc = T[j];
      if (rc < mostNeg) {
        mostNeg = rc;
        enter = j;
      }
    }
    if (enter == -1) {
      st = ORC_OPTIMAL;
      break;
    }
    int leave = -1;
    double bestRatio = kInf;
    for (int i = 1; i < R; i++) {
      double aij = at(T, C, i, enter);
      if (aij > EPS) {
        double ratio = at(T, C, i, C - 1) / aij;
        if (ratio < bestRatio - EPS) {
          bestRatio = ratio;
          leave = i;
        }
      }
    }
    if (leave == -1) {
      st = ORC_UNBOUNDED;
      break;
    }
    if (std::fabs(at(T, C, leave, enter)) < EPS) {
      st = ORC_PIVOT_TOO_SMALL;
      break;
    }
    log_pivot(pivot_log, log_cap, k, leave, enter);
    pivot_skip(R, C, T, leave, enter, EPS, true);
    basis[leave - 1] = enter;
    k++;
  }
  if (status) *status = st;
  if (n_pivots) *n_pivots = k;
  return 0;
}

/* ======================= CuttingPlaneSolver =============================================== */
/* IntegerProgramming/CuttingPlaneSolver.cs:76-107: choose the constraint row whose RHS fractional part is closest
 * to 0.5 and build cut = -Frac(row).  The reference collects the fractional rows in row order, SORTS them with
 * List<T>.Sort(Comparison) by |Frac(rhs) - 0.5| and takes element 0 (:94-96).  List<T>.Sort is the .NET Framework's
 * introspective sort (ArraySortHelper<T>.IntrospectiveSort, mscorlib 4.5+): insertion sort up to 16 elements -- stable,
 * so element 0 is the FIRST minimum -- and median-of-three quicksort (heapsort at the depth limit) above that, which is
 * not stable: when MORE THAN 16 rows are fractional AND several of them tie exactly for the smallest key, element 0 is
 * whichever of them the sort's swaps leave there.  dotnet_introsort_first() restates that sort literally (it was
 * checked against the reference executed by oracle/csharp, tests/test_reference_run.py); outside that corner it returns
 * the first minimum.  orc_set_gomory_first_min(1) switches to the plain first minimum, which is what the CUDA kernels
 * implement ("known deviation"). */
namespace {
int g_gomory_first_min = 0;
int64_t g_gomory_tie_corners = 0;  // selections in which the literal sort and the first minimum disagree

struct FracRow {
  int idx;
  double key;
};
inline int cmp_key(const FracRow& a, const FracRow& b) {  // double.CompareTo on non-NaN keys
  return a.key < b.key ? -1 : (a.key > b.key ? 1 : 0);
}

// ArraySortHelper<T>.IntrospectiveSort(keys, 0, n, comparer) with depthLimit = 2 * FloorLog2(capacity), where
// capacity is the length of the List<T>'s backing array (4, 8, 16, ... for a list filled by Add, List.cs EnsureCapacity)
struct DotnetIntroSort {
  std::vector<FracRow>& k;
  explicit DotnetIntroSort(std::vector<FracRow>& keys) : k(keys) {}
  void swap_if_greater(int a, int b) {
    if (a != b && cmp_key(k[a], k[b]) > 0) std::swap(k[a], k[b]);
  }
  void insertion(int lo, int hi) {
    for (int i = lo; i < hi; i++) {
      int j = i;
      FracRow t = k[i + 1];
      while (j >= lo && cmp_key(t, k[j]) < 0) {
        k[j + 1] = k[j];
        j--;
      }
      k[j + 1] = t;
    }
  }
  void down_heap(int i, int n, int lo) {
    FracRow d = k[lo + i - 1];
    while (i <= n / 2) {
      int child = 2 * i;
      if (child < n && cmp_key(k[lo + child - 1], k[lo + child]) < 0) child++;
      if (!(cmp_key(d, k[lo + child - 1]) < 0)) break;
      k[lo + i - 1] = k[lo + child - 1];
      i = child;
    }
    k[lo + i - 1] = d;
  }
  void heapsort(int lo, int hi) {
    int n = hi - lo + 1;
    for (int i = n / 2; i >= 1; i--) down_heap(i, n, lo);
    for (int i = n; i > 1; i--) {
      std::swap(k[lo], k[lo + i - 1]);
      down_heap(1, i - 1, lo);
    }
  }
  int partition(int lo, int hi) {
    int mid = lo + (hi - lo) / 2;
    swap_if_greater(lo, mid);
    swap_if_greater(lo, hi);
    swap_if_greater(mid, hi);
    FracRow pivot = k[mid];
    std::swap(k[mid], k[hi - 1]);
    int left = lo, right = hi - 1;
    while (left < right) {
      while (cmp_key(k[++left], pivot) < 0) {
      }
      while (cmp_key(pivot, k[--right]) < 0) {
      }
      if (left >= right) break;
      std::swap(k[left], k[right]);
    }
    std::swap(k[left], k[hi - 1]);
    return left;
  }
  void intro(int lo, int hi, int depth) {
    while (hi > lo) {
      int size = hi - lo + 1;
      if (size <= 16) {
        if (size == 1) return;
        if (size == 2) {
          swap_if_greater(lo, hi);
          return;
        }
        if (size == 3) {
          swap_if_greater(lo, hi - 1);
          swap_if_greater(lo, hi);
          swap_if_greater(hi - 1, hi);
          return;
        }
        insertion(lo, hi);
        return;
      }
      if (depth == 0) {
        heapsort(lo, hi);
        return;
      }
      depth--;
      int p = partition(lo, hi);
      intro(p + 1, hi, depth);
      hi = p - 1;
    }
  }
  void sort() {
    int n = (int)k.size();
    if (n < 2) return;
    int cap = 4;
    while (cap < n) cap *= 2;
    int log2 = 0;
    for (int c = cap; c >= 1; c /= 2) log2++;
    intro(0, n - 1, 2 * log2);
  }
};

int dotnet_introsort_first(std::vector<FracRow> rows) {
  DotnetIntroSort(rows).sort();
  return rows[0].idx;
}
}  // namespace

void orc_set_gomory_first_min(int on) { g_gomory_first_min = on; }
int64_t orc_gomory_tie_corners(void) { return g_gomory_tie_corners; }

int orc_gomory_cut(int R, int C, const double* T, double* cut) {
  const double EPS = 1e-9;
  int first_min = -1;
  double bestKey = kInf;
  std::vector<FracRow> rows;
  for (int i = 0; i < R - 1; i++) {
    double rhsFrac = orc_frac(at(T, C, i + 1, C - 1));
    if (rhsFrac > EPS) {
      double key = std::fabs(rhsFrac - 0.5);
      rows.push_back({i, key});
      if (key < bestKey) {
        bestKey = key;
        first_min = i;
      }
    }
  }
  if (first_min < 0) return -1;
  int chosen = first_min;
  if (rows.size() > 16) {
    int literal = dotnet_introsort_first(rows);
    if (literal != first_min) g_gomory_tie_corners++;
    if (!g_gomory_first_min) chosen = literal;
  }
  for (int j = 0; j < C; j++) cut[j] = -orc_frac(at(T, C, chosen + 1, j));
  return chosen;
}

static bool cp_obj_optimal(int C, const double* T) {  // :19-25
  for (int j = 0; j < C - 1; j++)
    if (T[j] < -1e-9) return false;
  return true;
}
static bool cp_any_negative_rhs(int R, int C, const double* T) {  // :27-35
  for (int i = 1; i < R; i++)
    if (at(T, C, i, C - 1) < -1e-9) return true;
  return false;
}
static bool cp_any_fractional_rhs(int R, int C, const double* T) {  // :37-45
  for (int i = 1; i < R; i++)
    if (orc_frac(at(T, C, i, C - 1)) > 1e-9) return true;
  return false;
}

/* :64-229, recursion unrolled into a loop */
int orc_cutting_plane(int* Rio, int C, double* T, int row_cap, int max_cuts, int* status,
                      int* n_cuts, int* cut_log, int cut_log_cap) {
  const double EPS = 1e-9;
  int R = *Rio;
  int cuts = 0;
  int st = ORC_RUNNING;
  std::vector<double> cut(C);
  while (true) {
    if (max_cuts >= 0 && cuts >= max_cuts) {
      st = ORC_ITER_LIMIT;
      break;
    }
    int chosen = orc_gomory_cut(R, C, T, cut.data());
    if (chosen < 0) {
      st = ORC_NO_CUT_NEEDED;  // :87-91 "All RHS are integers"
      break;
    }
    if (R + 1 > row_cap) {
      st = ORC_ITER_LIMIT;
      break;
    }
    std::memcpy(T + (size_t)R * C, cut.data(), sizeof(double) * C);  // :110
    int cutRow = R;                                                   // tableau row index
    R++;
    int pivotCol = -1;  // :113-132
    double bestRatio = kInf;
    for (int j = 0; j < C - 1; j++) {
      double a = cut[j];
      if (a < -EPS) {
        double num = T[j];
        if (std::fabs(num) > EPS) {
          double ratio = std::fabs(num / a);
          if (ratio < bestRatio - EPS ||
              (std::fabs(ratio - bestRatio) <= EPS && (pivotCol == -1 || j < pivotCol))) {
            bestRatio = ratio;
            pivotCol = j;
          }
        }
      }
    }
    int nd = 0, np = 0;
    auto log_cut = [&]() {
      if (cut_log && cuts < cut_log_cap) {
        cut_log[4 * cuts + 0] = chosen;
        cut_log[4 * cuts + 1] = pivotCol;
        cut_log[4 * cuts + 2] = nd;
        cut_log[4 * cuts + 3] = np;
      }
      cuts++;
    };
    if (pivotCol == -1) {  // :134-138
      log_cut();
      st = ORC_NO_PIVOT_COL;
      break;
    }
    if (std::fabs(at(T, C, cutRow, pivotCol)) <= EPS) {  // :145-150
      log_cut();
      st = ORC_PIVOT_TOO_SMALL;
      break;
    }
    pivot_skip(R, C, T, cutRow, pivotCol, EPS, false);  // :152-176
    bool needDual = cp_any_negative_rhs(R, C, T);
    bool needPrimal = !cp_obj_optimal(C, T);
    if (needDual) {  // :186-194 printSteps:true => iteration cap is live
      int dst;
      int64_t dn;
      orc_dual_solve(R, C, T, 10000, 1, &dst, &dn, nullptr, 0);
      nd = (int)dn;
      if (dst != ORC_OPTIMAL) {
        log_cut();
        st = (dst == ORC_PIVOT_TOO_SMALL) ? ORC_PIVOT_TOO_SMALL : ORC_INFEASIBLE;
        break;
      }
      needPrimal = !cp_obj_optimal(C, T);
    }
    if (needPrimal) {  // :196-212 -- result of Solve is ignored, tableau copied back
      int pst;
      int64_t pn;
      orc_primal2_solve(R, C, T, 10000, 1, &pst, &pn, nullptr, 0);
      np = (int)pn;
      if (pst == ORC_PIVOT_TOO_SMALL) {
        log_cut();
        st = ORC_PIVOT_TOO_SMALL;
        break;
      }
    }
    log_cut();
    if (cp_obj_optimal(C, T) && !cp_any_negative_rhs(R, C, T)) {  // :215-226
      if (cp_any_fractional_rhs(R, C, T)) continue;
      st = ORC_OPTIMAL;
      break;
    }
    st = ORC_CUT_STEP_DONE;  // :228 "further steps may be required"
    break;
  }
  *Rio = R;
  if (status) *status = st;
  if (n_cuts) *n_cuts = cuts;
  return 0;
}

/* ======================= BranchBoundSimplexSolver ========================================= */
/* RoundTableau BranchBoundSimplexSolver.cs:552-567 */
void orc_bb_round_tableau(int64_t count, double* T) {
  for (int64_t i = 0; i < count; i++) T[i] = orc_net_round4(T[i]);
}
static void bb_fix_negzero(int64_t count, double* T) {  // :307-313 (-0.0 == 0.0 is true in C#)
  for (int64_t i = 0; i < count; i++)
    if (T[i] == 0.0) T[i] = 0.0;
}
/* out-of-place pivot shared by :174-192 and :257-271 */
static void bb_pivot(int R, int C, const double* T, double* out, int r, int c) {
  double piv = at(T, C, r, c);
  for (int j = 0; j < C; j++) {
    double v = at(T, C, r, j) / piv;
    if (v == 0.0) v = 0.0;  // "== -0.0" is true for both zeros
    at(out, C, r, j) = v;
  }
  for (int i = 0; i < R; i++) {
    if (i == r) continue;
    double f = at(T, C, i, c);
    for (int j = 0; j < C; j++) at(out, C, i, j) = at(T, C, i, j) - (f * at(out, C, r, j));
  }
}
/* PerformDualPivot :115-201 */
int orc_bb_dual_pivot(int R, int C, const double* T, double* out, int* prow, int* pcol) {
  int r = -1;
  double minRhs = 0.0;
  for (int i = 0; i < R; i++) {  // row 0 included (:118-123); first index of the minimum
    double v = at(T, C, i, C - 1);
    if (v < 0 && (r == -1 || v < minRhs)) {
      minRhs = v;
      r = i;
    }
  }
  if (r == -1) return 0;
  bool allZeroOrInf = true;
  double minPos = kInf;
  for (int j = 0; j < C - 1; j++) {  // :126-143
    double th = kInf;
    if (at(T, C, r, j) < 0) th = std::fabs(T[j] / at(T, C, r, j));
    if (!(th == 0 || th == kInf)) allZeroOrInf = false;
    if (th > 0 && th < minPos) minPos = th;
  }
  double target = allZeroOrInf ? 0.0 : minPos;  // :146-148
  int c = -1;
  for (int j = 0; j < C - 1; j++) {  // IndexOf :154
    double th = kInf;
    if (at(T, C, r, j) < 0) th = std::fabs(T[j] / at(T, C, r, j));
    if (th == target) {
      c = j;
      break;
    }
  }
  if (c == -1) return 0;  // indexer throws => caught => (tableau, null) :165-172
  bb_pivot(R, C, T, out, r, c);
  if (prow) *prow = r;
  if (pcol) *pcol = c;
  return 1;
}
/* PerformPrimalPivot :203-279 */
static int bb_primal_pivot_ex(int R, int C, const double* T, double* out, int* prow, int* pcol, int is_min);
int orc_bb_primal_pivot(int R, int C, const double* T, double* out, int* prow, int* pcol) {
  return bb_primal_pivot_ex(R, C, T, out, prow, pcol, 0);
}
static int bb_primal_pivot_ex(int R, int C, const double* T, double* out, int* prow, int* pcol, int is_min) {
  int c = -1;
  double pv = 0.0;
  for (int j = 0; j < C - 1; j++) {  // :209-213 Min() of the negatives (max) / of the positives (min); IndexOf => first
    double v = T[j];
    if ((is_min ? v > 0 : v < 0) && (c == -1 || v < pv)) {
      pv = v;
      c = j;
    }
  }
  if (c == -1) return 0;
  if (R <= 1) return 0;  // thetas empty => All(<0) is true :228-231
  bool allNeg = true, anyPosFinite = false, hasZero = false;
  double minTheta = kInf;
  for (int i = 1; i < R; i++) {  // :222-225
    double a = at(T, C, i, c);
    double th = (a != 0) ? at(T, C, i, C - 1) / a : kInf;
    if (!(th < 0)) allNeg = false;
    if (th > 0 && th != kInf) {
      anyPosFinite = true;
      if (th < minTheta) minTheta = th;
    }
    if (th == 0) hasZero = true;
  }
  if (allNeg) return 0;
  if (!anyPosFinite) {  // :234-240
    if (hasZero)
      minTheta = 0.0;
    else
      return 0;
  }
  int r = -1;
  for (int i = 1; i < R; i++) {  // :249 IndexOf
    double a = at(T, C, i, c);
    double th = (a != 0) ? at(T, C, i, C - 1) / a : kInf;
    if (th == minTheta) {
      r = i;
      break;
    }
  }
  if (r == -1 || at(T, C, r, c) == 0) return 0;
  bb_pivot(R, C, T, out, r, c);
  if (prow) *prow = r;
  if (pcol) *pcol = c;
  return 1;
}
/* DoDualSimplex with tableauOverride :289-468 */
int orc_bb_node_solve(int R, int C, double* T, int64_t max_pivots, int64_t* n_pivots,
                      int* pivot_log, int64_t log_cap) {
  return orc_bb_node_solve_ex(R, C, T, 0, max_pivots, n_pivots, pivot_log, log_cap);
}
/* DoDualSimplex :289-468 with its isMinimization argument (only RunBranchAndBound :1271 passes true) */
int orc_bb_node_solve_ex(int R, int C, double* T, int is_min, int64_t max_pivots, int64_t* n_pivots,
                         int* pivot_log, int64_t log_cap) {
  const size_t N = (size_t)R * C;
  std::vector<double> a(T, T + N), b(N), prev;
  double* cur = a.data();
  double* nxt = b.data();
  int64_t npiv = 0;
  bool have_prev = false;
  auto all_rhs_ge = [&](const double* X, double lim) {
    for (int i = 0; i < R; i++)
      if (!(at(X, C, i, C - 1) >= lim)) return false;
    return true;
  };
  auto obj_opt = [&](const double* X) {  // :346-348
    for (int j = 0; j < C - 1; j++)
      if (is_min ? !(X[j] <= 0) : !(X[j] >= 0)) return false;
    return true;
  };
  int result = ORC_OPTIMAL;
  while (true) {  // dual phase :305-343
    bb_fix_negzero(N, cur);
    if (all_rhs_ge(cur, -1e-9)) break;
    if (max_pivots >= 0 && npiv >= max_pivots) {
      result = ORC_ITER_LIMIT;
      break;
    }
    int pr, pc;
    if (!orc_bb_dual_pivot(R, C, cur, nxt, &pr, &pc)) {
      result = ORC_INFEASIBLE;  // :324-331
      break;
    }
    bb_fix_negzero(N, nxt);
    log_pivot(pivot_log, log_cap, npiv, pr, pc);
    std::swap(cur, nxt);
    have_prev = true;
    npiv++;
  }
  if (result == ORC_OPTIMAL && !obj_opt(cur)) {  // :345-401
    while (true) {
      bb_fix_negzero(N, cur);
      if (obj_opt(cur)) break;
      if (max_pivots >= 0 && npiv >= max_pivots) {
        result = ORC_ITER_LIMIT;
        break;
      }
      int pr, pc;
      if (!bb_primal_pivot_ex(R, C, cur, nxt, &pr, &pc, is_min)) break;  // :375-387
      log_pivot(pivot_log, log_cap, npiv, pr, pc);
      std::swap(cur, nxt);
      have_prev = true;
      npiv++;
    }
    if (result == ORC_OPTIMAL && !all_rhs_ge(cur, 0.0)) {  // :392-400
      if (!have_prev || npiv == 0) {
        result = ORC_INFEASIBLE;  // pivotColumns.RemoveAt(-1) throws -> branch "failed"
      } else {
        std::swap(cur, nxt);  // drop the last tableau: previous one becomes the result
        npiv--;
      }
    }
  }
  if (result == ORC_OPTIMAL || result == ORC_ITER_LIMIT) std::memcpy(T, cur, N * sizeof(double));
  if (n_pivots) *n_pivots = npiv;
  return result;
}
/* FormulateTableau :28-113.  cons is m rows of `stride` doubles, row i has len[i] entries
 * [coefficients..., rhs, type flag]; rows whose flag == 1 are negated entirely (:42-51), the flag is dropped,
 * every entry but the last is copied into columns 0.. (:96-99: a row longer than n+2, like the ones
 * ConfigureProblem :1233-1251 appends, spills its extra 0 into the first slack column), the last entry is
 * the RHS, and row i gets a 1 at column i+n-1 (:103-109) whatever its type.  T is (m+1) x (n+m+1). */
void orc_bb_formulate(int n, int m, const double* objective, const double* cons, int stride, const int* len, double* T) {
  const int W = n + m + 1;
  for (int i = 0; i < (m + 1) * W; i++) T[i] = 0.0;
  for (int j = 0; j < n; j++) T[j] = -objective[j];
  for (int i = 0; i < m; i++) {
    const double* row = cons + (size_t)i * stride;
    const int L = len[i];
    const double sgn = (row[L - 1] == 1) ? -1.0 : 1.0;
    for (int j = 0; j < L - 2; j++) T[(i + 1) * W + j] = (sgn < 0) ? -1 * row[j] : row[j];
    T[(i + 1) * W + W - 1] = (sgn < 0) ? -1 * row[L - 2] : row[L - 2];
  }
  for (int i = 1; i <= m; i++)
    if (n < W - 1) T[i * W + i + n - 1] = 1;
}
/* IdentifyBasicVariables :642-692 */
int orc_bb_identify_basic(int R, int C, const double* T, int* basic) {
  const double epsilon = 1e-6;
  std::vector<int> vars;
  std::vector<int> key;
  for (int k = 0; k < C; k++) {
    double sum = 0.0;
    for (int i = 0; i < R; i++) sum += orc_net_round4(at(T, C, i, k));
    sum = orc_net_round4(sum);
    if (std::fabs(sum - 1.0) <= epsilon) {
      int first1 = R;
      for (int i = 0; i < R; i++)
        if (orc_net_round4(at(T, C, i, k)) == 1.0) {
          first1 = i;
          break;
        }
      vars.push_back(k);
      key.push_back(first1);
    }
  }
  std::vector<int> order(vars.size());
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return key[x] < key[y]; });
  for (size_t i = 0; i < order.size(); i++) basic[i] = vars[order[i]];
  return (int)vars.size();
}
/* AddConstraint :694-803, one new constraint [e_var | bound | type] */
void orc_bb_add_constraint(int R, int C, const double* base, int n_vars, int var, double bound,
                           int type, double* out) {
  const double epsilon = 1e-6;
  std::vector<double> work(base, base + (size_t)R * C);
  orc_bb_round_tableau((int64_t)R * C, work.data());  // :702
  std::vector<int> basic(C);
  int nb = orc_bb_identify_basic(R, C, work.data(), basic.data());
  const int C2 = C + 1, R2 = R + 1;
  for (int i = 0; i < R; i++) {  // :716-719 insert a zero column before the RHS
    for (int j = 0; j < C - 1; j++) at(out, C2, i, j) = at(work.data(), C, i, j);
    at(out, C2, i, C - 1) = 0.0;
    at(out, C2, i, C) = at(work.data(), C, i, C - 1);
  }
  for (int j = 0; j < C2; j++) at(out, C2, R, j) = 0.0;  // :721-725
  for (int i = 0; i < n_vars; i++) at(out, C2, R, i) = orc_net_round4(i == var ? 1.0 : 0.0);  // :727-730
  at(out, C2, R, C2 - 1) = orc_net_round4(bound);  // :732
  at(out, C2, R, C - 1) = (type == 1) ? -1.0 : 1.0;  // :734-742
  orc_bb_round_tableau((int64_t)R2 * C2, out);  // :747
  const int cr = R;
  for (int b = 0; b < nb; b++) {  // :756-796
    int colIndex = basic[b];
    double coefficient = orc_net_round4(at(out, C2, cr, colIndex));
    if (std::fabs(coefficient) > epsilon) {
      int pivotRow = -1;
      for (int r = 0; r < R; r++)
        if (std::fabs(orc_net_round4(at(out, C2, r, colIndex)) - 1.0) <= epsilon) {
          pivotRow = r;
          break;
        }
      if (pivotRow >= 0) {
        for (int col = 0; col < C2; col++) {
          double pivotVal = orc_net_round4(at(out, C2, pivotRow, col));
          double constraintVal = orc_net_round4(at(out, C2, cr, col));
          double newVal = (type == 1) ? pivotVal - coefficient * constraintVal
                                      : constraintVal - coefficient * pivotVal;
          at(out, C2, cr, col) = orc_net_round4(newVal);
        }
      }
    }
  }
  orc_bb_round_tableau((int64_t)R2 * C2, out);  // :799
}
static bool bb_is_integer(double v) {  // :595-599
  double r = orc_net_round4(v);
  return std::fabs(r - orc_net_round(r)) <= 1e-6;
}
/* ExtractSolution :899-921 (same scan as CheckIntegerBasicVar :809-827) */
void orc_bb_extract(int R, int C, const double* T, int n_vars, double* x) {
  for (int i = 0; i < n_vars; i++) {
    x[i] = 0.0;
    for (int j = 0; j < R; j++) {
      double val = orc_net_round4(at(T, C, j, i));
      if (std::fabs(val - 1.0) <= 1e-6) {
        x[i] = orc_net_round4(at(T, C, j, C - 1));
        break;
      }
    }
  }
}
/* CheckIntegerBasicVar :805-857 */
int orc_bb_branch_var(int R, int C, const double* T, int n_vars, double* value) {
  std::vector<double> dv(n_vars);
  orc_bb_extract(R, C, T, n_vars, dv.data());
  int best = -1;
  double minDist = kInf;
  for (int i = 0; i < n_vars; i++) {
    if (!bb_is_integer(dv[i])) {
      double fp = dv[i] - std::floor(dv[i]);
      double d = std::fabs(fp - 0.5);
      if (d < minDist) {
        minDist = d;
        best = i;
        if (value) *value = dv[i];
      }
    }
  }
  return best;
}
/* ExecuteBranchAndBound :1006-1233 */
int orc_bb_solve(int R0, int C0, const double* T0, int n_vars, int enable_pruning,
                 int64_t max_nodes, double* x_out, double* z_out, int* has_solution,
                 int64_t* nodes_processed, int64_t* pivots_total, int* node_log, double* node_z,
                 int64_t node_log_cap) {
  struct Node {
    std::vector<double> T;
    int R, C, depth;
  };
  std::vector<Node> stack;
  {
    Node root{std::vector<double>(T0, T0 + (size_t)R0 * C0), R0, C0, 0};
    orc_bb_round_tableau((int64_t)R0 * C0, root.T.data());  // :1021
    stack.push_back(std::move(root));
  }
  bool have = false;
  double best = -kInf;  // :1024
  std::vector<double> bestx(n_vars, 0.0), sol(n_vars);
  int64_t iteration = 0, processed = 0, pivots = 0;
  int st = ORC_OPTIMAL;
  while (!stack.empty()) {
    iteration++;
    if (max_nodes >= 0 && iteration > max_nodes) {  // :1038-1042
      st = ORC_NODE_LIMIT;
      break;
    }
    Node cur = std::move(stack.back());
    stack.pop_back();
    processed++;
    orc_bb_round_tableau((int64_t)cur.R * cur.C, cur.T.data());  // :1047
    const int R = cur.R, C = cur.C;
    double objVal = orc_net_round4(at(cur.T.data(), C, 0, C - 1));  // GetObjective :892-897
    int lg_var = -1, lg_int = 0, lg_pruned = 0;
    auto write_log = [&]() {
      if (node_log && processed - 1 < node_log_cap) {
        int64_t q = processed - 1;
        node_log[4 * q + 0] = cur.depth;
        node_log[4 * q + 1] = lg_var;
        node_log[4 * q + 2] = lg_int;
        node_log[4 * q + 3] = lg_pruned;
        if (node_z) node_z[q] = objVal;
      }
    };
    if (enable_pruning && have && objVal <= best) {  // ShouldPrunebranch :985-1004
      lg_pruned = 1;
      write_log();
      continue;
    }
    orc_bb_extract(R, C, cur.T.data(), n_vars, sol.data());  // UpdateOptimalSolution :935-983
    bool allint = true;
    for (int i = 0; i < n_vars; i++)
      if (!bb_is_integer(sol[i])) {
        allint = false;
        break;
      }
    if (allint && objVal > best) {
      best = objVal;
      bestx = sol;
      have = true;
    }
    double value = 0.0;
    int var = orc_bb_branch_var(R, C, cur.T.data(), n_vars, &value);  // CreateBranches :859-890
    lg_var = var;
    lg_int = allint ? 1 : 0;
    write_log();
    if (var < 0) continue;
    double lowerInt = (double)(int)std::floor(value), upperInt = (double)(int)std::ceil(value);
    std::vector<Node> children;
    for (int side = 0; side < 2; side++) {  // lower (<= floor) then upper (>= ceil)
      Node ch;
      ch.R = R + 1;
      ch.C = C + 1;
      ch.depth = cur.depth + 1;
      ch.T.resize((size_t)ch.R * ch.C);
      orc_bb_add_constraint(R, C, cur.T.data(), n_vars, var, side == 0 ? lowerInt : upperInt, side,
                            ch.T.data());
      int64_t np = 0;
      int res = orc_bb_node_solve(ch.R, ch.C, ch.T.data(), -1, &np, nullptr, 0);
      pivots += np;
      if (res == ORC_OPTIMAL) {
        orc_bb_round_tableau((int64_t)ch.R * ch.C, ch.T.data());  // :1124 / :1187
        children.push_back(std::move(ch));
      }
    }
    for (int i = (int)children.size() - 1; i >= 0; i--) stack.push_back(std::move(children[i]));  // :1210-1213
  }
  if (has_solution) *has_solution = have ? 1 : 0;
  if (z_out) *z_out = best;
  if (x_out)
    for (int i = 0; i < n_vars; i++) x_out[i] = have ? bestx[i] : 0.0;
  if (nodes_processed) *nodes_processed = processed;
  if (pivots_total) *pivots_total = pivots;
  return st;
}

/* ======================= RevisedPrimalSimplexSolver ======================================= */
/* Simplex/RevisedPrimalSimplexSolver.cs:82-287.  Loops are arranged row-wise for cache
 * friendliness but every scalar result accumulates its terms in the reference's order
 * (s = 0; s += term_0; s += term_1; ...), so results are bit-identical to :398-448. */
int orc_rev_solve(int m, int n, const double* A, const double* b, const double* c_orig,
                  int is_min, int64_t max_iter, int64_t* n_iter, int* basis, double* x_out,
                  double* z_out, double* y_out, double* xB_out, double* Binv_out, int* log,
                  int64_t log_cap) {
  const double EPS = 1e-9;
  std::vector<double> c(n), Binv((size_t)m * m, 0.0), nB((size_t)m * m), cB(m, 0.0), xB(m), y(m),
      rcX(n), u(m);
  std::vector<int> bas(m);
  std::vector<char> isbasic(n + m, 0);
  for (int j = 0; j < n; j++) c[j] = is_min ? -c_orig[j] : c_orig[j];  // :51
  for (int i = 0; i < m; i++) {
    Binv[(size_t)i * m + i] = 1.0;
    bas[i] = n + i;
    isbasic[n + i] = 1;
  }
  int64_t it = 0;
  int st = ORC_RUNNING;
  while (true) {
    for (int i = 0; i < m; i++) {  // xB = B^-1 b  :89, :398-410
      double s = 0;
      const double* row = &Binv[(size_t)i * m];
      for (int j = 0; j < m; j++) s += row[j] * b[j];
      xB[i] = s;
    }
    bool infeas = false;
    for (int i = 0; i < m; i++)
      if (xB[i] < -EPS) infeas = true;
    if (infeas) {  // :90-91
      st = ORC_INFEASIBLE;
      break;
    }
    std::fill(y.begin(), y.end(), 0.0);  // y = cB B^-1  :93, :412-424
    for (int i = 0; i < m; i++) {
      const double* row = &Binv[(size_t)i * m];
      double ci = cB[i];
      for (int j = 0; j < m; j++) y[j] += ci * row[j];
    }
    std::fill(rcX.begin(), rcX.end(), 0.0);  // Dot(y, A[:,j])  :96-98
    for (int i = 0; i < m; i++) {
      const double* row = &A[(size_t)i * n];
      double yi = y[i];
      for (int j = 0; j < n; j++) rcX[j] += yi * row[j];
    }
    for (int j = 0; j < n; j++) rcX[j] = c[j] - rcX[j];
    int enter = -1;  // :104-121
    double bestRC = -kInf;
    for (int v = 0; v < n + m; v++) {
      if (isbasic[v]) continue;
      double rc = (v < n) ? rcX[v] : -y[v - n];
      if (rc > EPS) {
        if (enter == -1 || rc > bestRC + EPS || (std::fabs(rc - bestRC) <= EPS && v < enter)) {
          bestRC = rc;
          enter = v;
        }
      }
    }
    if (enter == -1) {  // :124-146, ExtractSolution :277-287
      st = ORC_OPTIMAL;
      break;
    }
    if (max_iter >= 0 && it >= max_iter) {
      st = ORC_ITER_LIMIT;
      break;
    }
    if (enter < n) {  // :149-151
      for (int i = 0; i < m; i++) {
        double s = 0;
        const double* row = &Binv[(size_t)i * m];
        for (int j = 0; j < m; j++) s += row[j] * A[(size_t)j * n + enter];
        u[i] = s;
      }
    } else {
      for (int i = 0; i < m; i++) u[i] = Binv[(size_t)i * m + (enter - n)];
    }
    int leave = -1;  // :153-176
    double bestRatio = DBL_MAX;
    for (int i = 0; i < m; i++) {
      if (u[i] > EPS) {
        double r = xB[i] / u[i];
        if (r < bestRatio - EPS ||
            (std::fabs(r - bestRatio) <= EPS && (leave == -1 || bas[i] < bas[leave]))) {
          bestRatio = r;
          leave = i;
        }
      }
    }
    if (leave == -1) {  // :178-179
      st = ORC_UNBOUNDED;
      break;
    }
    int leavingVar = bas[leave];
    if (log && it < log_cap) {
      log[3 * it + 0] = leave;
      log[3 * it + 1] = enter;
      log[3 * it + 2] = leavingVar;
    }
    bas[leave] = enter;  // :194-212
    isbasic[enter] = 1;
    isbasic[leavingVar] = 0;
    cB[leave] = (enter < n) ? c[enter] : 0.0;
    double pivot = u[leave];  // UpdateBInverse :264-275 + MultiplyMatrices :426-441
    if (std::fabs(pivot) < EPS) {
      st = ORC_PIVOT_TOO_SMALL;
      break;
    }
    const double* prow = &Binv[(size_t)leave * m];
    for (int i = 0; i < m; i++) {
      double* out = &nB[(size_t)i * m];
      const double* row = &Binv[(size_t)i * m];
      if (i == leave) {
        double e = 1.0 / pivot;
        if (std::fabs(e) < EPS)
          for (int j = 0; j < m; j++) out[j] = 0.0;
        else
          for (int j = 0; j < m; j++) out[j] = 0.0 + e * prow[j];
      } else {
        double e = -u[i] / pivot;
        bool use = !(std::fabs(e) < EPS);
        /* k ascends: k == i contributes 1.0*B[i][j], k == leave contributes e*B[leave][j] */
        if (!use) {
          for (int j = 0; j < m; j++) out[j] = 0.0 + 1.0 * row[j];
        } else if (i < leave) {
          for (int j = 0; j < m; j++) out[j] = (0.0 + 1.0 * row[j]) + e * prow[j];
        } else {
          for (int j = 0; j < m; j++) out[j] = (0.0 + e * prow[j]) + 1.0 * row[j];
        }
      }
    }
    Binv.swap(nB);
    it++;
  }
  if (n_iter) *n_iter = it;
  if (basis) std::copy(bas.begin(), bas.end(), basis);
  if (xB_out) std::copy(xB.begin(), xB.end(), xB_out);
  if (y_out) std::copy(y.begin(), y.end(), y_out);
  if (Binv_out) std::copy(Binv.begin(), Binv.end(), Binv_out);
  if (x_out) {
    for (int j = 0; j < n; j++) x_out[j] = 0.0;
    if (st == ORC_OPTIMAL || st == ORC_ITER_LIMIT)
      for (int i = 0; i < m; i++) {
        int v = bas[i];
        if (v < n) x_out[v] = (0.0 > xB[i]) ? 0.0 : xB[i];  // Math.Max(0.0, xB[i]) :283
      }
    if (z_out) {
      double s = 0;
      for (int j = 0; j < n; j++) s += c_orig[j] * x_out[j];  // :286
      *z_out = s;
    }
  }
  return st;
}

/* ======================= Knapsack ========================================================== */
/* KnapsackBranchBoundSolver.Solve(int,int[],int[]) -- body missing in the reference
 * (IntegerProgramming/KnapsackBranchBoundSolver.cs:9-11); contract from Program.cs:467-470.
 * Classic 0/1 DP over capacities; reconstruction prefers leaving an item out on ties. */
double orc_knap_dp(int capacity, int n, const int* w, const int* v, uint8_t* chosen) {
  if (capacity < 0) capacity = 0;
  std::vector<int64_t> dp((size_t)(n + 1) * (capacity + 1), 0);
  auto D = [&](int i, int cc) -> int64_t& { return dp[(size_t)i * (capacity + 1) + cc]; };
  for (int i = 1; i <= n; i++)
    for (int cc = 0; cc <= capacity; cc++) {
      int64_t bestv = D(i - 1, cc);
      if (w[i - 1] <= cc) {
        int64_t t = D(i - 1, cc - w[i - 1]) + v[i - 1];
        if (t > bestv) bestv = t;
      }
      D(i, cc) = bestv;
    }
  if (chosen) {
    int cc = capacity;
    for (int i = n; i >= 1; i--) {
      chosen[i - 1] = 0;
      if (D(i, cc) != D(i - 1, cc)) {
        chosen[i - 1] = 1;
        cc -= w[i - 1];
      }
    }
  }
  return (double)D(n, capacity);
}

/* The same DP, value only, two rolling rows: O(capacity) memory, so that the arbiter of Program.cs:467-470 can
 * also be run at BASELINE cfg4's size (n = 10^4, capacity ~ 2.5e6: the full table above would need 200 GB). */
double orc_knap_dp_value(int capacity, int n, const int* w, const int* v) {
  if (capacity < 0) capacity = 0;
  std::vector<int64_t> prev((size_t)capacity + 1, 0), next((size_t)capacity + 1, 0);
  for (int i = 0; i < n; i++) {
    const int wi = w[i];
    const int64_t vi = v[i];
    const int64_t* p = prev.data();
    int64_t* q = next.data();
    const int lim = wi <= capacity ? wi : capacity + 1;
    for (int cc = 0; cc < lim; cc++) q[cc] = p[cc];
    for (int cc = wi; cc <= capacity; cc++) {
      const int64_t a = p[cc], t = p[cc - wi] + vi;
      q[cc] = t > a ? t : a;
    }
    prev.swap(next);
  }
  return (double)prev[capacity];
}

/* KnapsackBranchBoundSimplex (Program.cs:444-463 is the only contract).  Specification used
 * by this build (DESIGN.md "Knapsack B&B"): rank items by value/weight descending (ties:
 * lower original id); a node fixes some items to 0/1; its relaxation fills the free items
 * greedily in rank order and stops at the first one that does not fit (the critical item);
 * bound = value + v_k*(cap/w_k); zero residual capacity or no critical item => candidate;
 * incumbent replaced only on strict improvement; nodes with bound <= incumbent fathomed;
 * children: x_k = 0 first, then x_k = 1 (depth first). */
double orc_knap_bb(double capacity, int n, const double* w, const double* v, int64_t max_nodes,
                   uint8_t* chosen, int64_t* nodes_out, int* status) {
  std::vector<int> rank(n);
  std::iota(rank.begin(), rank.end(), 0);
  std::vector<double> ratio(n);
  for (int i = 0; i < n; i++) ratio[i] = v[i] / w[i];
  std::stable_sort(rank.begin(), rank.end(), [&](int a, int b) { return ratio[a] > ratio[b]; });
  struct Node {
    std::vector<int8_t> fix;  // -1 free, 0, 1 by rank position
  };
  std::vector<Node> stack;
  stack.push_back(Node{std::vector<int8_t>(n, -1)});
  double best = -kInf;
  std::vector<uint8_t> bestsel(n, 0), sel(n);
  int64_t nodes = 0;
  int st = ORC_OPTIMAL;
  while (!stack.empty()) {
    if (max_nodes >= 0 && nodes >= max_nodes) {
      st = ORC_NODE_LIMIT;
      break;
    }
    Node nd = std::move(stack.back());
    stack.pop_back();
    nodes++;
    double cap = capacity, val = 0.0;
    for (int p = 0; p < n; p++)
      if (nd.fix[p] == 1) {
        cap -= w[rank[p]];
        val += v[rank[p]];
      }
    if (cap < 0) continue;  // infeasible
    std::fill(sel.begin(), sel.end(), 0);
    int crit = -1;
    for (int p = 0; p < n; p++) {
      if (nd.fix[p] == 1) {
        sel[p] = 1;
        continue;
      }
      if (nd.fix[p] == 0) continue;
      double wi = w[rank[p]];
      if (wi <= cap) {
        cap -= wi;
        val += v[rank[p]];
        sel[p] = 1;
      } else {
        crit = p;
        break;
      }
    }
    /* items fixed to 1 after the critical item are already in val/sel through the first loop;
     * mark the remaining fixed-1 positions */
    if (crit >= 0)
      for (int p = crit + 1; p < n; p++)
        if (nd.fix[p] == 1) sel[p] = 1;
    if (crit < 0 || cap == 0) {  // candidate
      if (val > best) {
        best = val;
        bestsel = sel;
      }
      continue;
    }
    double bound = val + v[rank[crit]] * (cap / w[rank[crit]]);
    if (bound <= best) continue;
    Node one = nd, zero = std::move(nd);
    one.fix[crit] = 1;
    zero.fix[crit] = 0;
    stack.push_back(std::move(one));   // explored second
    stack.push_back(std::move(zero));  // explored first
  }
  if (chosen) {
    for (int i = 0; i < n; i++) chosen[i] = 0;
    if (best > -kInf)
      for (int p = 0; p < n; p++)
        if (bestsel[p]) chosen[rank[p]] = 1;
  }
  if (nodes_out) *nodes_out = nodes;
  if (status) *status = st;
  return best > -kInf ? best : 0.0;
}

}  // extern "C"
