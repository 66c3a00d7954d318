/*
 * oracle/lpr_oracle.h -- CPU oracle for the LPR_381_Group_V22 dense simplex pivot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may load it,
 * and there only as the checker or the timed CPU baseline.  The product path
 * (lpr_381_group_v22_b200 + liblprb200.so) never links, imports or calls this file.
 *
 * What it is: a plain C++17 restatement (scalar double loops, same loop order, same
 * tolerances, IEEE binary64, round-to-nearest-even, NO FMA contraction: build with
 * -ffp-contract=off) of the C# solver loops of Storm-Tarran/LPR_381_Group_V22.  Every
 * function cites the reference file:line it follows (paths relative to
 * /root/reference/LPR_381_Group_V22/).
 *
 * Parity pinning: the reference ships NO tests, golden vectors or expected outputs
 *  and cannot be compiled here (no .NET toolchain; Program.cs:444,468 do not
 * compile anyway).  The oracle is therefore pinned to the reference EXECUTED by other means:
 * (0) oracle/csharp/ is a C# interpreter that runs the reference's own, unmodified .cs files in
 * the build container; tests/golden/make_reference_run.py records what every solver class
 * returned on the reference's fixtures and on seeded models (tests/golden/reference_run.json)
 * and tests/test_reference_run.py compares every function below with that, bit for bit;
 * (1) the only result check the reference has -- knapsack B&B value == DP value,
 * Program.cs:467-470 (the knapsack bodies are missing upstream, so there is nothing to execute);
 * (2) the reference's shipped fixtures (data/TextFile.txt, README model, Program.cs:433-435)
 * whose known answers were derived independently in SURVEY.md Appendix C and are reproduced
 * bit-for-bit by tests/test_oracle_golden.py; (3) scipy HiGHS objective cross-checks.
 * What (0) does not give: output of the CLR itself -- the interpreter restates the C# semantics
 * and the BCL rules it needs (listed in oracle/csharp/csrun.py); the arithmetic is the same
 * IEEE binary64 in both.
 */
#ifndef LPR_ORACLE_H
#define LPR_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* status codes shared with include/lprb200.h */
#define ORC_RUNNING 0
#define ORC_OPTIMAL 1
#define ORC_UNBOUNDED 2
#define ORC_INFEASIBLE 3
#define ORC_ITER_LIMIT 4
#define ORC_NODE_LIMIT 5
#define ORC_PIVOT_TOO_SMALL 6
#define ORC_NO_CUT_NEEDED 7
#define ORC_NO_PIVOT_COL 8
#define ORC_CUT_STEP_DONE 9

/* rule ids shared with include/lprb200.h */
#define ORC_RULE_PRIMAL 0  /* Simplex/PrimalSimplexSolver.cs          */
#define ORC_RULE_PRIMAL2 1 /* Simplex/PrimalSimplexSolver2.cs         */
#define ORC_RULE_DUAL 2    /* Simplex/DualSimplex.cs                  */
#define ORC_RULE_SENS 3    /* SensitivityAnalysis/SensitivityAnalyzer */

/* ---- synthetic generator: u(seed,k) in [0,1) ---------------------------- */
uint64_t orc_splitmix64(uint64_t x);
double orc_u01(uint64_t seed, uint64_t k);
/* cfg2/cfg3 dense LP: A (m x n) = 0.1+u, b = (n/4)(1+u), c = 1+u */
void orc_gen_dense_lp(uint64_t seed, int m, int n, double* A, double* b, double* c);
/* cfg5 dense IP: A = 1+floor(20u), b = floor(rowsum/4), c = 1+floor(30u) */
void orc_gen_dense_ip(uint64_t seed, int m, int n, double* A, double* b, double* c);
/* cfg4 knapsack: w = 1+floor(1000u), v = max(1, w+floor(200u')-100), cap = floor(sum w/2) */
void orc_gen_knapsack(uint64_t seed, int n, double* w, double* v, double* capacity);

/* ---- Math.Round restatements ------------------------------------------------------------ */
double orc_net_round(double x);  /* Math.Round(double), banker's */
double orc_net_round4(double x); /* Math.Round(x, 4)             */
double orc_frac(double x);       /* CuttingPlaneSolver.Frac      */

/* ---- PrimalSimplexSolver (Simplex/PrimalSimplexSolver.cs) ------------------------------- */
/* ctor :27-87.  coef is m x coef_stride, coef_count[i] entries valid in row i;
 * relation: 0 "<=", 1 ">=", 2 "=" ; T is (m+1) x (n+m+1) row-major; basis has m entries. */
void orc_primal_build(int n, int m, const double* objective, const double* coef, int coef_stride,
                      const int* coef_count, const int* relation, const double* rhs,
                      int is_maximization, double* T, int* basis);
/* Solve :102-150 with FindEnteringVariable :152-167, FindLeavingVariable :169-191,
 * Pivot :193-211.  pivot_log holds (row, col) pairs; threads>1 parallelises the row loop of
 * Pivot only (element-wise, bit-identical).  max_pivots<0 = no cap (the reference has none). */
int orc_primal_solve(int R, int C, double* T, int* basis, int64_t max_pivots, int* status,
                     int64_t* n_pivots, int* pivot_log, int64_t log_cap, int threads);
int orc_primal_find_entering(int R, int C, const double* T);
int orc_primal_find_leaving(int R, int C, const double* T, int col);
void orc_primal_pivot(int R, int C, double* T, int prow, int pcol, int threads);
/* ExtractSolution :213-252 */
void orc_primal_extract(int R, int C, int n, const double* T, double* x);

/* ---- PrimalSimplexSolver2 (Simplex/PrimalSimplexSolver2.cs:46-164) ---------------------- */
int orc_primal2_solve(int R, int C, double* T, int max_iters, int print_steps, int* status,
                      int64_t* n_pivots, int* pivot_log, int64_t log_cap);
/* ---- DualSimplexSolver (Simplex/DualSimplex.cs:14-114,150-178) -------------------------- */
int orc_dual_solve(int R, int C, double* T, int max_iters, int print_steps, int* status,
                   int64_t* n_pivots, int* pivot_log, int64_t log_cap);
/* ---- SensitivityAnalyzer re-optimisation (SensitivityAnalyzer.cs:98-201) ---------------- */
int orc_sens_resolve(int R, int C, double* T, int* basis, int max_iter, int* status,
                     int64_t* n_pivots, int* pivot_log, int64_t log_cap);

/* RebuildBasicsFromTableau (SensitivityAnalyzer.cs:706-723) with GetBasicRow/IsPivotColumn (:64-83):
 * basis[i-1] = first column j < C-1 with |T[i,j]-1| < 1e-9 whose other constraint rows are all
 * within 1e-9 of 0, else -1 */
void orc_sens_rebuild_basis(int R, int C, const double* T, int* basis);
/* solution rebuild of ReOptimize (:158-164): x[j] = RHS of GetBasicRow(j) or 0, j < C-1 */
void orc_sens_solution(int R, int C, const double* T, double* x);
/* AddNewConstraintNonInteractive (:609-659) up to, not including, ResolveAll: Tout is (R+1) x (C+1);
 * tech has C-1 entries; rhs_minus_ax = rhs - sum_j tech[j]*solution[j] (computed by the caller as
 * the reference does, :643-647); basis gets the new slack column appended (R entries on return) */
void orc_sens_add_constraint(int R, int C, const double* T, int* basis, const double* tech,
                             double rhs_minus_ax, double* Tout);

/* ---- CuttingPlaneSolver (IntegerProgramming/CuttingPlaneSolver.cs:64-229) --------------- */
/* T has capacity row_cap x C; *R in/out (row 0 = objective).  max_cuts<0 = unlimited
 * (reference recursion is unbounded).  cut_log: per cut (chosen_row, pivot_col, n_dual, n_primal). */
int orc_cutting_plane(int* R, int C, double* T, int row_cap, int max_cuts, int* status,
                      int* n_cuts, int* cut_log, int cut_log_cap);
/* one cut row only (steps 1-4, :76-107): returns chosen constraint row (0-based among
 * constraint rows) or -1; cut has C entries */
int orc_gomory_cut(int R, int C, const double* T, double* cut);
/* the row choice above follows List<T>.Sort literally (the Framework's unstable introspective sort); 1 = take the
 * plain first minimum instead (what the CUDA kernels do: identical unless more than 16 fractional rows tie exactly
 * for the best key).  orc_gomory_tie_corners() counts the selections in which the two rules disagreed. */
void orc_set_gomory_first_min(int on);
int64_t orc_gomory_tie_corners(void);

/* ---- BranchBoundSimplexSolver (IntegerProgramming/BranchBoundSimplexSolver.cs) ---------- */
void orc_bb_round_tableau(int64_t count, double* T); /* RoundTableau :552-567 */
/* PerformDualPivot :115-201. out may alias nothing (out-of-place). returns 1 pivoted, 0 none/infeasible */
int orc_bb_dual_pivot(int R, int C, const double* T, double* out, int* prow, int* pcol);
/* PerformPrimalPivot :203-279 (isMinimization=false) */
int orc_bb_primal_pivot(int R, int C, const double* T, double* out, int* prow, int* pcol);
/* DoDualSimplex with tableauOverride :289-468.  T in/out (final tableau returned);
 * returns ORC_OPTIMAL (optimalValue != null) / ORC_INFEASIBLE (null or exception). */
int orc_bb_node_solve(int R, int C, double* T, int64_t max_pivots, int64_t* n_pivots,
                      int* pivot_log, int64_t log_cap);
int orc_bb_node_solve_ex(int R, int C, double* T, int is_min, int64_t max_pivots, int64_t* n_pivots,
                         int* pivot_log, int64_t log_cap);
/* FormulateTableau :28-113 (ragged rows [coefficients..., rhs, type flag]); T is (m+1) x (n+m+1) */
void orc_bb_formulate(int n, int m, const double* objective, const double* cons, int stride, const int* len, double* T);
/* IdentifyBasicVariables :642-692 (on an already rounded tableau). returns count */
int orc_bb_identify_basic(int R, int C, const double* T, int* basic);
/* AddConstraint :694-803 with one new constraint e_var (<= if type==0, >= if type==1).
 * base is R x C; out is (R+1) x (C+1). */
void orc_bb_add_constraint(int R, int C, const double* base, int n_vars, int var, double bound,
                           int type, double* out);
/* CheckIntegerBasicVar :805-857: returns branch var or -1; *value = its value */
int orc_bb_branch_var(int R, int C, const double* T, int n_vars, double* value);
/* ExtractSolution :899-921 */
void orc_bb_extract(int R, int C, const double* T, int n_vars, double* x);
/* ExecuteBranchAndBound :1006-1233 via BranchAndBoundAdapter.SolveFromPrimal.
 * max_nodes = 20 is the reference cap (:1038); <0 = unlimited.  node_log (optional):
 * per processed node 4 ints (depth, branch_var or -1, is_integer, pruned) and node_z its z. */
int orc_bb_solve(int R, int C, const double* T0, int n_vars, int enable_pruning,
                 int64_t max_nodes, double* x_out, double* z_out, int* has_solution,
                 int64_t* nodes_processed, int64_t* pivots_total, int* node_log, double* node_z,
                 int64_t node_log_cap);

/* ---- RevisedPrimalSimplexSolver (Simplex/RevisedPrimalSimplexSolver.cs:82-287) ---------- */
/* A m x n row-major (Relation ignored :55-61).  log: (leaveRow, enter, leaveVar) triples.
 * returns ORC_OPTIMAL / ORC_INFEASIBLE (":91") / ORC_UNBOUNDED (":179") /
 * ORC_PIVOT_TOO_SMALL (":267") / ORC_ITER_LIMIT */
int orc_rev_solve(int m, int n, const double* A, const double* b, const double* c_orig,
                  int is_min, int64_t max_iter, int64_t* n_iter, int* basis, double* x,
                  double* z, double* y, double* xB, double* Binv, int* log, int64_t log_cap);

/* ---- Knapsack (Program.cs:430-471; bodies missing in the reference => spec in DESIGN.md) - */
/* DP arbiter: KnapsackBranchBoundSolver.Solve(int,int[],int[]) */
double orc_knap_dp(int capacity, int n, const int* weights, const int* values, uint8_t* chosen);
/* value only, O(capacity) memory (cfg4 size) */
double orc_knap_dp_value(int capacity, int n, const int* weights, const int* values);
/* B&B: ratio-ranked fractional bound, branch on the fractional item x=0 then x=1, DFS,
 * strict-improvement incumbent.  chosen[i] over ORIGINAL ids. */
double orc_knap_bb(double capacity, int n, const double* weights, const double* values,
                   int64_t max_nodes, uint8_t* chosen, int64_t* nodes, int* status);

#ifdef __cplusplus
}
#endif
#endif
