/*
 * include/lprb200.h -- C ABI of liblprb200.so, the B200-native (sm_100a) dense simplex pivot path
 * that sits behind the solver classes of Storm-Tarran/LPR_381_Group_V22.
 *
 * The reference has no FFI today: the drop-in boundary is the public surface of
 * its C# solver classes.  Each entry point below names the reference member whose BODY it
 * replaces (paths relative to LPR_381_Group_V22/); the P/Invoke shim that keeps the C#
 * signatures is in csharp/ and described in INTEGRATION.md.
 *
 * Conventions
 *   - extern "C", cdecl, plain pointers and sizes; no exception crosses the boundary.
 *   - every function returns LPR_OK (0) or a negative LPR_E_* code; the message is available
 *     from lpr_last_error() (thread local).  Solver outcomes (optimal / unbounded / ...) are NOT
 *     errors: they come back in *status.
 *   - host arrays are caller owned, row-major IEEE binary64 (C# double[] / double[,] are
 *     blittable and pinned by the marshaller) and are copied during the call.
 *   - device memory is owned by the opaque handle; *_destroy frees it.
 *   - a handle is not thread safe; different handles may be used from different threads.
 *   - there is NO CPU fallback: without a CUDA device every compute call returns LPR_E_CUDA.
 */
#ifndef LPRB200_H
#define LPRB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LPR_VERSION 100

/* return codes */
#define LPR_OK 0
#define LPR_E_BADARG (-1)
#define LPR_E_CUDA (-2)
#define LPR_E_NOMEM (-3)
#define LPR_E_STATE (-4)
#define LPR_E_CAPACITY (-5)
#define LPR_E_NCCL (-6)

/* solver status (written to *status) */
#define LPR_RUNNING 0
#define LPR_OPTIMAL 1
#define LPR_UNBOUNDED 2
#define LPR_INFEASIBLE 3
#define LPR_ITER_LIMIT 4
#define LPR_NODE_LIMIT 5
#define LPR_PIVOT_TOO_SMALL 6
#define LPR_NO_CUT_NEEDED 7
#define LPR_NO_PIVOT_COL 8
#define LPR_CUT_STEP_DONE 9
#define LPR_DEPTH_LIMIT 10 /* B&B: some subtree was cut at the slab depth headroom (LPR_BB_MAX_DEPTH) */

/* pivot rule ids: the exact selection / tolerance variant */
#define LPR_RULE_PRIMAL 0  /* Simplex/PrimalSimplexSolver.cs:152-211                      */
#define LPR_RULE_PRIMAL2 1 /* Simplex/PrimalSimplexSolver2.cs:102-164                     */
#define LPR_RULE_DUAL 2    /* Simplex/DualSimplex.cs:27-70,150-178                        */
#define LPR_RULE_SENS 3    /* SensitivityAnalysis/SensitivityAnalyzer.cs:98-201           */

/* constraint relation codes (IO/InputFileParser.cs:70-82 Constraint.Relation) */
#define LPR_REL_LE 0
#define LPR_REL_GE 1
#define LPR_REL_EQ 2

typedef struct lpr_tab lpr_tab; /* dense simplex tableau resident in HBM            */
typedef struct lpr_rev lpr_rev; /* revised simplex state (A, B^-1, ...) in HBM      */
typedef struct lpr_bb lpr_bb;   /* branch & bound simplex open-node pool in HBM     */
typedef struct lpr_knap lpr_knap; /* knapsack branch & bound pool in HBM            */

/* ---- library ---------------------------------------------------------------------------- */
int lpr_version(void);
const char* lpr_last_error(void);
int lpr_device_count(int* count);
/* number of kernels this library has launched in this process (bench.py "gpu_launches") */
int64_t lpr_launch_count(void);

/* ---- dense tableau: PrimalSimplexSolver / PrimalSimplexSolver2 / DualSimplexSolver -------- */
/* Replaces the `double[,] tableau` field (PrimalSimplexSolver.cs:12,58; PrimalSimplexSolver2.cs:12,34).
 * rows x cols host array (row 0 = objective row, last column = RHS); row_cap/col_cap >= rows/cols
 * reserve room for appended cut / bound rows and columns (0 = no headroom). */
int lpr_tab_create(int device, int rows, int cols, int row_cap, int col_cap, const double* host,
                   lpr_tab** out);
/* PrimalSimplexSolver..ctor (PrimalSimplexSolver.cs:27-87): builds the (m+1) x (n+m+1) tableau on
 * the device from the model: >= rows negated, = treated as <=, row 0 = -c (c when !is_max), slack
 * identity, basis = slacks.  coef is m x coef_stride, coef_count[i] (NULL => n) entries of row i
 * are valid and only the first n are read (:68-72). */
int lpr_tab_create_primal(int device, int n, int m, const double* objective, const double* coef,
                          int coef_stride, const int* coef_count, const int* relation,
                          const double* rhs, int is_maximization, lpr_tab** out);
/* Synthetic dense LP of SURVEY.md 8(d) generated directly in HBM (counter based splitmix64, bit
 * identical to the oracle's generator): A = 0.1+u, b = (n/4)(1+u), c = 1+u, maximise. */
int lpr_tab_create_dense_lp(int device, uint64_t seed, int m, int n, lpr_tab** out);
int lpr_tab_destroy(lpr_tab* h);
int lpr_tab_dims(const lpr_tab* h, int* rows, int* cols, int* ld);
int lpr_tab_upload(lpr_tab* h, const double* host);      /* rows x cols dense                */
int lpr_tab_read(lpr_tab* h, double* host);              /* GetFinalTableau()  :269-273      */
int lpr_tab_read_row(lpr_tab* h, int row, double* host); /* cols values                      */
int lpr_tab_read_col(lpr_tab* h, int col, double* host); /* rows values                      */
int lpr_tab_get_basis(lpr_tab* h, int* basis);           /* BasicVariables :275-278, rows-1  */
int lpr_tab_set_basis(lpr_tab* h, const int* basis);
/* Solve() loop (PrimalSimplexSolver.cs:102-150; PrimalSimplexSolver2.cs:46-97; DualSimplex.cs:14-114;
 * SensitivityAnalyzer.cs:121-201) entirely on the device.  max_pivots < 0 = no cap.  pivot_log
 * receives (row, col) pairs in tableau indices.  flags bit0 = "printSteps" (only meaningful for
 * RULE_PRIMAL2 / RULE_DUAL whose iteration counters advance only when printing);
 * bit2 = use the unfused reference-shaped kernels (select reads the tableau directly);
 * bit3 = bracket every sweep launch with CUDA events (see lpr_tab_last_sweep_us);
 * bit4 = one tableau sweep per pivot (disable the delayed-update path that applies LPR_TAB_BLOCK=16 pivots
 *        per sweep for RULE_PRIMAL; both paths are bit-identical);
 * bit5 = keep the selection and the sweep of the delayed-update path on one stream (by default the
 *        selection of group g+1 overlaps the out-of-place sweep of group g). */
int lpr_tab_solve(lpr_tab* h, int rule, int64_t max_pivots, int flags, int* status,
                  int64_t* n_pivots, int* pivot_log, int64_t log_cap);
/* one pivot (FindEnteringVariable + FindLeavingVariable + Pivot) for snapshot-accurate tracing */
int lpr_tab_step(lpr_tab* h, int rule, int* enter_col, int* leave_row, int* status);
/* Pivot(row, col) at a caller chosen position (PrimalSimplexSolver.cs:193-211 arithmetic;
 * skip_eps/skip_mode select the |f| skip of the other solvers: 0 none, 1 skip |f|<=eps, 2 skip |f|<eps) */
int lpr_tab_pivot_at(lpr_tab* h, int row, int col, double skip_eps, int skip_mode);
/* ExtractSolution (PrimalSimplexSolver.cs:213-252) for the first n columns */
int lpr_tab_extract_solution(lpr_tab* h, int n, double* x);
/* FinalZ = T[0, cols-1] (:113) */
int lpr_tab_objective(lpr_tab* h, double* z);
/* device time (CUDA events on the handle's stream) and kernel launches of the last solve */
int lpr_tab_last_solve_ms(const lpr_tab* h, float* ms);
/* average duration of the sweep kernel in the last lpr_tab_solve run with flags bit3 set (each sweep
 * launch bracketed by a CUDA event pair on the handle's stream; roofline measurement aid) */
int lpr_tab_last_sweep_us(const lpr_tab* h, float* us);
/* SensitivityAnalysis/SensitivityAnalyzer.cs on a device-resident final tableau:
 * RebuildBasicsFromTableau :706-723 (basis[i-1] = first column that is a unit column with its 1 in row i,
 * else -1); the solution rebuild of ReOptimize :158-164 (x has cols-1 entries, slacks included);
 * AddNewConstraintNonInteractive :609-659 up to ResolveAll: tech has cols-1 entries, rhs_minus_ax =
 * rhs - sum_j tech[j]*solution[j] as the caller computes it (:643-647); the handle needs one row and one
 * column of headroom (row_cap / col_cap of lpr_tab_create), rows and cols grow by one, the RHS stays the
 * last column.  ResolveAll = lpr_tab_sens_rebuild_basis + lpr_tab_solve(rule = LPR_RULE_SENS). */
int lpr_tab_sens_rebuild_basis(lpr_tab* h);
int lpr_tab_sens_solution(lpr_tab* h, double* x);
int lpr_tab_sens_add_constraint(lpr_tab* h, const double* tech, double rhs_minus_ax);
/* append one row (Gomory cut, CuttingPlaneSolver.cs:110) -- needs row headroom */
int lpr_tab_append_row(lpr_tab* h, const double* row);
/* Gomory fractional cut rows 1-4 of CuttingPlaneSolver.cs:76-107 generated on the device:
 * returns the chosen constraint row (0-based among constraint rows, -1 = none) and the cut. */
int lpr_tab_gomory_cut(lpr_tab* h, int* chosen_row, double* cut_host /* cols, may be NULL */,
                       int append);
/* CuttingPlaneSolver.CuttingPlaneSolution (CuttingPlaneSolver.cs:64-229), recursion as a loop;
 * cut_log: (chosen_row, pivot_col, n_dual_pivots, n_primal_pivots) per cut */
int lpr_tab_cutting_plane(lpr_tab* h, int max_cuts, int* status, int* n_cuts, int* cut_log,
                          int cut_log_cap);

/* ---- RevisedPrimalSimplexSolver (Simplex/RevisedPrimalSimplexSolver.cs) -------------------- */
/* ctor :41-80 (Relation ignored: every row is <=).  A is m x n row-major. */
int lpr_rev_create(int device, int m, int n, const double* A, const double* b, const double* c,
                   int is_minimization, lpr_rev** out);
int lpr_rev_create_dense_lp(int device, uint64_t seed, int m, int n, lpr_rev** out);
int lpr_rev_destroy(lpr_rev* h);
/* Solve() :82-251.  refactor_every > 0 recomputes B^-1 from the basis columns every that many
 * iterations (FP64 tensor-core GEMM based inversion; the reference never refactorises, Q7).
 * log: (leaveRow, enter, leaveVar) triples.  status: OPTIMAL / INFEASIBLE (:91) / UNBOUNDED (:179)
 * / PIVOT_TOO_SMALL (:267) / ITER_LIMIT. */
int lpr_rev_solve(lpr_rev* h, int64_t max_iter, int refactor_every, int* status, int64_t* n_iter,
                  int* log, int64_t log_cap);
int lpr_rev_refactor(lpr_rev* h);
/* One iteration at a time, for snapshot-accurate tracing (the C# shim / the Python mirror use this path below a size
 * threshold, lpr_rev_solve above it): lpr_rev_begin resets to the slack basis; lpr_rev_step runs one pass of the loop
 * of Solve() :86-250 -- *status RUNNING after a pivot, OPTIMAL when none is left (x and z are then ready), INFEASIBLE /
 * UNBOUNDED / PIVOT_TOO_SMALL where the reference throws; lpr_rev_format_snapshot returns the text block
 * CaptureSnapshot (:294-387) builds for that step ("Iteration k" / "Optimal": post-pivot duals, reduced costs and
 * B^-1 A | B^-1 | RHS table beside the pre-pivot direction and ratio test).  B^-1 A (:360, O(m^2 n)) is only computed
 * when this function is called.  *text is library owned, valid until the thread's next formatting call. */
int lpr_rev_begin(lpr_rev* h);
int lpr_rev_step(lpr_rev* h, int* status, int* enter, int* leave_row, int* leave_var);
int lpr_rev_format_snapshot(lpr_rev* h, const char** text, int64_t* len);
int lpr_rev_read_basis(lpr_rev* h, int* basis);   /* BasicVariables :39, m entries          */
int lpr_rev_read_x(lpr_rev* h, double* x);        /* SolutionVector :277-287, n entries     */
int lpr_rev_read_z(lpr_rev* h, double* z);        /* FinalZ :286                            */
int lpr_rev_read_y(lpr_rev* h, double* y);        /* dual prices y = c_B B^-1 :93, m        */
int lpr_rev_read_xb(lpr_rev* h, double* xb);      /* x_B = B^-1 b :89, m                    */
int lpr_rev_read_binv(lpr_rev* h, double* binv);  /* m x m                                  */
int lpr_rev_last_solve_ms(const lpr_rev* h, float* ms);
int lpr_rev_last_refactor_ms(const lpr_rev* h, float* ms);
/* ||I - B X||_inf before the refactorisation (of the new inverse when mode 2 was asked for) and its FP64 flop count */
int lpr_rev_last_refactor_info(const lpr_rev* h, double* residual, double* flops);
/* Refactorisation with an explicit mode.  0 (what lpr_rev_refactor and refactor_every use): Newton-Schulz refresh
 * X <- X + X (I - B X) while the current inverse is good (||I - B X||_inf < 1e-5: one step then reaches working
 * precision), otherwise the full path; 1: refresh
 * only; 2: full refactorisation from the basis columns alone -- blocked Gauss-Jordan inversion with partial pivoting,
 * rank-64 trailing updates on the FP64 tensor cores, residual check and one polishing step.  lpr_rev_last_refactor_path:
 * which path ran (1 refresh, 2 full) and ||I - B X||_inf of the result of the full path (before its polishing step). */
int lpr_rev_refactor_ex(lpr_rev* h, int mode);
int lpr_rev_last_refactor_path(const lpr_rev* h, int* path, double* residual_after);
int lpr_rev_write_binv(lpr_rev* h, const double* binv); /* overwrite B^-1 (m x m row major) */

/* ---- BranchBoundSimplexSolver (IntegerProgramming/BranchBoundSimplexSolver.cs) ------------- */
/* building blocks, each on a device tableau */
int lpr_tab_round4(lpr_tab* h);                                  /* RoundTableau :552-567      */
int lpr_tab_bb_node_solve(lpr_tab* h, int64_t max_pivots, int* status, int64_t* n_pivots,
                          int* pivot_log, int64_t log_cap);     /* DoDualSimplex :289-468     */
/* the same with DoDualSimplex's isMinimization argument (optimal when the objective row is <= 0, entering =
 * smallest positive entry :209-213, :346-348); only RunBranchAndBound's initial solve :1271 passes true */
int lpr_tab_bb_node_solve_ex(lpr_tab* h, int is_minimization, int64_t max_pivots, int* status,
                             int64_t* n_pivots, int* pivot_log, int64_t log_cap);
/* DualSimplexSolverBB.FormulateTableau / PrepareInput (:28-113, :281-287) built on the device: cons is m rows
 * of `stride` doubles, row i holds len[i] entries [coefficients..., rhs, type flag] (flag 1 = ">=": the row is
 * negated; rows longer than n+2, like the ones ConfigureProblem :1233-1251 appends, spill into the slack
 * columns exactly as in the reference).  The tableau is (m+1) x (n+m+1). */
int lpr_tab_create_bb(int device, int n, int m, const double* objective, const double* cons, int stride,
                      const int* len, int row_cap, int col_cap, lpr_tab** out);
int lpr_tab_bb_add_constraint(lpr_tab* parent, int n_vars, int var, double bound, int type,
                              lpr_tab** child);                  /* AddConstraint :694-803     */
int lpr_tab_bb_branch_var(lpr_tab* h, int n_vars, int* var, double* value,
                          double* x /* n_vars, may be NULL */); /* :805-857, :899-921          */
/* BranchAndBoundAdapter.SolveFromPrimal + BranchAndBound.ExecuteBranchAndBound (:1006-1233):
 * depth-first tree on ONE device, reference order (used for parity; max_nodes = 20 is the
 * reference cap, < 0 lifts it).  node_log: (depth, branch_var, is_integer, pruned) per node. */
int lpr_bb_solve(int device, int rows, int cols, const double* final_tableau, int n_vars,
                 int enable_pruning, int64_t max_nodes, double* x, double* z, int* has_solution,
                 int64_t* nodes, int64_t* pivots, int* node_log, double* node_z,
                 int64_t node_log_cap, int* status);
/* Partitionable open-node pool for the multi-GPU driver (one process per GPU; incumbents and
 * node records are exchanged by the host layer over NCCL): see DESIGN.md "B&B pool". */
int lpr_bb_create(int device, int rows, int cols, const double* root_tableau, int n_vars,
                  int enable_pruning, lpr_bb** out);
int lpr_bb_destroy(lpr_bb* h);
int lpr_bb_open_count(lpr_bb* h, int64_t* n);
/* expand up to max_nodes open nodes (deepest first); returns nodes processed / pivots done */
int lpr_bb_run(lpr_bb* h, int64_t max_nodes, int64_t* processed, int64_t* pivots);
/* same with a time slice: stops after the batch during which max_seconds (> 0) of host time have elapsed.  The
 * multi-GPU driver runs time-sliced rounds so that ranks whose subtrees need more pivots per node do not hold the
 * others at the incumbent exchange (the result does not depend on where the rounds are cut). */
int lpr_bb_run_timed(lpr_bb* h, int64_t max_nodes, double max_seconds, int64_t* processed, int64_t* pivots);
/* keep the open nodes whose position in the stack is == offset (mod stride) and drop the others: after every rank
 * has expanded the same root for a few batches (bit-identical pools), each keeps its own share -- a start-up
 * partition without any transfer (lpr_381_group_v22_b200/distributed.py, replicated_root) */
int lpr_bb_keep_stride(lpr_bb* h, int offset, int stride);
/* totals since creation; depth_overflow counts nodes whose children were NOT generated because they would
 * exceed the slab depth headroom (LPR_BB_MAX_DEPTH, default 128) -- a non-zero value means the search was cut */
int lpr_bb_stats(lpr_bb* h, int64_t* processed, int64_t* pivots, int64_t* depth_overflow, int* max_depth);
/* incumbent as (z, dfs_key[], x[]) -- the key makes ties deterministic across GPU counts */
int lpr_bb_get_incumbent(lpr_bb* h, int* has, double* z, double* x, int* key, int* key_len);
int lpr_bb_set_incumbent(lpr_bb* h, double z, const double* x, const int* key, int key_len);
/* work stealing: pop up to max_nodes shallowest open nodes into a host byte buffer / push them */
int lpr_bb_export_nodes(lpr_bb* h, int max_nodes, void* buf, int64_t buf_cap, int64_t* bytes,
                        int* n_exported);
int lpr_bb_import_nodes(lpr_bb* h, const void* buf, int64_t bytes);

/* ---- Knapsack branch & bound (Program.cs:430-471; KnapsackBranchBoundSimplex is missing from
 *      the reference, specification in DESIGN.md) ----------------------------------------------- */
/* KnapsackBranchBoundSolver.Solve(int,int[],int[]) -- DP arbiter, on the device */
int lpr_knap_dp(int device, int capacity, int n, const int* weights, const int* values,
                double* best, uint8_t* chosen);
int lpr_knap_create(int device, double capacity, int n, const double* weights,
                    const double* values, lpr_knap** out);
int lpr_knap_destroy(lpr_knap* h);
/* expand the pool level by level (batches of the deepest nodes) until it is empty or max_nodes is reached; the
 * levels run back to back on the device, the host only looks at a control block every few levels */
int lpr_knap_run(lpr_knap* h, int64_t max_nodes, int64_t* processed, int* status);
/* same with a time slice: returns after the group of tree levels during which max_seconds (> 0) have elapsed */
int lpr_knap_run_timed(lpr_knap* h, int64_t max_nodes, double max_seconds, int64_t* processed, int* status);
int lpr_knap_open_count(lpr_knap* h, int64_t* n);
/* keep the open nodes at stack positions == offset (mod stride), drop the others (start-up partition of the multi-GPU
 * driver: every rank expands the same root identically, then keeps its share; same idea as lpr_bb_keep_stride) */
int lpr_knap_keep_stride(lpr_knap* h, int offset, int stride);
int lpr_knap_get_incumbent(lpr_knap* h, double* best, uint8_t* chosen /* n, original ids */,
                           uint64_t* key /* key_words */, int* key_bits);
int lpr_knap_set_incumbent(lpr_knap* h, double best, const uint8_t* chosen, const uint64_t* key,
                           int key_bits);
int lpr_knap_export_nodes(lpr_knap* h, int max_nodes, void* buf, int64_t buf_cap, int64_t* bytes,
                          int* n_exported);
int lpr_knap_import_nodes(lpr_knap* h, const void* buf, int64_t bytes);
/* KnapsackBranchBoundSimplex.Solve() single device convenience */
int lpr_knap_solve(int device, double capacity, int n, const double* weights,
                   const double* values, int64_t max_nodes, double* best, uint8_t* chosen,
                   int64_t* nodes, int* status);

/* ---- multi-GPU branch & bound inside the library ----------------------------------------
 * The reference's callers are single-process (Program.cs:385-389 B&B simplex, :444-468 knapsack), so the node pool
 * is partitioned over n_gpus devices of the box BY THE LIBRARY: one host thread per device, one NCCL communicator
 * per device (ncclCommInitAll, cached), per round one ncclAllReduce(MAX) of [incumbent value, changed flag, counts],
 * work stealing with ncclSend / ncclRecv of node records device to device, then one time slice of node work
 * (csrc/multi_gpu.cu).  devices NULL = 0 .. n_gpus-1.  The incumbent (value, then DFS-first key) does not depend on
 * n_gpus.  n_gpus = 1 needs no NCCL; with n_gpus > 1 a missing libnccl.so.2 is LPR_E_NCCL (no fallback). */
typedef struct lpr_mgpu_stats {
  int n_gpus, nccl_version;
  int ranks_per_gpu, reserved; /* pools (host thread + stream each) per device: LPR_MG_RANKS_PER_GPU, default 2 for
                                  B&B simplex (child construction of one pool overlaps the pivot chains of the other) */
  int64_t rounds, steals, nodes_moved, open_left, depth_overflow;
  double seconds;       /* seeding + rounds, max over the ranks */
  double setup_seconds; /* pool creation, NCCL communicator and channel set-up */
  double seed_seconds, exchange_seconds, steal_seconds; /* rank 0 */
  int64_t nodes_per_gpu[16];
  double run_seconds_per_gpu[16];
} lpr_mgpu_stats;
/* BranchAndBound.ExecuteBranchAndBound over n_gpus devices.  max_nodes: total node budget (< 0 none); max_rounds:
 * number of rounds (< 0 until the pools are empty); slice_seconds > 0: a round is a time slice of that length.
 * status: OPTIMAL (tree closed), NODE_LIMIT (budget / rounds spent, open nodes left), DEPTH_LIMIT. */
int lpr_bb_solve_mgpu(int n_gpus, const int* devices, int rows, int cols, const double* final_tableau, int n_vars,
                      int enable_pruning, int64_t max_nodes, int64_t max_rounds, double slice_seconds, double* x,
                      double* z, int* has_solution, int64_t* nodes, int64_t* pivots, int* status,
                      lpr_mgpu_stats* stats /* may be NULL */);
/* KnapsackBranchBoundSimplex.Solve() over n_gpus devices (slice_seconds <= 0: 2 ms) */
int lpr_knap_solve_mgpu(int n_gpus, const int* devices, double capacity, int n, const double* weights,
                        const double* values, int64_t max_nodes, int64_t max_rounds, double slice_seconds, double* best,
                        uint8_t* chosen, int64_t* nodes, int* status, lpr_mgpu_stats* stats /* may be NULL */);
int lpr_nccl_version(int* version); /* the NCCL the library resolved (dlopen), e.g. 22809 */

/* ---- model input: IO/InputFileParser.cs:19-68 and the CLI's extra rows ---------------- */
typedef struct lpr_model lpr_model;
/* InputFileParser.ReadInputFile (:19-68).  A missing file or fewer than three lines is not an error: like the
 * reference the parser prints a message (lpr_model_message) and stays empty (lpr_model_info: loaded = 0).  What
 * throws in the reference -- double.Parse FormatException, a constraint line with fewer than n+2 tokens
 * (IndexOutOfRangeException) -- returns LPR_E_BADARG with the exception's name at the start of the message. */
int lpr_model_parse_file(const char* path, lpr_model** out);
int lpr_model_parse_text(const char* text, int64_t len, lpr_model** out);
/* dense model from arrays (coef m x n row-major; relation NULL = all <=): the synthetic configs */
int lpr_model_from_dense(int n, int m, const double* objective, const double* coef, const int* relation,
                         const double* rhs, int is_maximization, lpr_model** out);
int lpr_model_destroy(lpr_model* m);
/* ObjectiveCoefficients.Count, Constraints.Count, SignRestrictions.Count (:12-15) */
int lpr_model_info(const lpr_model* m, int* loaded, int* n, int* n_constraints, int* n_signs);
int lpr_model_problem_type(const lpr_model* m, char* out, int cap); /* ProblemType :12, lower-cased :36 */
int lpr_model_message(const lpr_model* m, char* out, int cap);      /* the console line of :23,:32,:66 */
int lpr_model_objective(const lpr_model* m, double* c);
/* Constraints[i] (:70-82): Coefficients (count entries), Relation string, RHS */
int lpr_model_constraint(const lpr_model* m, int i, double* coef, int cap, int* count, char* relation,
                         int rel_cap, double* rhs);
int lpr_model_sign(const lpr_model* m, int j, char* out, int cap); /* SignRestrictions[j] :63-64 */
/* Program.cs:114-124 / :372-382: menu options 1 and 3 append `x_i <= 1` rows of length n+3 with a stray 1 */
int lpr_model_add_cli_bound_rows(lpr_model* m);
/* Program.cs:511-535 AddUpperBoundConstraints (menu option 2): `x_j <= 1` for "bin" / "<=1" restrictions */
int lpr_model_add_upper_bound_rows(lpr_model* m);
/* PrimalSimplexSolver..ctor on the parsed model: the tableau is built on the device from the model's arrays
 * (no List<Constraint> -> double[,] -> upload detour) */
int lpr_tab_create_from_model(int device, const lpr_model* m, int is_maximization, lpr_tab** out);
/* dense binary model file (header + raw doubles), so the synthetic configs need no text round trip */
int lpr_model_save_binary(const lpr_model* m, const char* path);
int lpr_model_load_binary(const char* path, lpr_model** out);

/* ---- snapshots: Utilities/TableIterationFormater.cs:22-48, NumFormat.N3 ------------------ */
/* $"{x:F3}" and NumFormat.N3 (RevisedPrimalSimplexSolver.cs:455-465) with .NET Framework semantics: 15
 * significant digits first, then half-away-from-zero on the digit string; a negative that rounds to zero loses
 * its sign */
int lpr_fmt_f3(double x, char* out, int cap);
int lpr_fmt_n3(double x, char* out, int cap);
/* $"{x:F<decimals>}" with the same rules (e.g. the F6 of the "Final Tableau (Optimal)" summary) */
int lpr_fmt_fixed(double x, int decimals, char* out, int cap);
/* TableIterationFormater.Format(tab, numOriginalVars, title, rowLabels) on a host array (ld doubles per row).
 * *text points to a buffer owned by the library, valid until the calling thread's next lpr_fmt_table /
 * lpr_tab_format call. */
int lpr_fmt_table(const double* tab, int rows, int cols, int64_t ld, int num_original_vars,
                  const char* title, const char* const* row_labels, int n_labels, const char** text,
                  int64_t* len);
/* the same for a device tableau: rows stream D2H in ~8 MB blocks through two pinned buffers while host threads
 * format the previous block (replaces the IterationSnapshots.Add(Format(...)) calls of
 * PrimalSimplexSolver.cs:86,:148,:122) */
int lpr_tab_format(lpr_tab* h, int num_original_vars, const char* title, const char* const* row_labels,
                   int n_labels, const char** text, int64_t* len);

/* ---- result files: IO/OutputFileWrite.cs:16-137, CanonicalFormConverter.cs:57-98 ------------ */
int lpr_fmt_general(double x, char* out, int cap); /* double.ToString(): 15 significant digits, "G" */
/* CanonicalFormConverter.CanonicalFormForFile on the model; *text is library owned like lpr_fmt_table's */
int lpr_model_canonical_form(const lpr_model* m, const char** text, int64_t* len);
/* OutputFileWrite.WriteFullResults (:16-78): header, canonical form, snapshots, Z* and x (NumFormat.N3).  x may be
 * NULL / n_x 0 (solutionVector == null).  timestamp NULL = local time now ("yyyy-MM-dd HH:mm:ss").  The file gets
 * a UTF-8 byte-order mark when it starts empty, like File.WriteAllText(..., Encoding.UTF8). */
int lpr_out_write_full_results(const char* path, const char* solver_used, const lpr_model* m,
                               const char* const* snapshots, int n_snapshots, double final_z, const double* x,
                               int n_x, int append, const char* timestamp);
/* OutputFileWrite.WriteSnapshotsOnly (:83-119) */
int lpr_out_write_snapshots_only(const char* path, const char* solver_used, const char* const* snapshots,
                                 int n_snapshots, double final_z, const double* x, int n_x, int append,
                                 const char* timestamp);

#ifdef __cplusplus
}
#endif
#endif /* LPRB200_H */
