// LprNative.cs -- P/Invoke binding of liblprb200 (include/lprb200.h).  Drop into the reference project
// (LPR_381_Group_V22/Native/); cannot be compiled in the build image (no .NET toolchain); executed by the interpreter of oracle/csharp -- see INTEGRATION.md.
using System;
using System.Runtime.InteropServices;

namespace LPR_381_Group_V22.Native
{
    internal static class Lpr
    {
        private const string Lib = "lprb200"; // liblprb200.so / lprb200.dll next to the executable

        public const int OK = 0;
        public const int RUNNING = 0, OPTIMAL = 1, UNBOUNDED = 2, INFEASIBLE = 3, ITER_LIMIT = 4, NODE_LIMIT = 5,
                         PIVOT_TOO_SMALL = 6, NO_CUT_NEEDED = 7, NO_PIVOT_COL = 8, CUT_STEP_DONE = 9, DEPTH_LIMIT = 10;
        public const int RULE_PRIMAL = 0, RULE_PRIMAL2 = 1, RULE_DUAL = 2, RULE_SENS = 3;

        [DllImport(Lib, CallingConvention = CallingConvention.Cdecl)] public static extern int lpr_version();
        [DllImport(Lib, CallingConvention = CallingConvention.Cdecl)] public static extern IntPtr lpr_last_error();

        // ---- dense tableau ---------------------------------------------------------------------
        [DllImport(Lib, CallingConvention = CallingConvention.Cdecl)]
        public static extern int lpr_tab_create(int device, int rows, int cols, int rowCap, int colCap, double[,] host, out IntPtr h);
        [DllImport(Lib, CallingConvention = CallingConvention.Cdecl)]
        public static extern int lpr_tab_create_primal(int device, int n, int m, double[] objective, double[] coef, int coefStride,
                                                       int[] coefCount, int[] relation, double[] rhs, int isMax, out IntPtr h);
        [DllImport(Lib, CallingConvention = CallingConvention.Cdecl)] public static extern int lpr_tab_destroy(IntPtr h);
        [DllImport(Lib, CallingConvention = CallingConvention.Cdecl)] public static extern int lpr_tab_dims(IntPtr h, out int rows, out int cols, out int ld);
        [DllImport(Lib, CallingConvention = CallingConvention.Cdecl)] public static extern int lpr_tab_read(IntPtr h, [Out] double[,] host);
        [DllImport(Lib, CallingConvention = CallingConvention.Cdecl)] public static extern int lpr_tab_get_basis(IntPtr h, [Out] int[] basis);
        [DllImport(Lib, CallingConvention = CallingConvention.Cdecl)]
        public static extern int lpr_tab_solve(IntPtr h, int rule, long maxPivots, int flags, out int status, out long nPivots, [Out] int[] pivotLog, long logCap);
        [DllImport(Lib, CallingConvention = CallingConvention.Cdecl)] public static extern int lpr_tab_step(IntPtr h, int rule, out int enterCol, out int leaveRow, out int status);
        [DllImport(Lib, CallingConvention = CallingConvention.Cdecl)] public static extern int lpr_tab_extract_solution(IntPtr h, int n, [Out] double[] x);
        [DllImport(Lib, CallingConvention = CallingConvention.Cdecl)] public static extern int lpr_tab_objective(IntPtr h, out double z);
        [DllImport(Lib, CallingConvention = CallingConvention.Cdecl)] public static extern int lpr_tab_cutting_plane(IntPtr h, int maxCuts, out int status, out int nCuts, [Out] int[] cutLog, int cutLogCap);

        // ---- RunBranchAndBound entry (BranchBoundSimplexSolver.cs:28-113, :281-468) -------------------------
        [DllImport(Lib, CallingConvention = CallingConvention.Cdecl)]
        public static extern int lpr_tab_create_bb(int device, int n, int m, double[] objective, double[,] consPadded, int stride, int[] len, int rowCap, int colCap, out IntPtr h);
        [DllImport(Lib, CallingConvention = CallingConvention.Cdecl)]
        public static extern int lpr_tab_bb_node_solve_ex(IntPtr h, int isMinimization, long maxPivots, out int status, out long nPivots, [Out] int[] pivotLog, long logCap);

        // ---- SensitivityAnalyzer on a device-resident tableau (SensitivityAnalyzer.cs:609-723) ----------
        [DllImport(Lib, CallingConvention = CallingConvention.Cdecl)] public static extern int lpr_tab_sens_rebuild_basis(IntPtr h);
        [DllImport(Lib, CallingConvention = CallingConvention.Cdecl)] public static extern int lpr_tab_sens_solution(IntPtr h, [Out] double[] x);
        [DllImport(Lib, CallingConvention = CallingConvention.Cdecl)] public static extern int lpr_tab_sens_add_constraint(IntPtr h, double[] tech, double rhsMinusAx);

        // ---- revised simplex ---------------------------------------------------------------------
        [DllImport(Lib, CallingConvention = CallingConvention.Cdecl)]
        public static extern int lpr_rev_create(int device, int m, int n, double[,] A, double[] b, double[] c, int isMin, out IntPtr h);
        [DllImport(Lib, CallingConvention = CallingConvention.Cdecl)] public static extern int lpr_rev_destroy(IntPtr h);
        [DllImport(Lib, CallingConvention = CallingConvention.Cdecl)]
        public static extern int lpr_rev_solve(IntPtr h, long maxIter, int refactorEvery, out int status, out long nIter, [Out] int[] log, long logCap);
        [DllImport(Lib, CallingConvention = CallingConvention.Cdecl)] public static extern int lpr_rev_read_basis(IntPtr h, [Out] int[] basis);
        [DllImport(Lib, CallingConvention = CallingConvention.Cdecl)] public static extern int lpr_rev_read_x(IntPtr h, [Out] double[] x);
        [DllImport(Lib, CallingConvention = CallingConvention.Cdecl)] public static extern int lpr_rev_read_z(IntPtr h, out double z);
        [DllImport(Lib, CallingConvention = CallingConvention.Cdecl)] public static extern int lpr_rev_read_y(IntPtr h, [Out] double[] y);
        // one iteration at a time + the CaptureSnapshot text of that iteration (RevisedPrimalSimplexSolver.cs:82-250, :294-387)
        [DllImport(Lib, CallingConvention = CallingConvention.Cdecl)] public static extern int lpr_rev_begin(IntPtr h);
        [DllImport(Lib, CallingConvention = CallingConvention.Cdecl)] public static extern int lpr_rev_step(IntPtr h, out int status, out int enter, out int leaveRow, out int leaveVar);
        [DllImport(Lib, CallingConvention = CallingConvention.Cdecl)] public static extern int lpr_rev_format_snapshot(IntPtr h, out IntPtr utf8Text, out long len);
        [DllImport(Lib, CallingConvention = CallingConvention.Cdecl)] public static extern int lpr_rev_refactor(IntPtr h);

        // ---- integer programming -------------------------------------------------------------------
        [DllImport(Lib, CallingConvention = CallingConvention.Cdecl)]
        public static extern int lpr_bb_solve(int device, int rows, int cols, double[,] finalTableau, int nVars, int enablePruning, long maxNodes,
                                              [Out] double[] x, out double z, out int hasSolution, out long nodes, out long pivots,
                                              [Out] int[] nodeLog, [Out] double[] nodeZ, long nodeLogCap, out int status);
        [DllImport(Lib, CallingConvention = CallingConvention.Cdecl)]
        public static extern int lpr_knap_solve(int device, double capacity, int n, double[] weights, double[] values, long maxNodes,
                                                out double best, [Out] byte[] chosen, out long nodes, out int status);
        [DllImport(Lib, CallingConvention = CallingConvention.Cdecl)]
        public static extern int lpr_knap_dp(int device, int capacity, int n, int[] weights, int[] values, out double best, [Out] byte[] chosen);

        // ---- multi-GPU branch & bound inside the library (host threads + NCCL; stats may be IntPtr.Zero) ----------------
        [DllImport(Lib, CallingConvention = CallingConvention.Cdecl)]
        public static extern int lpr_bb_solve_mgpu(int nGpus, int[] devices, int rows, int cols, double[,] finalTableau, int nVars, int enablePruning,
                                                   long maxNodes, long maxRounds, double sliceSeconds, [Out] double[] x, out double z, out int hasSolution,
                                                   out long nodes, out long pivots, out int status, IntPtr stats);
        [DllImport(Lib, CallingConvention = CallingConvention.Cdecl)]
        public static extern int lpr_knap_solve_mgpu(int nGpus, int[] devices, double capacity, int n, double[] weights, double[] values, long maxNodes,
                                                     long maxRounds, double sliceSeconds, out double best, [Out] byte[] chosen, out long nodes, out int status, IntPtr stats);
        // ---- B&B building blocks on one device tableau (BranchBoundSimplexSolver.cs:552-567, :694-803) -----------------
        [DllImport(Lib, CallingConvention = CallingConvention.Cdecl)] public static extern int lpr_tab_round4(IntPtr h);
        [DllImport(Lib, CallingConvention = CallingConvention.Cdecl)] public static extern int lpr_tab_bb_add_constraint(IntPtr parent, int nVars, int var, double bound, int type, out IntPtr child);
        [DllImport(Lib, CallingConvention = CallingConvention.Cdecl)] public static extern int lpr_fmt_fixed(double x, int decimals, [Out] byte[] buf, int cap);

        public static string Utf8(IntPtr p, long len)
        {
            var bytes = new byte[len];
            Marshal.Copy(p, bytes, 0, (int)len);
            return System.Text.Encoding.UTF8.GetString(bytes);
        }

        public static void Check(int rc)
        {
            if (rc != OK) throw new InvalidOperationException("liblprb200: " + Marshal.PtrToStringAnsi(lpr_last_error()));
        }

        // ---- partitionable B&B node pool (one process per GPU; see INTEGRATION.md section 3) ---------------------
        [DllImport(Lib, CallingConvention = CallingConvention.Cdecl)] public static extern int lpr_bb_create(int device, int rows, int cols, double[,] rootTableau, int nVars, int enablePruning, out IntPtr pool);
        [DllImport(Lib, CallingConvention = CallingConvention.Cdecl)] public static extern int lpr_bb_destroy(IntPtr pool);
        [DllImport(Lib, CallingConvention = CallingConvention.Cdecl)] public static extern int lpr_bb_open_count(IntPtr pool, out long n);
        [DllImport(Lib, CallingConvention = CallingConvention.Cdecl)] public static extern int lpr_bb_run(IntPtr pool, long maxNodes, out long processed, out long pivots);
        [DllImport(Lib, CallingConvention = CallingConvention.Cdecl)] public static extern int lpr_bb_run_timed(IntPtr pool, long maxNodes, double maxSeconds, out long processed, out long pivots);
        [DllImport(Lib, CallingConvention = CallingConvention.Cdecl)] public static extern int lpr_bb_keep_stride(IntPtr pool, int offset, int stride);
        [DllImport(Lib, CallingConvention = CallingConvention.Cdecl)] public static extern int lpr_bb_get_incumbent(IntPtr pool, out int has, out double z, [Out] double[] x, [Out] int[] key, ref int keyLen);
        [DllImport(Lib, CallingConvention = CallingConvention.Cdecl)] public static extern int lpr_bb_set_incumbent(IntPtr pool, double z, double[] x, int[] key, int keyLen);
        [DllImport(Lib, CallingConvention = CallingConvention.Cdecl)] public static extern int lpr_bb_export_nodes(IntPtr pool, int maxNodes, IntPtr buf, long bufCap, out long bytes, out int nExported);
        [DllImport(Lib, CallingConvention = CallingConvention.Cdecl)] public static extern int lpr_bb_import_nodes(IntPtr pool, IntPtr buf, long bytes);

        // ---- model input (IO/InputFileParser.cs:19-68, Program.cs:114-124, :511-535) ----------------------
        [DllImport(Lib, CallingConvention = CallingConvention.Cdecl, CharSet = CharSet.Ansi)] public static extern int lpr_model_parse_file(string path, out IntPtr model);
        [DllImport(Lib, CallingConvention = CallingConvention.Cdecl)] public static extern int lpr_model_parse_text(byte[] utf8, long len, out IntPtr model);
        [DllImport(Lib, CallingConvention = CallingConvention.Cdecl)] public static extern int lpr_model_from_dense(int n, int m, double[] objective, double[,] coef, int[] relation, double[] rhs, int isMax, out IntPtr model);
        [DllImport(Lib, CallingConvention = CallingConvention.Cdecl)] public static extern int lpr_model_destroy(IntPtr model);
        [DllImport(Lib, CallingConvention = CallingConvention.Cdecl)] public static extern int lpr_model_info(IntPtr model, out int loaded, out int n, out int nConstraints, out int nSigns);
        [DllImport(Lib, CallingConvention = CallingConvention.Cdecl)] public static extern int lpr_model_problem_type(IntPtr model, [Out] byte[] buf, int cap);
        [DllImport(Lib, CallingConvention = CallingConvention.Cdecl)] public static extern int lpr_model_message(IntPtr model, [Out] byte[] buf, int cap);
        [DllImport(Lib, CallingConvention = CallingConvention.Cdecl)] public static extern int lpr_model_objective(IntPtr model, [Out] double[] c);
        [DllImport(Lib, CallingConvention = CallingConvention.Cdecl)] public static extern int lpr_model_constraint(IntPtr model, int i, [Out] double[] coef, int cap, out int count, [Out] byte[] relation, int relCap, out double rhs);
        [DllImport(Lib, CallingConvention = CallingConvention.Cdecl)] public static extern int lpr_model_sign(IntPtr model, int j, [Out] byte[] buf, int cap);
        [DllImport(Lib, CallingConvention = CallingConvention.Cdecl)] public static extern int lpr_model_add_cli_bound_rows(IntPtr model);
        [DllImport(Lib, CallingConvention = CallingConvention.Cdecl)] public static extern int lpr_model_add_upper_bound_rows(IntPtr model);
        [DllImport(Lib, CallingConvention = CallingConvention.Cdecl)] public static extern int lpr_tab_create_from_model(int device, IntPtr model, int isMax, out IntPtr h);
        [DllImport(Lib, CallingConvention = CallingConvention.Cdecl, CharSet = CharSet.Ansi)] public static extern int lpr_model_save_binary(IntPtr model, string path);
        [DllImport(Lib, CallingConvention = CallingConvention.Cdecl, CharSet = CharSet.Ansi)] public static extern int lpr_model_load_binary(string path, out IntPtr model);

        // ---- snapshots (Utilities/TableIterationFormater.cs:22-48, NumFormat.N3) ------------------------------
        [DllImport(Lib, CallingConvention = CallingConvention.Cdecl)] public static extern int lpr_fmt_f3(double x, [Out] byte[] buf, int cap);
        [DllImport(Lib, CallingConvention = CallingConvention.Cdecl)] public static extern int lpr_fmt_n3(double x, [Out] byte[] buf, int cap);
        [DllImport(Lib, CallingConvention = CallingConvention.Cdecl, CharSet = CharSet.Ansi)]
        public static extern int lpr_fmt_table(double[,] tab, int rows, int cols, long ld, int numOriginalVars, string title, string[] rowLabels, int nLabels, out IntPtr text, out long len);
        [DllImport(Lib, CallingConvention = CallingConvention.Cdecl)] public static extern int lpr_fmt_general(double x, [Out] byte[] buf, int cap);
        [DllImport(Lib, CallingConvention = CallingConvention.Cdecl)] public static extern int lpr_model_canonical_form(IntPtr model, out IntPtr text, out long len);
        [DllImport(Lib, CallingConvention = CallingConvention.Cdecl, CharSet = CharSet.Ansi)]
        public static extern int lpr_out_write_full_results(string path, string solverUsed, IntPtr model, string[] snapshots, int nSnapshots, double finalZ, double[] x, int nX, int append, string timestamp);
        [DllImport(Lib, CallingConvention = CallingConvention.Cdecl, CharSet = CharSet.Ansi)]
        public static extern int lpr_out_write_snapshots_only(string path, string solverUsed, string[] snapshots, int nSnapshots, double finalZ, double[] x, int nX, int append, string timestamp);
        [DllImport(Lib, CallingConvention = CallingConvention.Cdecl, CharSet = CharSet.Ansi)]
        public static extern int lpr_tab_format(IntPtr h, int numOriginalVars, string title, string[] rowLabels, int nLabels, out IntPtr text, out long len);
    }

    /// <summary>Owns a native handle; Dispose/finalizer frees the device memory.</summary>
    internal sealed class TabHandle : SafeHandle
    {
        public TabHandle(IntPtr h) : base(IntPtr.Zero, true) { SetHandle(h); }
        public override bool IsInvalid => handle == IntPtr.Zero;
        protected override bool ReleaseHandle() { return Lpr.lpr_tab_destroy(handle) == Lpr.OK; }
    }
    internal sealed class RevHandle : SafeHandle
    {
        public RevHandle(IntPtr h) : base(IntPtr.Zero, true) { SetHandle(h); }
        public override bool IsInvalid => handle == IntPtr.Zero;
        protected override bool ReleaseHandle() { return Lpr.lpr_rev_destroy(handle) == Lpr.OK; }
    }
    internal sealed class ModelHandle : SafeHandle
    {
        public ModelHandle(IntPtr h) : base(IntPtr.Zero, true) { SetHandle(h); }
        public override bool IsInvalid => handle == IntPtr.Zero;
        protected override bool ReleaseHandle() { return Lpr.lpr_model_destroy(handle) == Lpr.OK; }
    }
}
