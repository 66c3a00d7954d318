// LprNative.cs -- P/Invoke binding of liblprb200 (include/lprb200.h).  Drop into the reference project
// (LPR_381_Group_V22/Native/); cannot be compiled in the build image (no .NET toolchain) -- see INTEGRATION.md.
using System;
using System.Runtime.InteropServices;

namespace LPR_381_Group_V22.Native
{
    internal static class Lpr
    {
        private const string Lib = "lprb200"; // liblprb200.so / lprb200.dll next to the executable

        public const int OK = 0;
        public const int RUNNING = 0, OPTIMAL = 1, UNBOUNDED = 2, INFEASIBLE = 3, ITER_LIMIT = 4, NODE_LIMIT = 5,
                         PIVOT_TOO_SMALL = 6, NO_CUT_NEEDED = 7, NO_PIVOT_COL = 8, CUT_STEP_DONE = 9;
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

        public static void Check(int rc)
        {
            if (rc != OK) throw new InvalidOperationException("liblprb200: " + Marshal.PtrToStringAnsi(lpr_last_error()));
        }
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
}
