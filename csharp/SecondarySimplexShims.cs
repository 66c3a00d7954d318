// Replacement bodies for Simplex/PrimalSimplexSolver2.cs (:24-229) and Simplex/DualSimplex.cs (:14-241): the pivot loops
// run on the GPU through lpr_tab_solve(rule = PRIMAL2 / DUAL); signatures, in-place semantics and exceptions are the
// reference's.  Not compiled here (no .NET toolchain); executed by tests/test_csharp_shims*.py (INTEGRATION.md).
using LPR_381_Group_V22.Native;
using LPR_381_Group_V22.Utilities;
using System;
using System.Collections.Generic;
using System.Linq;
using System.Text;

namespace LPR_381_Group_V22.Simplex
{
    internal static class RowTableau
    {
        /// <summary>(objective row, constraint rows) -> rectangular tableau; rows must all have the objective's width.</summary>
        public static double[,] Pack(double[] obj, List<double[]> rows)
        {
            var t = new double[rows.Count + 1, obj.Length];
            for (int j = 0; j < obj.Length; j++) t[0, j] = obj[j];
            for (int i = 0; i < rows.Count; i++) for (int j = 0; j < obj.Length; j++) t[i + 1, j] = rows[i][j];
            return t;
        }
        public static void Unpack(double[,] t, double[] obj, List<double[]> rows)
        {
            for (int j = 0; j < obj.Length; j++) obj[j] = t[0, j];
            for (int i = 0; i < rows.Count; i++) for (int j = 0; j < obj.Length; j++) rows[i][j] = t[i + 1, j];
        }
    }

    public class PrimalSimplexSolver2
    {
        private double[,] tableau;
        private readonly int rows, cols;
        private bool _isOptimal = false;
        public List<string> IterationSnapshots { get; } = new List<string>();
        public double FinalZ { get; private set; }
        public List<double> SolutionVector { get; private set; } = new List<double>();

        public PrimalSimplexSolver2(double[] objectiveRow, List<double[]> constraintRows)
        {
            if (objectiveRow == null) throw new ArgumentNullException(nameof(objectiveRow));
            if (constraintRows == null || constraintRows.Count == 0) throw new ArgumentException("No constraint rows.");
            if (constraintRows.Any(r => r.Length != objectiveRow.Length))
                throw new ArgumentException("All rows (obj & constraints) must have the same length.");
            rows = constraintRows.Count + 1;
            cols = objectiveRow.Length;
            tableau = RowTableau.Pack(objectiveRow, constraintRows);
        }

        /// <summary>text snapshots (three per pivot, like the reference) only for console-sized tableaux</summary>
        public static long TraceMaxElements = 4096;

        /// <summary>PrimalSimplexSolver2.Solve (:46-97).  Below TraceMaxElements the loop steps pivot by pivot and records the
        /// reference's snapshots ("Start of iter k", "Before pivot ...", "After pivot ...", :50-79) and console lines; above it
        /// the whole loop is one native call.  Either way the iteration counter only advances when printing (:70), so
        /// maxIters only bites with printSteps (flags bit0).</summary>
        public bool Solve(int maxIters = 10_000, bool printSteps = false)
        {
            Lpr.Check(Lpr.lpr_tab_create(0, rows, cols, 0, 0, tableau, out IntPtr h));
            using (var tab = new TabHandle(h))
            {
                int status;
                if ((long)rows * cols > TraceMaxElements)
                {
                    Lpr.Check(Lpr.lpr_tab_solve(h, Lpr.RULE_PRIMAL2, maxIters, printSteps ? 1 : 0, out status, out long _, null, 0));
                    Lpr.Check(Lpr.lpr_tab_read(h, tableau));
                }
                else
                {
                    int iter = 0;
                    while (true)
                    {
                        CaptureSnapshot($"Start of iter {iter}");                                        // :50
                        Lpr.Check(Lpr.lpr_tab_step(h, Lpr.RULE_PRIMAL2, out int pivotCol, out int pivotRow, out status));
                        if (status != Lpr.RUNNING)
                        {
                            if (printSteps) Console.WriteLine(status == Lpr.OPTIMAL ? "Optimal reached." : "Unbounded: no valid leaving row.");
                            break;
                        }
                        CaptureSnapshot($"Before pivot (iter {iter + 1}) at row {pivotRow}, col {pivotCol}");   // :68, tableau still pre-pivot
                        if (printSteps)
                        {
                            Console.WriteLine($"\nIter {++iter}: pivot @ row {pivotRow}, col {pivotCol}");
                            PrintTableau();
                        }
                        Lpr.Check(Lpr.lpr_tab_read(h, tableau));
                        CaptureSnapshot($"After pivot (iter {iter}) at row {pivotRow}, col {pivotCol}");        // :79
                        if (printSteps)
                        {
                            Console.WriteLine("After pivot:");
                            PrintTableau();
                        }
                        if (iter >= maxIters)                                                            // :87-92
                        {
                            if (printSteps) Console.WriteLine("Max iterations reached.");
                            status = Lpr.ITER_LIMIT;
                            break;
                        }
                    }
                }
                if (status == Lpr.PIVOT_TOO_SMALL) throw new InvalidOperationException("Pivot too small/zero.");   // :148-149
                _isOptimal = status == Lpr.OPTIMAL;
                if (_isOptimal) FinalZ = tableau[0, cols - 1];
                return _isOptimal;
            }
        }

        private void CaptureSnapshot(string title)   // :167-181 on the host copy of the tableau
        {
            var sb = new StringBuilder();
            if (!string.IsNullOrWhiteSpace(title)) sb.AppendLine(title);
            sb.AppendLine("Current Tableau:");
            sb.AppendLine("OBJ\t" + string.Join("\t", Enumerable.Range(0, cols).Select(j => tableau[0, j].ToString("0.###"))));
            for (int i = 1; i < rows; i++)
                sb.AppendLine($"r{i}\t" + string.Join("\t", Enumerable.Range(0, cols).Select(j => tableau[i, j].ToString("0.###"))));
            IterationSnapshots.Add(sb.ToString());
        }

        private void PrintTableau()   // :183-188
        {
            Console.WriteLine("OBJ: " + string.Join("\t", Enumerable.Range(0, cols).Select(j => tableau[0, j].ToString("0.####"))));
            for (int i = 1; i < rows; i++)
                Console.WriteLine($"r{i}: " + string.Join("\t", Enumerable.Range(0, cols).Select(j => tableau[i, j].ToString("0.####"))));
        }

        public double[] GetObjectiveRow(bool solveIfNeeded = true)
        {
            EnsureReady(solveIfNeeded);
            return Enumerable.Range(0, cols).Select(j => tableau[0, j]).ToArray();
        }
        public List<double[]> GetConstraintRows(bool solveIfNeeded = true)
        {
            EnsureReady(solveIfNeeded);
            return Enumerable.Range(1, rows - 1).Select(i => Enumerable.Range(0, cols).Select(j => tableau[i, j]).ToArray()).ToList();
        }
        public (double[] ObjectiveRow, List<double[]> ConstraintRows) GetRows(bool solveIfNeeded = true)
        {
            EnsureReady(solveIfNeeded);
            return (GetObjectiveRow(false), GetConstraintRows(false));
        }
        private void EnsureReady(bool solveIfNeeded)
        {
            if (!_isOptimal && solveIfNeeded && !Solve())
                throw new InvalidOperationException("Could not reach an optimal tableau (unbounded or infeasible).");   // :225
        }
    }
}

/// <summary>Simplex/DualSimplex.cs lives in the global namespace in the reference; so does this replacement.</summary>
public class DualSimplexSolver
{
    private const double EPS = 1e-9;

    /// <summary>Dual simplex (:14-114): mutates objectiveRow and constraintRows in place; true when every RHS ends >= 0.</summary>
    public bool Solve(double[] objectiveRow, List<double[]> constraintRows, int maxIters = 10_000, bool printSteps = true)
    {
        if (objectiveRow == null) throw new ArgumentNullException(nameof(objectiveRow));
        if (constraintRows == null || constraintRows.Count == 0) throw new ArgumentException("No constraint rows.");
        if (constraintRows.Any(r => r.Length != objectiveRow.Length))
            throw new ArgumentException("All rows (obj & constraints) must have the same length.");
        var t = LPR_381_Group_V22.Simplex.RowTableau.Pack(objectiveRow, constraintRows);
        Lpr.Check(Lpr.lpr_tab_create(0, t.GetLength(0), t.GetLength(1), 0, 0, t, out IntPtr h));
        using (var tab = new TabHandle(h))
        {
            // flags bit0 = printSteps: the reference's `iter` only advances when printing (:94, :108)
            Lpr.Check(Lpr.lpr_tab_solve(h, Lpr.RULE_DUAL, maxIters, printSteps ? 1 : 0, out int status, out long pivots, null, 0));
            Lpr.Check(Lpr.lpr_tab_read(h, t));
            LPR_381_Group_V22.Simplex.RowTableau.Unpack(t, objectiveRow, constraintRows);
            if (status == Lpr.PIVOT_TOO_SMALL) throw new InvalidOperationException("Pivot too small/zero.");   // :155-156
            if (printSteps) Console.WriteLine(status == Lpr.OPTIMAL ? $"Dual simplex finished after {pivots} pivots." : "Dual simplex stopped: infeasible or max iterations reached.");
            return status == Lpr.OPTIMAL;
        }
    }

    public double[] GetObjectiveRow(double[] objectiveRow, List<double[]> constraintRows, bool solveIfNeeded = true, int maxIters = 10_000, bool printSteps = false)
    {
        EnsureReady(objectiveRow, constraintRows, solveIfNeeded, maxIters, printSteps);
        return (double[])objectiveRow.Clone();
    }
    public List<double[]> GetConstraintRows(double[] objectiveRow, List<double[]> constraintRows, bool solveIfNeeded = true, int maxIters = 10_000, bool printSteps = false)
    {
        EnsureReady(objectiveRow, constraintRows, solveIfNeeded, maxIters, printSteps);
        return constraintRows.Select(r => (double[])r.Clone()).ToList();
    }
    public (double[] ObjectiveRow, List<double[]> ConstraintRows) GetRows(double[] objectiveRow, List<double[]> constraintRows, bool solveIfNeeded = true, int maxIters = 10_000, bool printSteps = false)
    {
        EnsureReady(objectiveRow, constraintRows, solveIfNeeded, maxIters, printSteps);
        return ((double[])objectiveRow.Clone(), constraintRows.Select(r => (double[])r.Clone()).ToList());
    }
    private void EnsureReady(double[] objectiveRow, List<double[]> constraintRows, bool solveIfNeeded, int maxIters, bool printSteps)
    {
        if (!solveIfNeeded || !AnyNegativeRhs(constraintRows)) return;
        if (!Solve(objectiveRow, constraintRows, maxIters, printSteps))
            throw new InvalidOperationException("Dual phase could not reach feasibility (infeasible or max iterations).");   // :146
    }
    public static bool AnyNegativeRhs(List<double[]> rows)   // :180-185
    {
        if (rows == null || rows.Count == 0) return false;
        int rhs = rows[0].Length - 1;
        return rows.Any(r => r[rhs] < -EPS);
    }
    public static void PrintTableau(double[] objectiveRow, List<double[]> constraintRows, int numVars, string title = null)   // :193-207
    {
        var table = LPR_381_Group_V22.Simplex.RowTableau.Pack(objectiveRow, constraintRows ?? new List<double[]>());
        Console.WriteLine(TableIterationFormater.Format(table, numVars, title ?? "Dual Simplex Tableau"));
    }
}
