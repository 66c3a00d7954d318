// Replacement bodies for IntegerProgramming/BranchBoundSimplexSolver.cs: DualSimplexSolverBB (:12-469), TeeTextWriter
// (:471-487) and BranchAndBound (:490-1312).  List<List<double>> tableaux are marshalled as rectangular arrays; every
// pivot, AddConstraint and the tree search itself run in liblprb200.  Not compiled here (no .NET toolchain in the build image); executed by tests/test_csharp_shims*.py.
using LPR_381_Group_V22.Native;
using System;
using System.Collections.Generic;
using System.IO;
using System.Linq;
using System.Text;

namespace LPR_381_Group_V22.IntegerProgramming
{
    public class BranchBoundSimplexSolver
    {
        internal static double[,] ToArray(List<List<double>> t)
        {
            var a = new double[t.Count, t[0].Count];
            for (int i = 0; i < t.Count; i++) for (int j = 0; j < t[0].Count; j++) a[i, j] = t[i][j];
            return a;
        }
        internal static List<List<double>> ToLists(double[,] a)
        {
            var t = new List<List<double>>(a.GetLength(0));
            for (int i = 0; i < a.GetLength(0); i++) t.Add(Enumerable.Range(0, a.GetLength(1)).Select(j => a[i, j]).ToList());
            return t;
        }
        internal static List<List<double>> ReadTableau(IntPtr h)
        {
            Lpr.Check(Lpr.lpr_tab_dims(h, out int r, out int c, out int _));
            var a = new double[r, c];
            Lpr.Check(Lpr.lpr_tab_read(h, a));
            return ToLists(a);
        }

        public class DualSimplexSolverBB
        {
            public DualSimplexSolverBB() { }

            /// <summary>:28-113.  Rows are [coefficients..., rhs, type flag]; like the reference the caller's rows are mutated
            /// (">=" rows negated, flags removed) and the tableau is (m+1) x (n+m+1), built on the device.</summary>
            public List<List<double>> FormulateTableau(List<double> objectiveFunction, List<List<double>> constraints)
            {
                int n = objectiveFunction.Count, m = constraints.Count, stride = Math.Max(2, constraints.Max(r => r.Count));
                var cons = new double[m, stride];
                var len = new int[m];
                for (int i = 0; i < m; i++) { len[i] = constraints[i].Count; for (int j = 0; j < len[i]; j++) cons[i, j] = constraints[i][j]; }
                Lpr.Check(Lpr.lpr_tab_create_bb(0, n, m, objectiveFunction.ToArray(), cons, stride, len, 0, 0, out IntPtr h));
                using (var tab = new TabHandle(h))
                {
                    foreach (var row in constraints)   // what :42-56 does to the caller's lists
                    {
                        if (row[row.Count - 1] == 1) for (int j = 0; j < row.Count; j++) row[j] = -1 * row[j];
                        row.RemoveAt(row.Count - 1);
                    }
                    return ReadTableau(h);
                }
            }

            private static (List<List<double>>, List<double>) OnePivot(List<List<double>> tableau, bool dual, bool isMinimization)
            {
                var a = ToArray(tableau);
                int cols = a.GetLength(1);
                // theta values are display-only in the reference (:139-150, :222-241); recomputed here from the input tableau
                var thetas = new List<double>();
                Lpr.Check(Lpr.lpr_tab_create(0, a.GetLength(0), cols, 0, 0, a, out IntPtr h));
                using (var tab = new TabHandle(h))
                {
                    var log = new int[2];
                    // one pivot of the DoDualSimplex state machine: a dual pivot while some RHS is negative (which is when the
                    // reference calls PerformDualPivot, :315-343), otherwise a primal one (:345-390).  As standalone calls the
                    // two members therefore equal the reference's only in those states, a primal pivot that leaves a negative
                    // RHS is dropped like DoDualSimplex drops it (:392-400), and -0.0 entries come back as 0.0 (:307-313);
                    // nothing outside DoDualSimplex calls them upstream.
                    Lpr.Check(Lpr.lpr_tab_bb_node_solve_ex(h, isMinimization ? 1 : 0, 1, out int _, out long done, log, 1));
                    if (done == 0) return (tableau, null);   // no pivot possible: the reference's catch -> (tableau, null) (:152-172)
                    int r = log[0], c = log[1];
                    if (dual) for (int j = 0; j < cols - 1; j++) thetas.Add(a[r, j] < 0 ? Math.Abs(a[0, j] / a[r, j]) : double.PositiveInfinity);
                    else for (int i = 1; i < a.GetLength(0); i++) thetas.Add(a[i, c] != 0 ? a[i, cols - 1] / a[i, c] : double.PositiveInfinity);
                    return (ReadTableau(h), thetas);
                }
            }
            public (List<List<double>> updatedTableau, List<double> thetaValues) PerformDualPivot(List<List<double>> tableau) => OnePivot(tableau, true, false);
            public (List<List<double>> updatedTableau, List<double> thetaValues) PerformPrimalPivot(List<List<double>> tableau, bool isMinimization) => OnePivot(tableau, false, isMinimization);

            public (List<List<double>> tableau, bool isMinimization, int surplusCount, int slackCount, int objectiveLength) PrepareInput(List<double> objectiveFunction, List<List<double>> constraints, bool isMinimization)
            {
                int surplus = constraints.Count(c => c[c.Count - 1] == 1 || c[c.Count - 1] == 2);
                int slack = constraints.Count - surplus;
                return (FormulateTableau(objectiveFunction, constraints), isMinimization, surplus, slack, objectiveFunction.Count);
            }

            /// <summary>:289-468: dual phase while a RHS is negative, then primal phase; optimalValue == null means infeasible.
            /// The list of tableaux holds the start and the final tableau (the intermediate ones are display-only upstream).</summary>
            public (List<List<List<double>>> tableaux, List<double> decisionVariables, double? optimalValue, List<int> pivotCols, List<int> pivotRows, List<string> headerRow)
                DoDualSimplex(List<double> objectiveFunction, List<List<double>> constraints, bool isMinimization, List<List<double>> tableauOverride = null)
            {
                var start = tableauOverride ?? FormulateTableau(objectiveFunction, constraints);
                var a = ToArray(start);
                Lpr.Check(Lpr.lpr_tab_create(0, a.GetLength(0), a.GetLength(1), 0, 0, a, out IntPtr h));
                using (var tab = new TabHandle(h))
                {
                    var log = new int[2 * 4096];
                    Lpr.Check(Lpr.lpr_tab_bb_node_solve_ex(h, isMinimization ? 1 : 0, -1, out int status, out long nPivots, log, 4096));
                    var final = ReadTableau(h);
                    var header = Enumerable.Range(0, a.GetLength(1) - 1).Select(j => j < objectiveFunction.Count ? $"x{j + 1}" : $"s{j - objectiveFunction.Count + 1}").Concat(new[] { "rhs" }).ToList();
                    if (status != Lpr.OPTIMAL) return (new List<List<List<double>>> { start }, null, null, null, null, header);
                    int k = (int)Math.Min(nPivots, 4096);
                    return (new List<List<List<double>>> { start, final }, null, final[0][final[0].Count - 1],
                            Enumerable.Range(0, k).Select(q => log[2 * q + 1]).ToList(), Enumerable.Range(0, k).Select(q => log[2 * q]).ToList(), header);
                }
            }
        }

        public sealed class TeeTextWriter : TextWriter
        {
            private readonly TextWriter _console;
            private readonly StringWriter _buffer;
            public TeeTextWriter(TextWriter console, StringWriter buffer) { _console = console; _buffer = buffer; }
            public override Encoding Encoding => _console.Encoding;
            public override void Write(char value) { _console.Write(value); _buffer.Write(value); }
            public override void Write(string value) { _console.Write(value); _buffer.Write(value); }
            public override void WriteLine(string value) { _console.WriteLine(value); _buffer.WriteLine(value); }
        }

        public class BranchAndBound
        {
            private List<double> objectiveCoefficients = new List<double> { 0.0, 0.0 };
            private List<List<double>> constraintMatrix = new List<List<double>>();
            private List<List<List<double>>> simplexTableaux;
            /// <summary>node budget: 20 = the reference's cap (:1038-1042); < 0 lifts it</summary>
            public long MaxNodes { get; set; } = 20;
            /// <summary>GPUs the node pool is partitioned over when the cap is lifted (lpr_bb_solve_mgpu)</summary>
            public int Gpus { get; set; } = 1;

            public BranchAndBound() { }
            public void SetNumVars(int n) { objectiveCoefficients = Enumerable.Repeat(0.0, n).ToList(); }   // :497-500
            public double RoundNumber(double number) => Math.Round(number, 4);                              // :540-550
            public List<double> RoundVector(List<double> v) => v.Select(RoundNumber).ToList();
            public bool IsInteger(double value) { double r = RoundNumber(value); return Math.Abs(r - Math.Round(r)) <= 1e-6; }   // :595-599

            public List<List<double>> RoundTableau(List<List<double>> tableau)   // :552-567, on the device
            {
                var a = ToArray(tableau);
                Lpr.Check(Lpr.lpr_tab_create(0, a.GetLength(0), a.GetLength(1), 0, 0, a, out IntPtr h));
                using (var tab = new TabHandle(h)) { Lpr.Check(Lpr.lpr_tab_round4(h)); return ReadTableau(h); }
            }
            public List<List<List<double>>> RoundAllTableaux(List<List<List<double>>> tableaux) => tableaux.Select(RoundTableau).ToList();

            /// <summary>:694-803 for the one shape the reference ever passes (:1104, :1171): a single bound row
            /// [e_k..., bound, type] (type 0 = "<=", 1 = ">=").</summary>
            public (List<List<double>> outputTab, List<List<double>> updatedTab) AddConstraint(List<List<double>> newConstraints, List<List<double>> baseTableau = null)
            {
                if (baseTableau == null) { Console.WriteLine("Input tableau required"); return (null, null); }
                if (newConstraints.Count != 1) throw new NotSupportedException("one bound row per call, as ExecuteBranchAndBound does");
                var row = newConstraints[0];
                int var = row.IndexOf(1.0); double bound = row[row.Count - 2]; int type = (int)row[row.Count - 1];
                var a = ToArray(baseTableau);
                Lpr.Check(Lpr.lpr_tab_create(0, a.GetLength(0), a.GetLength(1), 0, 0, a, out IntPtr h));
                using (var parent = new TabHandle(h))
                {
                    Lpr.Check(Lpr.lpr_tab_bb_add_constraint(h, objectiveCoefficients.Count, var < 0 ? 0 : var, bound, type, out IntPtr c));
                    using (var child = new TabHandle(c)) { var t = ReadTableau(c); return (t, t); }
                }
            }

            /// <summary>:1006-1233 -- the depth-first search in the reference's order (20-node cap unless MaxNodes says otherwise).</summary>
            public (List<double> optimalSolution, double optimalValue) ExecuteBranchAndBound(List<List<List<double>>> initialTableaux, bool enablePruning = false)
            {
                var root = ToArray(initialTableaux[initialTableaux.Count - 1]);
                int n = objectiveCoefficients.Count;
                var x = new double[n];
                double z; int has;
                if (Gpus > 1)
                    Lpr.Check(Lpr.lpr_bb_solve_mgpu(Gpus, null, root.GetLength(0), root.GetLength(1), root, n, enablePruning ? 1 : 0, MaxNodes, -1, 0.0,
                                                    x, out z, out has, out long _, out long _, out int _, IntPtr.Zero));
                else
                    Lpr.Check(Lpr.lpr_bb_solve(0, root.GetLength(0), root.GetLength(1), root, n, enablePruning ? 1 : 0, MaxNodes, x, out z, out has,
                                               out long _, out long _, null, null, 0, out int _));
                Console.WriteLine(has != 0 ? $"Best integer solution: z = {z}" : "No integer solution found.");
                return (has != 0 ? x.ToList() : null, has != 0 ? z : double.NegativeInfinity);
            }

            public (List<double> objective, List<List<double>> constraints) ConfigureProblem(List<double> objective, List<List<double>> constraints)   // :1233-1251
            {
                for (int i = 0; i < objective.Count; i++)
                {
                    var row = Enumerable.Repeat(0.0, objective.Count + 3).ToList();   // one entry longer than a model row: upstream quirk
                    row[i] = 1.0; row[objective.Count + 1] = 1.0;
                    constraints.Add(row);
                }
                return (objective, constraints);
            }

            public void RunBranchAndBound(List<double> objectivePassed, List<List<double>> constraintsPassed, bool isMin)   // :1253-1298
            {
                objectiveCoefficients = objectivePassed.ToList();
                constraintMatrix = constraintsPassed.Select(r => r.ToList()).ToList();
                (objectiveCoefficients, constraintMatrix) = ConfigureProblem(objectiveCoefficients, constraintMatrix);
                var lp = new DualSimplexSolverBB().DoDualSimplex(objectiveCoefficients.ToList(), constraintMatrix.Select(r => r.ToList()).ToList(), isMin);
                if (lp.optimalValue == null) throw new InvalidOperationException("initial LP relaxation failed");
                simplexTableaux = RoundAllTableaux(new List<List<List<double>>> { lp.tableaux[lp.tableaux.Count - 1] });
                ExecuteBranchAndBound(simplexTableaux, false);
            }
        }
    }
}
