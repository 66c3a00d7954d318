// Replacement body for Simplex/PrimalSimplexSolver.cs: same public members (IterationSnapshots, FinalZ,
// SolutionVector, FinalTableau, FinalLabels, FinalTable, Solve(), GetFinalTableau(), BasicVariables), pivots on the GPU.
using LPR_381_Group_V22.Native;
using LPR_381_Group_V22.Utilities;
using System;
using System.Collections.Generic;
using System.Linq;
using System.Text;
using IOConstraint = LPR_381_Group_V22.IO.InputFileParser.Constraint;

namespace LPR_381_Group_V22.Simplex
{
    public class PrimalSimplexSolver : IDisposable
    {
        private const int TraceMaxElements = 4096;   // text snapshots only for console-sized models
        private readonly TabHandle tab;
        private readonly int numVariables, numConstraints, rows, cols;
        private readonly bool trace;

        public List<string> IterationSnapshots = new List<string>();
        public double FinalZ { get; private set; }
        public List<double> SolutionVector { get; private set; }
        public double[,] FinalTableau { get; private set; }
        public IReadOnlyList<string> FinalLabels => BasicVariables.Select(ColLabel).ToList().AsReadOnly();
        public string FinalTable => FinalTableau == null ? "" : TableIterationFormater.Format(FinalTableau, numVariables, "Final Table", FinalLabels);

        public PrimalSimplexSolver(List<double> objective, List<IOConstraint> constraints, bool isMaximization = true)
        {
            numVariables = objective.Count;
            numConstraints = constraints.Count;
            int stride = Math.Max(numVariables, constraints.Max(c => c.Coefficients.Count));
            var coef = new double[numConstraints * stride];
            var count = new int[numConstraints];
            var rel = new int[numConstraints];
            var rhs = new double[numConstraints];
            for (int i = 0; i < numConstraints; i++)
            {
                var c = constraints[i];
                count[i] = c.Coefficients.Count;
                for (int j = 0; j < count[i]; j++) coef[i * stride + j] = c.Coefficients[j];
                rel[i] = c.Relation == ">=" ? 1 : (c.Relation == "=" ? 2 : 0);
                rhs[i] = c.RHS;
            }
            Lpr.Check(Lpr.lpr_tab_create_primal(0, numVariables, numConstraints, objective.ToArray(), coef, stride, count, rel, rhs,
                                                isMaximization ? 1 : 0, out IntPtr h));
            tab = new TabHandle(h);
            rows = numConstraints + 1;
            cols = numVariables + numConstraints + 1;
            trace = (long)rows * cols <= TraceMaxElements;
            if (trace) IterationSnapshots.Add(TableIterationFormater.Format(Read(), numVariables, "Initial Tableau"));
        }

        public void Solve()
        {
            int status;
            if (trace)
            {
                int iteration = 0;
                var before = Read();
                while (true)
                {
                    Lpr.Check(Lpr.lpr_tab_step(tab.DangerousGetHandle(), Lpr.RULE_PRIMAL, out int e, out int r, out status));
                    if (status != Lpr.RUNNING) break;
                    // the console lines of :137-146, in the reference's order (the pivot position is known after the step)
                    Console.WriteLine($"\nIteration {++iteration}: pivot @ constraint {r}, column {ColLabel(e)}");
                    Console.WriteLine(TableIterationFormater.Format(before, numVariables, "Before pivot"));
                    var t = Read();
                    Console.WriteLine($"After pivot (constraint {r}, column {ColLabel(e)}):");
                    Console.WriteLine(TableIterationFormater.Format(t, numVariables, "After pivot"));
                    IterationSnapshots.Add(TableIterationFormater.Format(t, numVariables, $"Iteration {iteration} - After pivot"));
                    before = t;
                }
            }
            else
            {
                Lpr.Check(Lpr.lpr_tab_solve(tab.DangerousGetHandle(), Lpr.RULE_PRIMAL, -1, 0, out status, out long _, null, 0));
            }
            FinalTableau = Read();
            if (status == Lpr.OPTIMAL)
            {
                Lpr.Check(Lpr.lpr_tab_objective(tab.DangerousGetHandle(), out double z));
                FinalZ = z;
                var x = new double[numVariables];
                Lpr.Check(Lpr.lpr_tab_extract_solution(tab.DangerousGetHandle(), numVariables, x));
                SolutionVector = x.ToList();
                Console.WriteLine("Optimal Solution Found!");
                if (trace)
                {
                    // :118-122: the last snapshot is the final tableau followed by the solution summary
                    var finalBlock = new StringBuilder();
                    finalBlock.AppendLine(TableIterationFormater.Format(FinalTableau, numVariables, "Final Tableau (Optimal)"));
                    finalBlock.AppendLine(SolutionSummary());
                    IterationSnapshots.Add(finalBlock.ToString());
                }
                Console.WriteLine(SolutionSummary());
                Console.WriteLine(new string('-', 100));
            }
            else if (status == Lpr.UNBOUNDED)
            {
                Console.WriteLine("Unbounded Solution!");   // FinalZ stays 0, SolutionVector stays null
                if (trace) IterationSnapshots.Add(TableIterationFormater.Format(FinalTableau, numVariables, "Unbounded Tableau"));
            }
        }

        public double[,] GetFinalTableau() => Read();

        public List<int> BasicVariables
        {
            get { var b = new int[numConstraints]; Lpr.Check(Lpr.lpr_tab_get_basis(tab.DangerousGetHandle(), b)); return b.ToList(); }
        }

        private double[,] Read() { var t = new double[rows, cols]; Lpr.Check(Lpr.lpr_tab_read(tab.DangerousGetHandle(), t)); return t; }
        private string ColLabel(int col) => col < numVariables ? $"x{col + 1}" : $"t{col - numVariables + 1}";

        private string SolutionSummary(string title = "Optimal solution")   // :256-267
        {
            var sb = new StringBuilder();
            sb.AppendLine(title + ":");
            sb.AppendLine($"Z = {FinalZ:F6}");
            if (SolutionVector != null)
                for (int i = 0; i < numVariables; i++) sb.AppendLine($"x{i + 1} = {SolutionVector[i]:F6}");
            return sb.ToString();
        }
        public void Dispose() { tab.Dispose(); }
    }
}
