// Replacement body for Simplex/RevisedPrimalSimplexSolver.cs (NumFormat stays as in the reference file).
using LPR_381_Group_V22.Native;
using System;
using System.Collections.Generic;
using System.Linq;
using static LPR_381_Group_V22.IO.InputFileParser;

namespace LPR_381_Group_V22.Simplex
{
    public class RevisedPrimalSimplexSolver : IDisposable
    {
        private readonly RevHandle rev;
        private readonly int numVariables, numConstraints;
        public List<string> IterationSnapshots { get; private set; } = new List<string>();
        public double FinalZ { get; private set; }
        public List<double> SolutionVector { get; private set; } = new List<double>();
        public List<int> BasicVariables
        {
            get { var b = new int[numConstraints]; Lpr.Check(Lpr.lpr_rev_read_basis(rev.DangerousGetHandle(), b)); return b.ToList(); }
        }
        /// <summary>y = c_B B^-1; only printed by the reference, exported for numerical checks.</summary>
        public double[] DualPrices
        {
            get { var y = new double[numConstraints]; Lpr.Check(Lpr.lpr_rev_read_y(rev.DangerousGetHandle(), y)); return y; }
        }

        public RevisedPrimalSimplexSolver(List<double> objective, List<Constraint> constraints, bool isMinimization)
        {
            if (objective == null || objective.Count == 0) throw new ArgumentException("Objective cannot be null or empty.");
            if (constraints == null || constraints.Count == 0) throw new ArgumentException("Constraints cannot be null or empty.");
            numVariables = objective.Count;
            numConstraints = constraints.Count;
            var A = new double[numConstraints, numVariables];
            var b = new double[numConstraints];
            for (int i = 0; i < numConstraints; i++)
            {
                if (constraints[i].Coefficients.Count != numVariables)
                    throw new ArgumentException($"Constraint {i + 1} has incorrect number of coefficients.");
                for (int j = 0; j < numVariables; j++) A[i, j] = constraints[i].Coefficients[j];
                b[i] = constraints[i].RHS;   // Relation is ignored, as in the reference
            }
            Lpr.Check(Lpr.lpr_rev_create(0, numConstraints, numVariables, A, b, objective.ToArray(), isMinimization ? 1 : 0, out IntPtr h));
            rev = new RevHandle(h);
        }

        /// <summary>Tracing threshold (SURVEY 8b "Snapshots"): below it Solve() steps iteration by iteration and records the
        /// reference's CaptureSnapshot text for each one; above it the whole loop runs on the device and no text is built
        /// (one snapshot of BASELINE cfg3 would be gigabytes).</summary>
        public static long TraceMaxElements = 4096;

        public void Solve()
        {
            IntPtr h = rev.DangerousGetHandle();
            int status;
            if ((long)numConstraints * (numVariables + numConstraints + 1) <= TraceMaxElements)
            {
                Lpr.Check(Lpr.lpr_rev_begin(h));
                while (true)
                {
                    Lpr.Check(Lpr.lpr_rev_step(h, out status, out int _, out int _, out int _));
                    if (status != Lpr.RUNNING && status != Lpr.OPTIMAL) break;
                    Lpr.Check(Lpr.lpr_rev_format_snapshot(h, out IntPtr text, out long len));   // "Iteration k" / "Optimal" (:226-246, :124-146)
                    IterationSnapshots.Add(Lpr.Utf8(text, len));
                    if (status == Lpr.OPTIMAL) break;
                }
            }
            else
            {
                Lpr.Check(Lpr.lpr_rev_solve(h, -1, 0, out status, out long _, null, 0));
            }
            if (status == Lpr.INFEASIBLE) throw new Exception("Infeasible basis (negative basic value).");
            if (status == Lpr.UNBOUNDED) throw new Exception("Unbounded problem (no positive component in direction).");
            if (status == Lpr.PIVOT_TOO_SMALL) throw new Exception("Pivot too small.");
            var x = new double[numVariables];
            Lpr.Check(Lpr.lpr_rev_read_x(h, x));
            Lpr.Check(Lpr.lpr_rev_read_z(h, out double z));
            SolutionVector = x.ToList();
            FinalZ = z;
        }
        public void Dispose() { rev.Dispose(); }
    }
}
