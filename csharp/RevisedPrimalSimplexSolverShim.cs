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

        public void Solve()
        {
            Lpr.Check(Lpr.lpr_rev_solve(rev.DangerousGetHandle(), -1, 0, out int status, out long _, null, 0));
            if (status == Lpr.INFEASIBLE) throw new Exception("Infeasible basis (negative basic value).");
            if (status == Lpr.UNBOUNDED) throw new Exception("Unbounded problem (no positive component in direction).");
            if (status == Lpr.PIVOT_TOO_SMALL) throw new Exception("Pivot too small.");
            var x = new double[numVariables];
            Lpr.Check(Lpr.lpr_rev_read_x(rev.DangerousGetHandle(), x));
            Lpr.Check(Lpr.lpr_rev_read_z(rev.DangerousGetHandle(), out double z));
            SolutionVector = x.ToList();
            FinalZ = z;
            IterationSnapshots.Add("Optimal\nDual prices (y = c_B^T B^{-1}):\n" + string.Join("\t", DualPrices.Select(NumFormat.N3)));
        }
        public void Dispose() { rev.Dispose(); }
    }
}
