// Replacement bodies for the re-optimisation members of SensitivityAnalysis/SensitivityAnalyzer.cs
// (ctor :22-41, Pivot/ReOptimize/DualSimplexIfNeeded/ResolveAll :98-209, AddNewConstraintNonInteractive :609-659,
// RebuildBasicsFromTableau :706-723): the final tableau lives in HBM between menu actions; the interactive
// menus, ranging and printing members stay as they are and read CurrentTableau.
using LPR_381_Group_V22.Native;
using System;
using System.Collections.Generic;

namespace LPR_381_Group_V22.SensitivityAnalysis
{
    public partial class SensitivityAnalyzer : IDisposable
    {
        private const int RULE_SENS = 3, OPTIMAL = 1, UNBOUNDED = 2, INFEASIBLE = 3, ITER_LIMIT = 4, HEADROOM = 16;
        private IntPtr h;
        private List<double> solutionVector;
        private double finalZ;

        public SensitivityAnalyzer(double[,] finalTableau, List<double> solution, double zValue, List<int> basicVariables)
        {
            var t = (double[,])finalTableau.Clone();
            int rows = t.GetLength(0), cols = t.GetLength(1);
            t[0, cols - 1] = zValue;                                           // :33
            solutionVector = new List<double>(solution);
            finalZ = zValue;
            Lpr.Check(Lpr.lpr_tab_create(0, rows, cols, rows + HEADROOM, cols + HEADROOM, t, out h));
            Lpr.Check(Lpr.lpr_tab_sens_rebuild_basis(h));                       // :36
            ValidateBinaryConstraints();                                        // :39
        }

        private void ValidateBinaryConstraints()                               // :43-51
        {
            for (int i = 0; i < Math.Min(solutionVector.Count, 6); i++)
                if (Math.Abs(solutionVector[i] - Math.Round(solutionVector[i])) > 1e-9)
                    Console.WriteLine($"Warning: x{i + 1} = {solutionVector[i]:0.###} violates binary constraint.");
        }

        public double[,] CurrentTableau
        {
            get { Lpr.Check(Lpr.lpr_tab_dims(h, out int r, out int c, out int _)); var t = new double[r, c]; Lpr.Check(Lpr.lpr_tab_read(h, t)); return t; }
        }
        public double CurrentZ => finalZ;

        private void ResolveAll()                                               // :203-209
        {
            Lpr.Check(Lpr.lpr_tab_sens_rebuild_basis(h));
            Lpr.Check(Lpr.lpr_tab_solve(h, RULE_SENS, 10000, 0, out int status, out long _, null, 0));
            if (status == INFEASIBLE) throw new InvalidOperationException("Infeasible after RHS change (dual simplex).");
            if (status == UNBOUNDED) throw new InvalidOperationException("Unbounded during re-optimization.");
            if (status == ITER_LIMIT) throw new InvalidOperationException("Re-optimization exceeded iteration limit.");
            Lpr.Check(Lpr.lpr_tab_objective(h, out finalZ));
            Lpr.Check(Lpr.lpr_tab_dims(h, out int _, out int cols, out int _));
            var x = new double[cols - 1];
            Lpr.Check(Lpr.lpr_tab_sens_solution(h, x));
            solutionVector = new List<double>(x);
            ValidateBinaryConstraints();                                        // :165
        }

        public void AddNewConstraintNonInteractive(double[] tech, double rhs)  // :609-659
        {
            double aX = 0.0;
            for (int j = 0; j < Math.Min(tech.Length, solutionVector.Count); j++) aX += tech[j] * solutionVector[j];
            Lpr.Check(Lpr.lpr_tab_sens_add_constraint(h, tech, rhs - aX));
            ResolveAll();
        }

        public void Dispose() { if (h != IntPtr.Zero) { Lpr.lpr_tab_destroy(h); h = IntPtr.Zero; } }
    }
}
