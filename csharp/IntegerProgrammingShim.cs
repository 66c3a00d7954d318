// Replacement bodies for IntegerProgramming/BranchAndBoundAdapter.cs and the missing knapsack classes.
using LPR_381_Group_V22.Native;
using LPR_381_Group_V22.Simplex;
using System;
using System.Collections.Generic;
using System.Linq;

namespace LPR_381_Group_V22.IntegerProgramming
{
    public static class BranchAndBoundAdapter
    {
        public static (List<double> x, double z) SolveFromPrimal(PrimalSimplexSolver primal, bool enablePruning = false, bool isMin = false)
        {
            if (primal.FinalTableau == null) throw new InvalidOperationException("Primal simplex has not been solved yet.");
            var t = primal.FinalTableau;
            int n = primal.SolutionVector?.Count ?? Math.Max(1, t.GetLength(1) - 1);
            var x = new double[n];
            Lpr.Check(Lpr.lpr_bb_solve(0, t.GetLength(0), t.GetLength(1), t, n, enablePruning ? 1 : 0, 20 /* reference cap */, x,
                                       out double z, out int has, out long _, out long _, null, null, 0, out int _));
            return (has != 0 ? x.ToList() : new List<double>(), has != 0 ? z : double.NegativeInfinity);
        }
    }

    public sealed class KnapsackItem { public int Id; public double Value; public double Weight; }

    public class KnapsackBranchBoundSimplex
    {
        private readonly double capacity; private readonly double[] weights, values; private byte[] chosen; private long nodes; private double best;
        public KnapsackBranchBoundSimplex(int capacity, double[] weights, double[] values) { this.capacity = capacity; this.weights = weights; this.values = values; }
        /// <summary>GPUs the open-node pool is partitioned over (inside the library: host threads + NCCL)</summary>
        public int Gpus { get; set; } = 1;
        public double Solve()
        {
            chosen = new byte[weights.Length];
            if (Gpus > 1)
                Lpr.Check(Lpr.lpr_knap_solve_mgpu(Gpus, null, capacity, weights.Length, weights, values, -1, -1, 0.0, out best, chosen, out nodes, out int _, IntPtr.Zero));
            else
                Lpr.Check(Lpr.lpr_knap_solve(0, capacity, weights.Length, weights, values, -1, out best, chosen, out nodes, out int _));
            return best;
        }
        public void PrintIterations() { Console.WriteLine($"Knapsack B&B: {nodes} nodes processed, best value {best}"); }
        public List<KnapsackItem> GetSelectedItemsOriginal() =>
            Enumerable.Range(0, weights.Length).Where(i => chosen[i] != 0).Select(i => new KnapsackItem { Id = i, Value = values[i], Weight = weights[i] }).ToList();
    }

    internal class KnapsackBranchBoundSolver
    {
        public static double Solve(int capacity, int[] weights, int[] values)
        {
            Lpr.Check(Lpr.lpr_knap_dp(0, capacity, weights.Length, weights, values, out double best, new byte[weights.Length]));
            return best;
        }
    }
}
