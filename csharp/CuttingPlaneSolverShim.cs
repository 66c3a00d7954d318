// Replacement body for IntegerProgramming/CuttingPlaneSolver.cs:64-229.  The whole recursion (Gomory row generated on the
// device, cut pivot, dual then primal clean-up, next cut) is one native call; objectiveRow / constraintRows are mutated in
// place and the cut rows appended, exactly what the reference's in-place version leaves behind.  Not compiled here (no .NET); executed by tests/test_csharp_shims*.py.
using LPR_381_Group_V22.Native;
using LPR_381_Group_V22.Simplex;
using System;
using System.Collections.Generic;
using System.Linq;

namespace LPR_381_Group_V22.IntegerProgramming
{
    public class CuttingPlaneSolver
    {
        /// <summary>cuts performed by the last call as (chosen row, pivot column, dual pivots, primal pivots)</summary>
        public List<(int row, int col, int dualPivots, int primalPivots)> CutLog { get; } = new List<(int, int, int, int)>();
        public int Status { get; private set; }

        public void CuttingPlaneSolution(double[] objectiveRow, List<double[]> constraintRows)
        {
            if (objectiveRow == null) throw new ArgumentNullException(nameof(objectiveRow));
            if (constraintRows == null || constraintRows.Count == 0) throw new ArgumentException("No constraint rows.");
            int width = objectiveRow.Length;
            if (constraintRows.Any(r => r.Length != width))
                throw new ArgumentException("All rows (objective & constraints) must have the same length.");
            const int headroom = 64;   // the reference recurses without a bound; 64 cuts is far beyond anything it terminates on
            var t = RowTableau.Pack(objectiveRow, constraintRows);
            Lpr.Check(Lpr.lpr_tab_create(0, t.GetLength(0), width, t.GetLength(0) + headroom, 0, t, out IntPtr h));
            using (var tab = new TabHandle(h))
            {
                var log = new int[4 * headroom];
                Lpr.Check(Lpr.lpr_tab_cutting_plane(h, headroom - 1, out int status, out int nCuts, log, headroom));
                Status = status;
                CutLog.Clear();
                for (int k = 0; k < nCuts; k++) CutLog.Add((log[4 * k], log[4 * k + 1], log[4 * k + 2], log[4 * k + 3]));
                Lpr.Check(Lpr.lpr_tab_dims(h, out int rows, out int cols, out int _));
                var full = new double[rows, cols];
                Lpr.Check(Lpr.lpr_tab_read(h, full));
                for (int i = constraintRows.Count; i < rows - 1; i++) constraintRows.Add(new double[width]);   // appended cut rows (:110)
                for (int j = 0; j < width; j++) objectiveRow[j] = full[0, j];
                for (int i = 0; i < rows - 1; i++) for (int j = 0; j < width; j++) constraintRows[i][j] = full[i + 1, j];
                if (status == Lpr.PIVOT_TOO_SMALL) throw new InvalidOperationException("Pivot too small/zero.");
                Console.WriteLine($"Cutting plane: {nCuts} cut(s), final objective value {objectiveRow[width - 1]:0.###}");
            }
        }
    }
}
