// InputOutputShim.cs -- replacements for IO/InputFileParser.cs (ReadInputFile :19-68, Constraint :70-82) and
// Utilities/TableIterationFormater.cs (Format :19-48) over liblprb200 (SURVEY 8f rows 2-3).  Same public members
// and signatures as the reference's classes; not compiled here (no .NET toolchain), executed by tests/test_csharp_shims.py -- the same entry points are driven by
// lpr_381_group_v22_b200/io.py, utilities.py and host/lpr_solvers.hpp in the tests.
using System;
using System.Collections.Generic;
using System.Runtime.InteropServices;
using System.Text;
using LPR_381_Group_V22.Native;

namespace LPR_381_Group_V22.IO
{
    public class InputFileParser
    {
        public string ProblemType { get; set; }
        public List<double> ObjectiveCoefficients { get; set; } = new List<double>();
        public List<Constraint> Constraints { get; set; } = new List<Constraint>();
        public List<string> SignRestrictions { get; set; } = new List<string>();

        public class Constraint
        {
            public List<double> Coefficients { get; }
            public string Relation { get; }
            public double RHS { get; }
            public Constraint(List<double> coefficients, string relation, double rhs) { Coefficients = coefficients; Relation = relation; RHS = rhs; }
        }

        internal ModelHandle NativeModel; // lpr_tab_create_from_model builds the device tableau straight from it

        private static string Str(Func<byte[], int, int> call)
        {
            var buf = new byte[4096];
            if (call(buf, buf.Length) != Lpr.OK) throw new InvalidOperationException("liblprb200: " + Marshal.PtrToStringAnsi(Lpr.lpr_last_error()));
            return Encoding.UTF8.GetString(buf, 0, Array.IndexOf(buf, (byte)0));
        }

        public void ReadInputFile(string filePath)
        {
            if (Lpr.lpr_model_parse_file(filePath, out IntPtr m) != Lpr.OK)
            {
                // the native parser names the exception the reference would have thrown
                string msg = Marshal.PtrToStringAnsi(Lpr.lpr_last_error());
                if (msg.StartsWith("FormatException")) throw new FormatException(msg);
                if (msg.StartsWith("IndexOutOfRangeException")) throw new IndexOutOfRangeException(msg);
                throw new InvalidOperationException("liblprb200: " + msg);
            }
            var model = new ModelHandle(m);
            Lpr.lpr_model_info(m, out int loaded, out int n, out int rows, out int nSigns);
            Console.WriteLine(Str((b, c) => Lpr.lpr_model_message(m, b, c)));
            if (loaded == 0) { model.Dispose(); return; } // :21-34: message only, members untouched

            ProblemType = Str((b, c) => Lpr.lpr_model_problem_type(m, b, c));
            var obj = new double[n];
            Lpr.lpr_model_objective(m, obj);
            ObjectiveCoefficients.AddRange(obj);
            for (int i = 0; i < rows; i++)
            {
                Lpr.lpr_model_constraint(m, i, null, 0, out int cnt, null, 0, out double rhs);
                var coef = new double[cnt];
                var rel = new byte[64];
                Lpr.lpr_model_constraint(m, i, coef, cnt, out cnt, rel, rel.Length, out rhs);
                Constraints.Add(new Constraint(new List<double>(coef), Encoding.UTF8.GetString(rel, 0, Array.IndexOf(rel, (byte)0)), rhs));
            }
            for (int j = 0; j < nSigns; j++) SignRestrictions.Add(Str((b, c) => Lpr.lpr_model_sign(m, j, b, c)));
            NativeModel = model;
        }
    }
}

namespace LPR_381_Group_V22.Utilities
{
    public static class TableIterationFormater
    {
        public static string Format(double[,] tab, int numOriginalVars, string title) => Format(tab, numOriginalVars, title, null);

        public static string Format(double[,] tab, int numOriginalVars, string title, IReadOnlyList<string> rowLabels)
        {
            string[] labels = rowLabels == null ? null : new List<string>(rowLabels).ToArray();
            int rc = Lpr.lpr_fmt_table(tab, tab.GetLength(0), tab.GetLength(1), tab.GetLength(1), numOriginalVars, title,
                                       labels, labels?.Length ?? 0, out IntPtr text, out long len);
            if (rc != Lpr.OK) throw new InvalidOperationException("liblprb200: " + Marshal.PtrToStringAnsi(Lpr.lpr_last_error()));
            return Marshal.PtrToStringAnsi(text, (int)len);
        }

        // device-resident tableau: used by PrimalSimplexSolverShim for IterationSnapshots (no double[,] round trip)
        internal static string Format(TabHandle tableau, int numOriginalVars, string title, IReadOnlyList<string> rowLabels = null)
        {
            string[] labels = rowLabels == null ? null : new List<string>(rowLabels).ToArray();
            int rc = Lpr.lpr_tab_format(tableau.DangerousGetHandle(), numOriginalVars, title, labels, labels?.Length ?? 0,
                                        out IntPtr text, out long len);
            if (rc != Lpr.OK) throw new InvalidOperationException("liblprb200: " + Marshal.PtrToStringAnsi(Lpr.lpr_last_error()));
            return Marshal.PtrToStringAnsi(text, (int)len);
        }
    }
}
