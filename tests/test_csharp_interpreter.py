"""Conformance tests of the C# interpreter under oracle/csharp/ (test infrastructure): each case is a small C# program
whose result follows from the language specification (ECMA-334) or from the .NET Framework reference source, derived by
hand.  The interpreter is what pins the oracle to the reference (tests/test_reference_run.py), so the rules it relies on
are spelled out here one by one.  No file of the reference is needed.
"""
import math
import os
import sys

import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))

from csharp import CsException, Interpreter  # noqa: E402
from csharp.csparse import CsSyntaxError, parse_source  # noqa: E402
from csharp.csrun import format_double  # noqa: E402

HEAD = "using System; using System.Collections.Generic; using System.Linq; using System.Text;\n"


def run(body, method="F", *args, cls="T"):
    it = Interpreter()
    it.load_source(HEAD + "public static class T {\n" + body + "\n}")
    return it.call_static(cls, method, *args)


def test_operator_precedence_and_associativity():
    assert run("public static int F() { return 2 + 3 * 4 - 10 / 4 % 3; }") == 12          # 2 + 12 - ((10/4)%3 = 2)
    assert run("public static bool F() { return 1 + 1 == 2 && 3 > 2 || false && 1 / 0 == 0; }") is True
    assert run("public static int F() { int a = 1, b = 2, c = 3; a = b = c; return a + b; }") == 6
    assert run("public static int F() { int x = 5; return x > 3 ? x > 4 ? 1 : 2 : 3; }") == 1
    assert run("public static int F() { return 7 & 3 | 8 ^ 1; }") == (7 & 3 | 8 ^ 1)
    assert run("public static int F() { return -2 * -3 - -1; }") == 7
    assert run("public static int F() { int i = 5; return i++ + ++i; }") == 12           # 5 + 7
    assert run("public static int F() { return 1 << 3 >> 1; }") == 4
    assert run("public static bool F() { int? n = null; return (n ?? 3) == 3 && !(n > 0) && !(n <= 0); }") is True


def test_integer_and_floating_point_arithmetic():
    assert run("public static int F() { return -7 / 2; }") == -3 and run("public static int F() { return -7 % 2; }") == -1
    assert run("public static double F() { return -7 / 2; }") == -3.0                   # integer division, then converted
    assert run("public static double F() { return -7 / 2.0; }") == -3.5
    assert run("public static double F() { double z = 0; return 1 / z; }") == math.inf
    assert math.isnan(run("public static double F() { double z = 0; return z / z; }"))
    assert run("public static double F() { double z = 0; return -1 / z; }") == -math.inf
    assert math.copysign(1.0, run("public static double F() { return -1 * 0.0; }")) == -1.0
    assert run("public static bool F() { return -0.0 == 0.0; }") is True
    assert run("public static double F() { return 5.5 % 2; }") == 1.5 and run("public static double F() { return -5.5 % 2; }") == -1.5
    assert run("public static double F() { return 0.1 + 0.2; }") == 0.30000000000000004  # no extended precision, no FMA
    assert run("public static double F() { double a = 1e16; return (a + 1) - a; }") == 0.0
    assert run("public static int F() { return (int)3.99 + (int)-3.99; }") == 0
    assert run("public static double F() { int n = 3; double s = 0; for (int i = 1; i <= n; i++) s += 1 / i; return s; }") == 1.0
    assert run("public static double F() { int n = 3; double s = 0; for (int i = 1; i <= n; i++) s += 1.0 / i; return s; }") == 1.0 + 0.5 + 1.0 / 3
    with pytest.raises(CsException) as e:
        run("public static int F() { int z = 0; return 5 % z; }")
    assert e.value.tname == "DivideByZeroException"


def test_implicit_numeric_conversion_sites():
    src = """
    private static double half(double v) { return v / 2; }
    private static double asDouble(int v) { return v; }
    public static double Field = 3;
    public static double Prop { get; set; } = 7;
    public static double F() {
        double a = 1; double[] arr = new double[2]; arr[0] = 1; var l = new List<double> { 1 }; l.Add(1); l[0] = 3; l.Insert(0, 5);
        double[,] m = new double[1, 1]; m[0, 0] = 1;
        var t = Pair();
        Prop = 9;
        return a / 2 + arr[0] / 2 + l[1] / 2 + l[0] / 2 + m[0, 0] / 2 + half(1) + asDouble(1) / 2 + Field / 2 + Prop / 2 + t.x / 2;
    }
    private static (double x, int y) Pair() { return (1, 2); }"""
    assert run(src) == 0.5 + 0.5 + 1.5 + 2.5 + 0.5 + 0.5 + 0.5 + 1.5 + 4.5 + 0.5


def test_math_and_rounding_rules():
    src = """public static string F() {
        return Math.Round(2.5) + "," + Math.Round(3.5) + "," + Math.Round(-2.5) + "," + Math.Round(2.50000001) + "|"
             + Math.Round(1.00005, 4) + "," + Math.Round(1.00015, 4) + "," + Math.Round(2.675, 2) + "|"
             + Math.Round(1.0005, 3, MidpointRounding.AwayFromZero) + "," + Math.Round(-2.5, MidpointRounding.AwayFromZero) + "|"
             + Math.Floor(-1.5) + "," + Math.Ceiling(-1.5) + "," + Math.Abs(-3) + "," + Math.Max(2, 3.5) + "," + Math.Min(-0.0, 0.0) + "|"
             + (int)Math.Ceiling(2.000001) + "," + Math.Sqrt(16) + "," + Math.Pow(2, 10) + "," + Math.Sign(-2.5) + "," + Math.Truncate(-2.7);
    }"""
    # the CLR scales, rounds to even, unscales -- in binary64: 1.00005 * 1e4 = 10000.500000000002 (up), 2.675 * 100 = 267.5
    # exactly (tie -> 268), where exact decimal rounding of the stored values would give 1.0001 and 2.67
    assert run(src) == "2,4,-2,3|1.0001,1.0002,2.68|1.001,-3|-2,-1,3,3.5,0|3,4,1024,-1,-2"
    assert 1.00005 * 1e4 == 10000.500000000002 and 2.675 * 100 == 267.5
    with pytest.raises(CsException) as e:
        run("public static double F() { return Math.Round(1.5, 16); }")
    assert e.value.tname == "ArgumentOutOfRangeException"


def test_number_to_string():
    cases = {1.0: "1", -1.5: "-1.5", 0.1 + 0.2: "0.3", 1 / 3: "0.333333333333333", 2 / 3: "0.666666666666667",
             1e15: "1E+15", 123456789012345.0: "123456789012345", 1e-5: "1E-05", 0.0001: "0.0001", 1.5e-7: "1.5E-07",
             -0.0: "0", 1e300: "1E+300", 15.399999999999999: "15.4", math.inf: "Infinity", -math.inf: "-Infinity"}
    for v, s in cases.items():
        assert format_double(v) == s, v
    assert format_double(math.nan) == "NaN"
    for v, fmt, s in ((2.5, "F0", "3"), (3.5, "F0", "4"), (0.125, "F2", "0.13"), (-0.0004, "F3", "0.000"), (1234.5678, "N2", "1,234.57"),
                      (15.4, "F6", "15.400000"), (0.5, "0.###", "0.5"), (12.34567, "0.###", "12.346"), (-0.0001, "0.###", "0"),
                      (3.0, "0.####", "3"), (0.5, "0.00", "0.50"), (1234.5, "#,##0.0", "1,234.5"), (0.256, "P1", "25.6 %"),
                      (12345.678, "E2", "1.23E+004"), (0.1, "R", "0.1"), (0.30000000000000004, "R", "0.30000000000000004")):
        assert format_double(v, fmt) == s, (v, fmt)
    src = """public static string F() {
        double v = 3.14159; int n = 42; string s = "ab"; double? q = null;
        return $"[{v:F2}|{v,8:F1}|{v,-8:F1}|{n:D4}|{n,5}|{s,4}|{q}|{{literal}}|{(n > 1 ? "y" : "n")}]" + string.Format("{0}-{1:F1}-{0,3}", n, v);
    }"""
    assert run(src) == "[3.14|     3.1|3.1     |0042|   42|  ab||{literal}|y]42-3.1- 42"
    assert run('public static string F() { return "a" + 1 + 2.5 + true + null + \'c\' + 1.0; }') == "a12.5Truec1"


def test_collections_follow_the_framework():
    src = """public static string F() {
        var l = new List<int> { 3, 1, 2 };
        var sb = new StringBuilder();
        l.Sort(); sb.Append(string.Join(",", l)).Append('|');
        l.Insert(1, 9); l.RemoveAt(0); l.Remove(2); sb.Append(string.Join(",", l)).Append('|');
        sb.Append(l.IndexOf(7)).Append(l.Contains(9)).Append(l.Count).Append('|');
        var d = new Dictionary<string, int>(); d["a"] = 1; d["a"]++; d["b"] = 5;
        sb.Append(d["a"]).Append(d.ContainsKey("c")).Append(d.Count).Append('|');
        try { var z = d["zz"]; } catch (KeyNotFoundException) { sb.Append("K"); }
        try { d.Add("a", 1); } catch (ArgumentException) { sb.Append("A"); }
        var st = new Stack<(int a, string b)>(); st.Push((1, "x")); st.Push((2, "y"));
        var (p, q) = st.Pop(); sb.Append(p).Append(q).Append(st.Count).Append(st.Peek().b).Append('|');
        var hs = new HashSet<int>(new[] { 1, 2, 2, 3 }); sb.Append(hs.Count).Append(hs.Contains(2)).Append(hs.Add(3)).Append('|');
        try { new Stack<int>().Pop(); } catch (InvalidOperationException) { sb.Append("E"); }
        double[,] m = { { 1, 2, 3 }, { 4, 5, 6 } };
        sb.Append(m.GetLength(0)).Append(m.GetLength(1)).Append(m.Length).Append(m[1, 2]).Append('|');
        var c = (double[,])m.Clone(); c[0, 0] = 9; sb.Append(m[0, 0]);
        double[] src = { 1, 2, 3 }, dst = new double[3]; Array.Copy(src, dst, 2); sb.Append(dst[1]).Append(dst[2]);
        return sb.ToString();
    }"""
    assert run(src) == "1,2,3|9,3|-1True2|2False2|KA2y1x|3TrueFalse|E2366|120"


def test_linq_semantics():
    src = """public static string F() {
        var xs = new List<double> { 4, -1, 2.5, -1 };
        var sb = new StringBuilder();
        sb.Append(xs.Where(v => v > 0).Select(v => v * 2).Sum()).Append('|');
        sb.Append(xs.Min()).Append(xs.Max()).Append(xs.IndexOf(xs.Min())).Append('|');          // first of equal minima
        sb.Append(xs.Any(v => v > 4)).Append(xs.All(v => v > -2)).Append(xs.Count(v => v < 0)).Append('|');
        sb.Append(string.Join(",", xs.OrderBy(v => v))).Append('|');
        sb.Append(string.Join(",", Enumerable.Range(1, 4).Select(i => i * i).Reverse())).Append('|');
        sb.Append(xs.Take(2).Last()).Append(xs.Skip(3).First()).Append(xs.DefaultIfEmpty(7).Count()).Append('|');
        sb.Append(new List<double>().DefaultIfEmpty(double.PositiveInfinity).Min()).Append('|');
        sb.Append(string.Join(",", xs.Zip(new[] { 1, 2, 3 }, (a, b) => a * b))).Append('|');
        sb.Append(string.Join(",", new[] { "b", "a", "c" }.Select((s, i) => s + i)));
        var stable = new[] { (1, "x"), (0, "y"), (1, "z"), (0, "w") }.OrderBy(t => t.Item1).Select(t => t.Item2);
        sb.Append('|').Append(string.Join("", stable));
        return sb.ToString();
    }"""
    assert run(src) == "13|-141|FalseTrue2|-1,-1,2.5,4|16,9,4,1|-1-14|Infinity|4,-2,7.5|b0,a1,c2|ywxz"
    assert math.isnan(run("public static double F() { return new List<double> { 1, double.NaN, 0 }.Min(); }"))
    assert run("public static double F() { return new List<double> { double.NaN, 1, 3 }.Max(); }") == 3.0


def test_control_flow_exceptions_and_closures():
    src = """public static string F() {
        var sb = new StringBuilder();
        for (int i = 0; i < 6; i++) { if (i == 1) continue; if (i == 4) break; sb.Append(i); }
        int k = 0; do { k += 2; } while (k < 5); sb.Append('|').Append(k);
        switch (k) { case 4: sb.Append("four"); break; case 6: case 7: sb.Append("six"); break; default: sb.Append("?"); break; }
        try { try { throw new InvalidOperationException("boom"); } finally { sb.Append("|fin"); } }
        catch (ArgumentException) { sb.Append("wrong"); }
        catch (Exception e) when (e.Message == "boom") { sb.Append("|" + e.Message + e.GetType().Name.Length); }
        try { try { throw new ArgumentNullException("p"); } catch (ArgumentException) { sb.Append("|base"); throw; } }
        catch (Exception e) { sb.Append(e.Message.Contains("p")); }
        var fs = new List<Func<int>>();
        foreach (var v in new[] { 1, 2, 3 }) fs.Add(() => v * 10);       // foreach: a fresh variable per iteration
        sb.Append('|').Append(string.Join(",", fs.Select(f => f())));
        int total = 0; Action<int> add = n => total += n; add(2); add(3); sb.Append('|').Append(total);
        int outv; bool ok = int.TryParse("12", out outv); double dv; bool no = double.TryParse("x", out dv);
        sb.Append('|').Append(ok).Append(outv).Append(no).Append(dv);
        return sb.ToString();
    }"""
    got = run(src)
    assert got.startswith("023|6six|fin|boom")
    assert "|baseTrue|" in got and got.endswith("|5|True12False0")
    assert "|10,20,30|" in got


def test_classes_properties_tuples_and_nullables():
    it = Interpreter()
    it.load_source(HEAD + """
    namespace N {
      public class Acc {
        private double total; private readonly List<double> log = new List<double>();
        public int Count { get; private set; }
        public double Mean => Count == 0 ? 0 : total / Count;
        public double Last { get { return log.Count > 0 ? log[log.Count - 1] : double.NaN; } }
        public static int Instances = 0;
        public const double Eps = 1e-9;
        public Acc(double start = 0) { total = start; Instances++; }
        public Acc Add(double v, bool record = true) { total += v; Count++; if (record) log.Add(v); return this; }
        public (double sum, int n) State() => (total, Count);
        public double? Find(double v) { foreach (var x in log) if (Math.Abs(x - v) < Eps) return x; return null; }
        public class Inner { public int Twice(int v) => 2 * v + Instances; }
      }
      public static class T {
        public static string F() {
          var a = new Acc(start: 1).Add(2).Add(3, record: false);
          var (s, n) = a.State();
          var st = a.State();
          double? hit = a.Find(2), miss = a.Find(5);
          var inner = new Acc.Inner();
          return $"{s}|{n}|{st.sum + st.n}|{a.Mean}|{a.Last}|{hit.HasValue}{hit.Value}|{miss.HasValue}{miss ?? -1}|{inner.Twice(4)}|{Acc.Instances}";
        }
      }
    }""")
    assert it.call_static("T", "F") == "6|2|8|3|2|True2|False-1|9|1"
    with pytest.raises(CsException) as e:
        it2 = Interpreter()
        it2.load_source("public static class T { public static double F() { double? q = null; return q.Value; } }")
        it2.call_static("T", "F")
    assert e.value.tname == "InvalidOperationException"


def test_strings():
    src = """public static string F() {
        string s = "  +1 2  3 ";
        var parts = s.Trim().Split(' ');
        var clean = s.Split(new[] { ' ' }, StringSplitOptions.RemoveEmptyEntries);
        return parts.Length + "," + clean.Length + "," + parts[2].Length + "|" + "AbC".ToLower() + "abc".ToUpper() + "|"
             + "hello".Substring(1, 3) + "hello".IndexOf("l") + "hello".Replace("l", "L") + "|" + new string('-', 3)
             + string.IsNullOrWhiteSpace("  ") + string.IsNullOrEmpty(null) + "|" + double.Parse("1,234.5") + double.Parse(" -1e1 ")
             + "|" + "a,b".Contains(",") + "x".PadLeft(3) + "x".PadRight(3) + "|" + "A".Equals("a", StringComparison.OrdinalIgnoreCase);
    }"""
    assert run(src) == "4,3,0|abcABC|ell2heLLo|---TrueTrue|1234.5-10|True  xx  |True"
    with pytest.raises(CsException) as e:
        run('public static double F() { return double.Parse(""); }')
    assert e.value.tname == "FormatException"


def test_the_parser_rejects_what_it_does_not_understand():
    for bad in ("class A { void F() { int x = ; } }", "class A { void F() { foo(; } }", "class A { int F() => ; }",
                "class A { operator +(A a, A b) { } }"):
        with pytest.raises(CsSyntaxError):
            parse_source(bad)
    unit = parse_source("namespace A.B { public sealed class C<T> : D, I { [Attr(1)] public static extern int f(out int x, [In] double[,] m); } }")
    cls = unit[2][0]
    assert cls[1] == "C" and cls[6] == "A.B" and cls[4][0][0] == "method" and "extern" in cls[4][0][1] and cls[4][0][5] is None


def test_three_statements_of_the_number_formatting_rules_agree():
    """csrc/host_io.cu (native), tests/net_reference.py and oracle/csharp/csrun.py were written separately; on random and
    edge doubles "F3", "F6", NumFormat.N3's building blocks and double.ToString() must come out the same from all three"""
    import ctypes as C
    import numpy as np
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import net_reference as R
    from lpr_381_group_v22_b200 import _native as N
    from lpr_381_group_v22_b200.utilities import F3, NumFormat

    def general(x):
        buf = C.create_string_buffer(64)
        N.check(N.lib().lpr_fmt_general(float(x), buf, 64))
        return buf.value.decode()
    rng = np.random.default_rng(21)
    xs = list((rng.normal(size=1500) * 10.0 ** rng.integers(-9, 17, 1500)).tolist())
    xs += [round(float(v), 4) + 0.0005 for v in rng.uniform(-50, 50, 300)]        # ties of the third decimal
    xs += [0.0, -0.0, 0.0005, -0.0005, 1e-13, 2.5, 0.125, 1e15, 1e16, 123456789.123456789, -7.9995, 1 / 3, 5e-5, 99999.9995]
    for x in xs:
        assert F3(x) == R.F3(x) == format_double(x, "F3"), x
        assert NumFormat.Fixed(x, 6) == R.net_fixed(x, 6) == format_double(x, "F6"), x
        assert general(x) == R.net_general(x) == format_double(x), x
        assert NumFormat.N3(x) == R.N3(x), x
