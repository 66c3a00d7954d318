"""CPU tests of the native model-input and snapshot-formatting entry points (csrc/host_io.cu):
the parser against the reference's fixtures and its documented failure modes, the .NET number formatting against an
independent Python restatement (tests/net_reference.py) and hand-derived strings."""
import os

import numpy as np
import pytest

import lpr_381_group_v22_b200 as L
from lpr_381_group_v22_b200.io import Model
from lpr_381_group_v22_b200.utilities import F3, NumFormat, TableIterationFormater

import net_reference as R

HERE = os.path.dirname(os.path.abspath(__file__))


# ---- parser -------------------------------------------------------------------------------------------------
def test_parse_reference_fixture_text():
    # data/TextFile.txt of the reference, verbatim (no trailing newline, '+' signs)
    m = Model.parse_text("max +2 +3 +3 +5 +2 +4\n+11 +8 +6 +14 +10 +10 <= 40\nbin bin bin bin bin bin")
    assert m.info() == (True, 6, 1, 6) and m.problem_type == "max"
    assert m.objective() == [2, 3, 3, 5, 2, 4]
    c = m.constraints()[0]
    assert c.Coefficients == [11, 8, 6, 14, 10, 10] and c.Relation == "<=" and c.RHS == 40
    assert m.signs() == ["bin"] * 6
    assert m.message == "Your file was read and is in the correct format!"


def test_parse_line_endings_bom_and_blanks():
    text = "﻿MAX 1 2.5\r\n  1   1  >=  3  ignored tokens\r\n1,000 -1e1 = 7\r\n + urs \r\n"
    m = Model.parse_text(text)
    assert m.info() == (True, 2, 2, 2) and m.problem_type == "max"  # ToLower() :36
    cons = m.constraints()
    assert [c.Relation for c in cons] == [">=", "="] and cons[0].Coefficients == [1, 1] and cons[0].RHS == 3
    assert cons[1].Coefficients == [1000.0, -10.0] and cons[1].RHS == 7  # AllowThousands, exponent
    assert m.signs() == ["+", "urs"]
    assert Model.parse_text("min 1\r1 <= 2\r+").info() == (True, 1, 1, 1)  # lone CR line ends


def test_parse_failure_modes_match_the_reference():
    short = Model.parse_text("max 1 2\n+ +")
    assert short.info() == (False, 0, 0, 0) and short.message == "The input file is not formatted correctly."
    missing = Model.parse_file("/nonexistent/dir/model.txt")
    assert missing.info()[0] is False and "can't find your file" in missing.message
    with pytest.raises(ValueError, match="FormatException"):  # "max 1  2": Split(' ') keeps the empty token
        Model.parse_text("max 1  2\n1 1 <= 2\n+ +")
    with pytest.raises(ValueError, match="FormatException"):
        Model.parse_text("max 1 x\n1 1 <= 2\n+ +")
    with pytest.raises(IndexError, match="IndexOutOfRangeException"):
        Model.parse_text("max 1 2\n1\n+ +")  # the line runs out of tokens
    with pytest.raises(IndexError, match="IndexOutOfRangeException"):
        Model.parse_text("max 1 2\n1 1 <=\n+ +")  # no right-hand side
    with pytest.raises(ValueError, match="FormatException"):
        Model.parse_text("max 1 2\n1 <= 2\n+ +")  # "<=" is read as the second coefficient first (:50-53)
    with pytest.raises(ValueError):
        Model.parse_text("max 1 2\n1 1 <= 0x10\n+ +")  # strtod extensions are not .NET syntax
    with pytest.raises(ValueError):
        Model.parse_text("max 1 2\n1 1 <= 1e999\n+ +")  # OverflowException on the Framework
    for bad in (",5", "1e", ".", "+", "1-", "(1)", "$1", "\u00a05", "0x1p3", "inf", "nan", "1_0"):
        with pytest.raises((ValueError, IndexError)):
            Model.parse_text(f"max 1 2\n1 1 <= {bad}\n+ +")
    for ok, val in (("1,,0", 10.0), (".5", 0.5), ("5.", 5.0), ("-1E+2", -100.0), ("\t7\t", 7.0), ("-.25e1", -2.5), ("NaN", None)):
        got = Model.parse_text(f"max 1 2\n1 1 <= {ok}\n+ +").constraints()[0].RHS
        assert (got != got) if val is None else got == val, ok
    with pytest.raises(L.LprError):
        Model.parse_text("max 1\n+").to_device()  # not loaded


def test_input_file_parser_class_and_model_rows(tmp_path, capsys):
    p = L.InputFileParser()
    p.ReadInputFile(os.path.join(HERE, "golden", "model_a.txt"))
    assert "correct format" in capsys.readouterr().out
    assert p.ProblemType == "max" and p.SignRestrictions == ["bin"] * 6
    n = len(p.ObjectiveCoefficients)
    # native CLI rows == list helpers
    ref = L.add_cli_bound_rows(n, list(p.Constraints))
    p.Model.add_cli_bound_rows()
    got = p.Model.constraints()
    assert [(c.Coefficients, c.Relation, c.RHS) for c in got] == [(c.Coefficients, c.Relation, c.RHS) for c in ref]
    q = Model.parse_text("max 1 2 3\n1 1 1 <= 2\nbin + 0<=x<=1")
    q.add_upper_bound_rows()
    ref = L.add_upper_bound_constraints(3, ["bin", "+", "0<=x<=1"], [])
    assert [c.Coefficients for c in q.constraints()[1:]] == [c.Coefficients for c in ref] == [[1, 0, 0], [0, 0, 1]]
    r = Model.parse_text("max 1 2 3\n1 1 1 <= 2\nx≤1")  # one restriction applies to every variable (:519)
    r.add_upper_bound_rows()
    assert len(r.constraints()) == 4


def test_binary_model_round_trip(tmp_path):
    rng = np.random.default_rng(3)
    A, b, c = rng.normal(size=(7, 5)), rng.normal(size=7), rng.normal(size=5)
    m = Model.from_dense(c, A, b, relation=["<=", ">=", "=", "<=", "<=", ">=", "<="], is_maximization=False)
    path = str(tmp_path / "model.lprm")
    m.save_binary(path)
    k = Model.load_binary(path)
    assert k.info() == m.info() == (True, 5, 7, 0) and k.problem_type == "min"
    assert k.objective() == c.tolist()
    for a, bb, i in zip(k.constraints(), m.constraints(), range(7)):
        assert a.Coefficients == bb.Coefficients == A[i].tolist() and a.Relation == bb.Relation and a.RHS == b[i]
    open(path, "r+b").truncate(40)
    with pytest.raises(L.LprError):
        Model.load_binary(path)
    # a parsed model with ragged CLI rows and sign strings survives too
    p = Model.parse_text("max 1 2\n1 1 <= 2\nbin bin").add_cli_bound_rows()
    p.save_binary(path)
    q = Model.load_binary(path)
    assert [c.Coefficients for c in q.constraints()] == [c.Coefficients for c in p.constraints()]
    assert q.signs() == ["bin", "bin"] and len(q.constraints()[1].Coefficients) == 5


# ---- number formatting ----------------------------------------------------------------------------------------
def test_f3_known_strings():
    known = {1.0: "1.000", 0.0005: "0.001", -0.0001: "0.000", -2.5: "-2.500", 15.399999999999999: "15.400",
             2.0005: "2.001", 0.0: "0.000", -0.0: "0.000", 999.9995: "1000.000", -0.0005: "-0.001", 1e15: "1000000000000000.000",
             123456789012345678.0: "123456789012346000.000", 1e-300: "0.000", 0.9995: "1.000", 0.99949999: "0.999",
             float("nan"): "NaN", float("inf"): "Infinity", float("-inf"): "-Infinity", 2.675: "2.675", 1.0005: "1.001"}
    for x, s in known.items():
        assert F3(x) == s, (x, F3(x), s)


def test_f3_against_independent_restatement():
    rng = np.random.default_rng(11)
    xs = np.concatenate([rng.normal(size=4000) * 10.0 ** rng.integers(-6, 18, 4000), rng.integers(-10**6, 10**6, 3000) / 2000.0,
                         (rng.integers(-10**7, 10**7, 3000) + 0.5) / 1000.0, [5e-324, 1.7976931348623157e308, -4.9e-4, 4.9999999999e-4]])
    for x in xs.tolist():
        assert F3(x) == R.F3(x), x


def test_fixed_f6_of_the_optimal_summary():
    """{v:F6} of PrimalSimplexSolver.cs:256-267 through lpr_fmt_fixed (Python's :.6f prints "-0.000000" for
    a negative that rounds to zero and rounds half to even on the binary value)."""
    from lpr_381_group_v22_b200.utilities import NumFormat
    known = {-1e-17: "0.000000", -0.0: "0.000000", 0.0000005: "0.000001", -0.0000005: "-0.000001", 15.4: "15.400000",
             2.0000005: "2.000001", 1.0000015: "1.000002", 0.1234565: "0.123457", 9.0: "9.000000", -4.9e-7: "0.000000"}
    for x, t in known.items():
        assert NumFormat.Fixed(x, 6) == t, (x, NumFormat.Fixed(x, 6), t)
    rng = np.random.default_rng(12)
    xs = np.concatenate([rng.normal(size=2000) * 10.0 ** rng.integers(-8, 12, 2000), (rng.integers(-10**8, 10**8, 2000) + 0.5) / 1e6])
    for x in xs.tolist():
        assert NumFormat.Fixed(x, 6) == R.net_fixed(x, 6), x
        assert NumFormat.Fixed(x, 3) == R.F3(x)


def test_n3_literal_framework_rounding():
    known = {1e-13: "0", 2.0: "2", 0.125: "0.125", 0.2000000000000001: "0.2", -1.9999999: "-2", 0.0005: "0.001",
             -0.0: "0", 1234.5678: "1234.568", -0.25: "-0.25", 1e15: "1E+15", 123456789012345678.0: "1.23456789012346E+17",
             999999999999999.0: "999999999999999", 1.0005: "1.001", 4.0005: "4", 2.5: "2.5", -7.0: "-7", 0.0004999: "0"}
    for x, s in known.items():
        assert NumFormat.N3(x) == s, (x, NumFormat.N3(x), s)
    # 4.0005 * 1e3 = 4000.4999999999995 in binary64: the Framework's scale-and-split rounds DOWN (a decimal
    # half-up on the shortest string would give 4.001), while 1.0005 * 1e3 is exactly 1000.5 and goes up
    assert 4.0005 * 1e3 < 4000.5 and 1.0005 * 1e3 == 1000.5
    rng = np.random.default_rng(12)
    xs = np.concatenate([rng.normal(size=3000) * 10.0 ** rng.integers(-5, 8, 3000), rng.integers(-10**6, 10**6, 2000) / 2000.0])
    for x in xs.tolist():
        assert NumFormat.N3(x) == R.N3(x), x


def test_table_format_small_and_labels():
    tab = [[0.0, -1.0, 2.0], [1.0, 0.5, 3.0], [-0.00049, 7.25, 1e6]]
    txt = TableIterationFormater.Format(tab, 1, "Initial Tableau")
    assert txt == R.format_table(tab, 1, "Initial Tableau")
    assert txt.startswith("\nInitial Tableau:\r\n" + "-" * 80 + "\r\nTable\tx1\tt1\tRHS\r\nZ\t0.000\t-1.000\t2.000\t\r\n1\t1.000\t0.500\t3.000\t\r\n")
    lab = TableIterationFormater.Format(tab, 2, "Final Table", ["x2"])  # fewer labels than rows: index fallback (:41)
    assert lab == R.format_table(tab, 2, "Final Table", ["x2"]) and "\r\nx2\t1.000" in lab and "\r\n2\t0.000\t7.250" in lab


def test_table_format_threaded_path_equals_restatement():
    rng = np.random.default_rng(13)
    T = rng.normal(size=(301, 257)) * 10.0 ** rng.integers(-4, 6, (301, 257))  # > 16k cells: several formatter threads
    txt = TableIterationFormater.Format(T, 100, "Iteration 3 - After pivot")
    assert txt == R.format_table(T.tolist(), 100, "Iteration 3 - After pivot")


# ---- parser fuzz against the independent restatement ------------------------------------------------------------
def _outcome(fn):
    try:
        return ("ok", fn())
    except ValueError:
        return ("format", None)
    except IndexError:
        return ("index", None)


def _native_parse(text):
    m = Model.parse_text(text)
    if not m.info()[0]:
        return None
    return m.problem_type, m.objective(), [(c.Coefficients, c.Relation, c.RHS) for c in m.constraints()], m.signs()


def test_parser_fuzz_against_restatement():
    from hypothesis import given, settings, strategies as st
    tok = st.one_of(
        st.sampled_from(["1", "-2", "+3.5", ".5", "5.", "1e3", "1E-2", "1,000", "1,,0", ",5", "", "x", "1e", "NaN", "Infinity",
                         "-Infinity", "<=", ">=", "=", "\t7", "7\t", "0x10", "1_0", "--1", "1.2.3", "1e999", "-0", "00012"]),
        st.from_regex(r"[+-]?[0-9]{1,4}(\.[0-9]{0,3})?", fullmatch=True))
    line = st.lists(tok, min_size=0, max_size=6).map(" ".join)
    eol = st.sampled_from(["\n", "\r\n", "\r"])

    @settings(max_examples=600, deadline=None)
    @given(st.sampled_from(["max", "MIN", "Max", ""]), st.lists(line, min_size=0, max_size=5), eol, st.booleans(), st.booleans())
    def run(ptype, lines, nl, trailing, pad):
        body = [(ptype + " " + lines[0]) if lines else ptype] + lines[1:]
        if pad:
            body = ["  " + b + " \t" for b in body]
        text = nl.join(body) + (nl if trailing else "")
        ref = _outcome(lambda: R.parse_model(text))
        got = _outcome(lambda: _native_parse(text))
        assert ref[0] == got[0], (text, ref, got)
        if ref[0] == "ok" and ref[1] is not None:
            assert repr(ref[1]) == repr(got[1]), (text, ref, got)  # repr: NaN-safe, keeps the sign of zero
        else:
            assert ref[1] == got[1], (text, ref, got)

    run()


# ---- result files: OutputFileWrite / CanonicalFormForFile ------------------------------------------------------------
def test_general_format_is_double_tostring():
    import ctypes as C
    from lpr_381_group_v22_b200 import _native as N

    def G(x):
        buf = C.create_string_buffer(64)
        N.check(N.lib().lpr_fmt_general(float(x), buf, 64))
        return buf.value.decode()

    known = {0.0: "0", -0.0: "0", 40.0: "40", -11.0: "-11", 0.1: "0.1", 2.5: "2.5", 1e15: "1E+15", 123456789012345.0: "123456789012345",
             0.0001: "0.0001", 0.00001: "1E-05", 1.5e-7: "1.5E-07", 1 / 3: "0.333333333333333", 2 / 3: "0.666666666666667",
             1e300: "1E+300", 15.399999999999999: "15.4", -1234.5678: "-1234.5678", float("nan"): "NaN"}
    for x, s in known.items():
        assert G(x) == s == R.net_general(x), (x, G(x), s)
    rng = np.random.default_rng(14)
    for x in (rng.normal(size=3000) * 10.0 ** rng.integers(-9, 18, 3000)).tolist():
        assert G(x) == R.net_general(x), x


def test_canonical_form_and_result_files(tmp_path):
    text = "max +2 +3 +3 +5 +2 +4\n+11 +8 +6 +14 +10 +10 <= 40\n1 0 -2.5 0 1 0 >= 3\nbin bin bin bin bin bin"
    m = Model.parse_text(text)
    obj = [2, 3, 3, 5, 2, 4]
    cons = [([11, 8, 6, 14, 10, 10], "<=", 40.0), ([1, 0, -2.5, 0, 1, 0], ">=", 3.0)]
    assert m.canonical_form() == R.canonical_form(obj, cons, ["bin"] * 6)
    assert "Z -2x1 -3x2 -3x3 -5x4 -2x5 -4x6 = 0\n+ 11x1 + 8x2" in m.canonical_form()  # negated objective, "+ " prefix
    snaps = ["\nInitial Tableau:\r\n...\r\n", "no newline at the end"]
    x = [0.0, 1.0, 1.0, 1.0, 0.2, 1.0]
    ts = "2025-08-29 10:11:12"
    path = str(tmp_path / "sub" / "dir" / "output_results.txt")  # EnsureDirectory creates the parents
    L.io.OutputFileWrite.WriteFullResults(path, "Primal Simplex Algorithm", m, snaps, 15.4, x, append=False, timestamp=ts)
    want = R.full_results_text("Primal Simplex Algorithm", "max", obj, cons, ["bin"] * 6, snaps, 15.4, x, ts)
    assert open(path, "rb").read() == b"\xef\xbb\xbf" + want.encode("utf-8")  # File.WriteAllText(..., Encoding.UTF8): BOM
    assert "Z* = 15.4\r\nx1 = 0\r\nx2 = 1\r\n" in want and "x5 = 0.2\r\n" in want
    # append to existing content: no second BOM; append to a missing file behaves like a fresh write
    L.io.OutputFileWrite.WriteSnapshotsOnly(path, "Branch and Bound Simplex Algorithm", snaps, 15.0, None, append=True, timestamp=ts)
    tail = R.snapshots_only_text("Branch and Bound Simplex Algorithm", snaps, 15.0, None, ts)
    assert open(path, "rb").read() == b"\xef\xbb\xbf" + (want + tail).encode("utf-8")
    assert "no newline at the end\r\n\r\n=== Final Results ===" in tail and "...\r\n\r\n=== Final" not in tail.replace("end\r\n\r\n", "")
    fresh = str(tmp_path / "fresh.txt")
    L.io.OutputFileWrite.WriteSnapshotsOnly(fresh, "s", [], 1.0, [2.0], append=True, timestamp=ts)
    assert open(fresh, "rb").read() == b"\xef\xbb\xbf" + R.snapshots_only_text("s", [], 1.0, [2.0], ts).encode("utf-8")
    L.io.OutputFileWrite.WriteFullResults(fresh, "s", m, [], 0.0, [], timestamp=ts)  # overwrite, no snapshots, empty x
    assert open(fresh, "rb").read() == b"\xef\xbb\xbf" + R.full_results_text("s", "max", obj, cons, ["bin"] * 6, [], 0.0, [], ts).encode("utf-8")
    L.io.OutputFileWrite.WriteSnapshotsOnly(str(tmp_path / "now.txt"), "s", ["a\n"], 1.0, None, append=False)
    import re
    assert re.search(rb"Timestamp: \d{4}-\d\d-\d\d \d\d:\d\d:\d\d\r\n", open(str(tmp_path / "now.txt"), "rb").read())
