"""CPU tests of the host-side mirrors that need no GPU: parser, CLI pre-processing, text formatting."""
import os

import lpr_381_group_v22_b200 as L
from lpr_381_group_v22_b200.utilities import F3, NumFormat, TableIterationFormater

HERE = os.path.dirname(os.path.abspath(__file__))


def test_input_file_parser_fixtures(capsys):
    p = L.InputFileParser()
    p.ReadInputFile(os.path.join(HERE, "golden", "model_a.txt"))
    assert p.ProblemType == "max" and p.ObjectiveCoefficients == [2, 3, 3, 5, 2, 4]
    assert len(p.Constraints) == 1 and p.Constraints[0].Coefficients == [11, 8, 6, 14, 10, 10]
    assert p.Constraints[0].Relation == "<=" and p.Constraints[0].RHS == 40
    assert p.SignRestrictions == ["bin"] * 6
    q = L.InputFileParser()
    q.ReadInputFile(os.path.join(HERE, "golden", "model_b.txt"))
    assert [c.Relation for c in q.Constraints] == ["<=", ">="] and q.SignRestrictions == ["+", "+", "+"]
    L.InputFileParser().ReadInputFile("/nonexistent/file.txt")
    assert "can't find your file" in capsys.readouterr().out


def test_cli_bound_rows_quirk_q1():
    cons = []
    L.add_cli_bound_rows(3, cons)
    assert len(cons) == 3 and all(len(c.Coefficients) == 6 for c in cons)
    assert cons[1].Coefficients == [0, 1, 0, 0, 1, 0] and cons[1].Relation == "<=" and cons[1].RHS == 1
    cons = []
    L.add_upper_bound_constraints(3, ["bin", "+", "0<=x<=1"], cons)
    assert [c.Coefficients for c in cons] == [[1, 0, 0], [0, 0, 1]]
    assert L.add_upper_bound_constraints(2, [], []) == []


def test_net_style_formatting():
    assert F3(1.0) == "1.000" and F3(0.0005) == "0.001" and F3(-0.0001) == "0.000" and F3(-2.5) == "-2.500"
    assert F3(15.399999999999999) == "15.400" and F3(2.0005) == "2.001"
    assert NumFormat.N3(1e-13) == "0" and NumFormat.N3(2.0) == "2" and NumFormat.N3(0.125) == "0.125"
    assert NumFormat.N3(0.2000000000000001) == "0.2" and NumFormat.N3(-1.9999999) == "-2"
    txt = TableIterationFormater.Format([[0.0, -1.0, 2.0], [1.0, 0.5, 3.0]], 1, "Initial Tableau")
    lines = txt.split("\r\n")
    assert lines[0] == "\nInitial Tableau:" or lines[0] == "" or "Initial Tableau:" in txt
    assert "Table\tx1\tt1\tRHS" in txt and "Z\t0.000\t-1.000\t2.000\t" in txt and "1\t1.000\t0.500\t3.000\t" in txt


def test_division_free_round4_tail(tmp_path):
    """tools/div1e4_check.c: the fma-based k / 1e4 of the B&B kernels (common.cuh net_round4) equals the IEEE quotient
    and rounding twice equals rounding once.  The full claim (every |k| <= 2^31) takes ~40 core-seconds and was run
    when the kernels were written; here the first 2^23 integers of both signs plus the top of the range."""
    import subprocess
    exe = str(tmp_path / "div1e4_check")
    src = os.path.join(os.path.dirname(HERE), "tools", "div1e4_check.c")
    subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fopenmp", src, "-lm", "-o", exe])
    for args in (["8388608"], ["2147483648", "2139095040"]):
        res = subprocess.run([exe] + args, capture_output=True, text=True, timeout=300)
        assert res.returncode == 0 and "mismatches 0, idempotence mismatches 0" in res.stdout, res.stdout
