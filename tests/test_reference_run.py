"""CPU tests: the oracle (oracle/lpr_oracle.cpp) and the native host code against OUTPUTS OF THE REFERENCE'S OWN C#
SOURCES, executed in the build container by the interpreter under oracle/csharp/ and committed as
tests/golden/reference_run.json (generator: tests/golden/make_reference_run.py).

This is what pins the oracle: every value below was produced by running Storm-Tarran/LPR_381_Group_V22's unmodified
.cs files, not by reading them.  The GPU leg of the same comparison is tests/test_reference_run_gpu.py.
"""
import hashlib
import json
import os
import subprocess
import sys

import numpy as np
import pytest

import oracle_lib as O

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = json.load(open(os.path.join(HERE, "golden", "reference_run.json")))
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def unmat(m):
    return np.array([float.fromhex(h) for h in m["hex"]], dtype=np.float64).reshape(m["shape"])


def unhex(v):
    return np.array([float.fromhex(h) for h in v], dtype=np.float64)


def same_bits(a, b):
    a = np.ascontiguousarray(a, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    return a.shape == b.shape and np.array_equal(a.view(np.uint64), b.view(np.uint64))


def cons_of(rec):
    return [(co, rel, rhs) for co, rel, rhs in rec["constraints"]]


# ------------------------------------------------------------------------------------------- the golden file itself
def test_golden_covers_the_reference_fixtures_and_every_solver():
    assert GOLD["meta"]["generator"] == "tests/golden/make_reference_run.py"
    assert len(GOLD["meta"]["reference_sources"]) == 13
    for key, least in (("parser", 9), ("primal", 28), ("primal2", 16), ("dual", 16), ("cutting_plane", 12),
                       ("revised", 20), ("bb", 13), ("bb_formulate", 6), ("sensitivity", 12), ("sensitivity_rhs", 6),
                       ("output", 3), ("mid_size", 3), ("bb_parts", 40), ("program", 6), ("cutting_plane_ties", 8),
                       ("run_bb", 5), ("accessors", 12)):
        assert len(GOLD[key]) >= least, key
    # data/TextFile.txt parsed by the reference's own InputFileParser
    p = GOLD["parser"][0]
    assert p["problem_type"] == "max" and unhex(p["objective"]).tolist() == [2, 3, 3, 5, 2, 4]
    assert p["signs"] == ["bin"] * 6 and p["constraints"][0][1] == "<="


@pytest.mark.skipif(not os.path.isdir("/root/reference/LPR_381_Group_V22"), reason="the reference tree is only in the build container")
def test_golden_is_what_the_reference_sources_produce_today():
    r = subprocess.run([sys.executable, os.path.join(HERE, "golden", "make_reference_run.py"), "--check"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr


# ------------------------------------------------------------------------------------------- primal tableau simplex
@pytest.mark.parametrize("i", range(len(GOLD["primal"])))
def test_primal_simplex_solver(i):
    g = GOLD["primal"][i]
    T0, basis0 = O.primal_build(g["objective"], cons_of(g), g["is_max"])
    assert same_bits(T0, unmat(g["initial_tableau"]))
    r = O.primal_solve(T0, basis0)
    assert r["log"].tolist() == g["pivots"]
    assert r["status"] == (O.OPTIMAL if g["status"] == "optimal" else O.UNBOUNDED)
    assert same_bits(r["T"], unmat(g["final_tableau"])) and same_bits(r["T"], unmat(g["get_final_tableau"]))
    assert r["basis"].tolist() == g["basis"]
    if g["status"] == "optimal":
        assert float(r["T"][0, -1]).hex() == g["final_z"]
        assert same_bits(O.primal_extract(r["T"], len(g["objective"])), unhex(g["x"]))
    else:
        assert g["x"] is None and float.fromhex(g["final_z"]) == 0.0   # FinalZ / SolutionVector are never set


@pytest.mark.parametrize("i", range(len(GOLD["primal2"])))
def test_primal_simplex_solver2(i):
    g = GOLD["primal2"][i]
    r = O.primal2_solve(unmat(g["tableau"]), g["max_iters"], g["print_steps"])
    assert same_bits(r["T"], unmat(g["final_tableau"]))
    assert (r["status"] == O.OPTIMAL) == g["returned"]
    assert r["log"].tolist() == g["pivots"]
    if g["returned"]:
        assert float(r["T"][0, -1]).hex() == g["final_z"]


@pytest.mark.parametrize("i", range(len(GOLD["dual"])))
def test_dual_simplex_solver(i):
    g = GOLD["dual"][i]
    r = O.dual_solve(unmat(g["tableau"]), g["max_iters"], g["print_steps"])
    assert same_bits(r["T"], unmat(g["final_tableau"]))
    if g["exception"] is not None:
        assert g["exception"] == "InvalidOperationException" and r["status"] == O.PIVOT_TOO_SMALL
    else:
        assert (r["status"] == O.OPTIMAL) == g["returned"]
    if g["print_steps"]:   # the reference prints "constraint {pivotRow + 1}" (DualSimplex.cs:94) = the tableau row
        assert r["log"].tolist() == g["printed_pivots"]


CUT_END = {"Displayed the Optimal Tableau.": O.OPTIMAL, "All RHS are integers. No Gomory cut needed.": O.NO_CUT_NEEDED,
           "No valid pivot column on the cut (need a negative cut coeff with non-zero obj coeff).": O.NO_PIVOT_COL,
           "Dual Simplex failed (infeasible or max iters).": None, "Pivot too small/zero.": O.PIVOT_TOO_SMALL,
           "Cutting-plane step finished (further steps may be required).": O.CUT_STEP_DONE}


@pytest.mark.parametrize("i", range(len(GOLD["cutting_plane"])))
def test_cutting_plane_solver(i):
    g = GOLD["cutting_plane"][i]
    T = unmat(g["tableau"])
    r = O.cutting_plane(T, extra_rows=64, literal_sort=True)
    assert same_bits(r["T"], unmat(g["final_tableau"]))
    # a cut that finds no pivot column is appended all the same (CuttingPlaneSolver.cs:107-131)
    appended = unmat(g["final_tableau"]).shape[0] - T.shape[0]
    assert appended in (len(g["cut_pivots_1based"]), len(g["cut_pivots_1based"]) + 1)
    got = [[T.shape[0] + k, int(c) + 1] for k, c in enumerate(r["log"][:len(g["cut_pivots_1based"]), 1])]
    assert got == g["cut_pivots_1based"]          # "row {cutRowIdx + 1}, col {pivotCol + 1}" (:136)
    want = CUT_END[g["last_console_line"]]
    if want is None:
        assert r["status"] in (O.INFEASIBLE, O.ITER_LIMIT)
    else:
        assert r["status"] == want


def test_gomory_row_choice_follows_list_sort():
    """CuttingPlaneSolver.cs:94-96 sorts the fractional rows with List<T>.Sort and takes element 0.  That sort is the
    Framework's UNSTABLE introspective sort: with more than 16 fractional rows and exact ties for the best key the row
    it leaves in front is not the first minimum.  The oracle restates the sort literally and reproduces the executed
    reference on tie-heavy tableaux; the plain first minimum -- what the CUDA kernels implement, DESIGN.md section 2
    "known deviation" -- does not, on the same inputs."""
    before = O.gomory_tie_corners()
    first_min_differs = 0
    for g in GOLD["cutting_plane_ties"]:
        T = unmat(g["tableau"])
        want = unmat(g["final_tableau"])
        r = O.cutting_plane(T, extra_rows=128, literal_sort=True)
        assert same_bits(r["T"], want)
        got = [[T.shape[0] + k, int(c) + 1] for k, c in enumerate(r["log"][:len(g["cut_pivots_1based"]), 1])]
        assert got == g["cut_pivots_1based"]
        first_min_differs += not same_bits(O.cutting_plane(T, extra_rows=128)["T"], want)
    assert O.gomory_tie_corners() > before
    assert first_min_differs >= 2         # the corner is real: the stable choice leaves the reference's path


# ------------------------------------------------------------------------------------------- revised simplex
@pytest.mark.parametrize("i", range(len(GOLD["revised"])))
def test_revised_primal_simplex_solver(i):
    g = GOLD["revised"][i]
    n = len(g["c"])
    r = O.rev_solve(np.array(g["A"], dtype=float), g["b"], g["c"], g["is_min"], want_binv=True)
    labels = [f"x{e + 1}" if e < n else f"S{e - n + 1}" for e in r["log"][:, 1].tolist()]
    assert labels == g["entering_labels"] and r["n_iter"] == g["n_iterations"]
    assert r["basis"].tolist() == g["basis"]
    assert same_bits(r["Binv"], unmat(g["binv"]))
    if g["exception"] is None:
        assert r["status"] == O.OPTIMAL
        assert same_bits(r["xB"], unhex(g["xb"])) and same_bits(r["x"], unhex(g["x"]))
        assert float(r["z"]).hex() == g["final_z"]
    elif g["exception"].startswith("Unbounded"):
        assert r["status"] == O.UNBOUNDED
    elif g["exception"].startswith("Infeasible"):
        assert r["status"] == O.INFEASIBLE
    else:
        assert r["status"] == O.PIVOT_TOO_SMALL


def test_revised_snapshot_text_restatement():
    """tests/net_reference.py (the checker of the native CaptureSnapshot text, b9) against the text the reference's
    own CaptureSnapshot produced"""
    import net_reference as R
    for g in GOLD["revised"]:
        if g["exception"] is not None:
            continue
        snaps, *_ = R.revised_solve_with_snapshots(g["c"], g["A"], g["b"], g["is_min"])
        text = "".join(snaps)
        assert len(snaps) == g["n_snapshots"]
        assert hashlib.sha256(text.encode("utf-8")).hexdigest() == g["snapshots_sha256"]
        if "snapshots" in g:
            assert snaps == g["snapshots"]


# ------------------------------------------------------------------------------------------- B&B simplex
@pytest.mark.parametrize("i", range(len(GOLD["bb"])))
def test_branch_and_bound_adapter(i):
    g = GOLD["bb"][i]
    r = O.bb_solve(unmat(g["root_tableau"]), g["n_vars"], prune=g["enable_pruning"], max_nodes=20)
    nodes = g["nodes"]
    assert r["nodes"] == len(nodes)
    log = r["node_log"]
    assert log[:, 0].tolist() == [nd["depth"] for nd in nodes]
    assert log[:, 3].astype(bool).tolist() == [nd["pruned"] for nd in nodes]
    for k, nd in enumerate(nodes):
        if not nd["pruned"]:
            assert int(log[k, 1]) == nd["branch_var"], (k, nd)
            assert bool(log[k, 2]) == nd["integer"], (k, nd)
    x, z = unhex(g["x"]), float.fromhex(g["z"])
    if r["has_solution"]:
        assert same_bits(r["x"], x) and float(r["z"]).hex() == g["z"]
    else:
        assert x.size == 0 and z == float("-inf")     # (x ?? new List<double>(), z) with z = -infinity
    assert (r["status"] == O.NODE_LIMIT) == g["hit_node_cap"]


@pytest.mark.parametrize("i", range(len(GOLD["bb_formulate"])))
def test_dual_simplex_solver_bb_formulate_and_solve(i):
    g = GOLD["bb_formulate"][i]
    assert "error" not in g
    T = O.bb_formulate(g["objective"], g["rows"])
    assert same_bits(T, unmat(g["tableau"]))
    r = O.bb_node_solve_ex(T, g["is_min"])
    if g["optimal_value"] is None:
        assert r["status"] == O.INFEASIBLE
    else:
        assert r["status"] == O.OPTIMAL
        assert same_bits(r["T"], unmat(g["final_tableau"]))
        assert float(r["T"][0, -1]).hex() == g["optimal_value"]
        assert r["log"][:, 0].tolist() == g["pivot_rows"] and r["log"][:, 1].tolist() == g["pivot_cols"]


@pytest.mark.parametrize("i", range(len(GOLD["bb_parts"])))
def test_branch_and_bound_members_one_by_one(i):
    """RoundTableau, IdentifyBasicVariables, AddConstraint, CheckIntegerBasicVar, ExtractSolution, PerformDualPivot and
    PerformPrimalPivot (both senses) of the reference, executed on adversarial tableaux: columns that sum to 1 without
    being unit columns, entries a hair off 0 and 1, rounding ties, negative right-hand sides"""
    g = GOLD["bb_parts"][i]
    T = unmat(g["tableau"])
    Tr = O.bb_round(T)
    assert same_bits(Tr, unmat(g["rounded"]))
    assert O.bb_identify_basic(Tr).tolist() == g["identify_basic"]
    assert same_bits(O.bb_add_constraint(T, g["n_vars"], g["var"], g["bound"], g["type"]), unmat(g["add_constraint"]))
    var, val = O.bb_branch_var(Tr, g["n_vars"])
    assert var == g["branch_var"] and (var < 0 or float(val).hex() == g["branch_value"])
    assert same_bits(O.bb_extract(Tr, g["n_vars"]), unhex(g["extract"]))
    ok, out, _, _ = O.bb_dual_pivot(Tr)
    assert bool(ok) == (g["dual_pivot"] is not None) and (not ok or same_bits(out, unmat(g["dual_pivot"])))
    ok, out, _, _ = O.bb_primal_pivot(Tr)
    assert bool(ok) == (g["primal_pivot"] is not None) and (not ok or same_bits(out, unmat(g["primal_pivot"])))
    # isMinimization = true has no single-pivot function in the oracle: one pivot of its DoDualSimplex state machine is
    # the same pivot when no RHS is negative before or after it (no dual phase, no "drop the last tableau"), up to the
    # -0.0 -> 0.0 clean-up DoDualSimplex applies to every tableau (:307-313)
    if g["primal_pivot_min"] is not None and not (Tr[1:, -1] < -1e-9).any():
        want = unmat(g["primal_pivot_min"])
        if not (want[:, -1] < 0).any():
            r = O.bb_node_solve_ex(Tr, True, max_pivots=1)
            assert r["n_pivots"] == 1 and same_bits(r["T"] + 0.0, want + 0.0)


@pytest.mark.parametrize("i", range(len(GOLD["run_bb"])))
def test_run_branch_and_bound_entry(i):
    """BranchAndBound.RunBranchAndBound executed (its result is only printed: parsed from the reference's console)
    against the oracle's composite ConfigureProblem -> FormulateTableau -> DoDualSimplex -> RoundTableau ->
    ExecuteBranchAndBound"""
    g = GOLD["run_bb"][i]
    assert g["exception"] is None
    o2, c2 = O.bb_configure_problem(g["objective"], g["rows"])
    T = O.bb_formulate(o2, c2)
    lp = O.bb_node_solve_ex(T, g["is_min"])
    assert lp["status"] == O.OPTIMAL
    root = O.bb_round(lp["T"])
    import net_reference as R
    assert R.net_general(float(root[0, -1])) == g["initial_objective_text"]
    r = O.bb_solve(root, len(g["objective"]), prune=False, max_nodes=20)
    assert r["node_log"][:, 0].tolist() == g["node_depths"]
    if g["no_integer_solution"]:
        assert not r["has_solution"]
    else:
        assert r["has_solution"]
        assert ", ".join(R.net_general(float(v)) for v in r["x"]) == g["solution_text"]
        assert R.net_general(float(r["z"])) == g["value_text"]


def test_rounding_rules():
    """Math.Round(x, 4) as BranchAndBound.RoundNumber calls it, IsInteger, CuttingPlaneSolver.Frac"""
    g = GOLD["rounding"]
    vals = unhex(g["values"])
    lib = O.lib()
    assert [float(lib.orc_net_round4(float(v))).hex() for v in vals] == g["round4"]
    assert [float(lib.orc_frac(float(v))).hex() for v in vals] == g["frac"]
    assert [abs(lib.orc_net_round4(float(v)) - lib.orc_net_round(lib.orc_net_round4(float(v)))) <= 1e-6 for v in vals] == g["is_integer"]
    # COMDouble::Round against rint: one double apart (x + 0.5 rounds to exactly 1.0, then the odd-tie correction)
    assert lib.orc_net_round(0.5000000000000001) == 0.0 and np.rint(0.5000000000000001) == 1.0
    assert lib.orc_net_round4(5.000000000000001e-05) == 0.0 and lib.orc_net_round4(5.0000000000000016e-05) == 0.0
    rng = np.random.default_rng(0)
    xs = np.concatenate([rng.normal(size=200000) * 10.0 ** rng.integers(-8, 16, 200000),
                         rng.integers(-10**6, 10**6, 200000) + 0.5, np.nextafter(np.arange(0, 2000) + 0.5, np.inf)])
    xs = xs[xs != 0.5000000000000001]
    got = np.array([lib.orc_net_round(float(v)) for v in xs[:60000]])
    assert np.array_equal(got, np.rint(xs[:60000]))


# ------------------------------------------------------------------------------------------- sensitivity
def _sens_root(g):
    T0, b0 = O.primal_build(g["objective"], cons_of(g), True)
    opt = O.primal_solve(T0, b0)
    assert opt["status"] == O.OPTIMAL
    return opt["T"]


@pytest.mark.parametrize("i", range(len(GOLD["sensitivity"])))
def test_sensitivity_add_constraint(i):
    g = GOLD["sensitivity"][i]
    T = _sens_root(g)
    assert same_bits(T, unmat(g["final_tableau"]))
    n = len(g["objective"])
    basis = O.sens_rebuild_basis(T)
    assert basis.tolist() == g["basis_rebuilt"]
    x = O.primal_extract(T, n)
    ax = 0.0
    for j in range(min(len(g["tech"]), n)):
        ax += g["tech"][j] * x[j]
    T1, b1 = O.sens_add_constraint(T, basis, g["tech"], g["rhs"] - ax)
    res = O.sens_resolve(T1, O.sens_rebuild_basis(T1))
    assert same_bits(res["T"], unmat(g["tableau_after"]))
    assert res["basis"].tolist() == g["basis_after"]
    if g["exception"] is None:
        assert res["status"] == O.OPTIMAL
        assert float(res["T"][0, -1]).hex() == g["z_after"]
        assert same_bits(O.sens_solution(res["T"]), unhex(g["x_after"]))
    else:
        assert res["status"] in (O.INFEASIBLE, O.UNBOUNDED, O.ITER_LIMIT)


@pytest.mark.parametrize("i", range(len(GOLD["sensitivity_rhs"])))
def test_sensitivity_resolve_after_rhs_change(i):
    g = GOLD["sensitivity_rhs"][i]
    T = unmat(g["tableau_before_resolve"])
    res = O.sens_resolve(T, O.sens_rebuild_basis(T))
    assert same_bits(res["T"], unmat(g["tableau_after"]))
    assert res["basis"].tolist() == g["basis_after"]
    if g["exception"] is None:
        assert res["status"] == O.OPTIMAL
        assert float(res["T"][0, -1]).hex() == g["z_after"]
        assert same_bits(O.sens_solution(res["T"]), unhex(g["x_after"]))
    else:
        assert res["status"] in (O.INFEASIBLE, O.UNBOUNDED, O.ITER_LIMIT)


# ------------------------------------------------------------------------------------------- native host code
def test_native_parser_against_the_reference_parser():
    from lpr_381_group_v22_b200.io import Model
    for g in GOLD["parser"]:
        if g["exception"] == "FormatException":
            with pytest.raises(ValueError, match="FormatException"):
                Model.parse_text(g["text"])
            continue
        if g["exception"] == "IndexOutOfRangeException":
            with pytest.raises(IndexError, match="IndexOutOfRangeException"):
                Model.parse_text(g["text"])
            continue
        m = Model.parse_text(g["text"])
        assert m.message == g["console"].strip()
        if g["problem_type"] is None:
            assert m.info()[0] is False
            continue
        assert m.problem_type == g["problem_type"]
        assert same_bits(m.objective(), unhex(g["objective"]))
        cons = m.constraints()
        assert len(cons) == len(g["constraints"])
        for c, (co, rel, rhs) in zip(cons, g["constraints"]):
            assert same_bits(c.Coefficients, unhex(co)) and c.Relation == rel and float(c.RHS).hex() == rhs
        assert m.signs() == g["signs"]


def test_python_mirror_of_the_input_classes_against_the_reference(tmp_path, capsys):
    """lpr_381_group_v22_b200.InputFileParser (class surface of IO/InputFileParser.cs) and the list helpers of
    Program.cs's bound rows against what the reference's classes did with the same texts"""
    import lpr_381_group_v22_b200 as L
    for k, g in enumerate(GOLD["parser"]):
        path = tmp_path / f"m{k}.txt"
        path.write_bytes(g["text"].encode("utf-8"))
        p = L.InputFileParser()
        if g["exception"] is not None:
            with pytest.raises((ValueError, IndexError)):
                p.ReadInputFile(str(path))
            continue
        p.ReadInputFile(str(path))
        assert capsys.readouterr().out.strip() == g["console"].strip()
        assert p.ProblemType == g["problem_type"]
        if g["problem_type"] is None:
            continue
        assert same_bits(p.ObjectiveCoefficients, unhex(g["objective"])) and list(p.SignRestrictions) == g["signs"]
        for c, (co, rel, rhs) in zip(p.Constraints, g["constraints"]):
            assert same_bits(c.Coefficients, unhex(co)) and c.Relation == rel and float(c.RHS).hex() == rhs
    for g in GOLD["output"]:
        p = L.InputFileParser()
        path = tmp_path / "o.txt"
        path.write_bytes(g["text"].encode("utf-8"))
        p.ReadInputFile(str(path))
        cons = list(p.Constraints)
        before = len(cons)
        if g["add_upper_bound_rows"]:
            L.add_upper_bound_constraints(len(p.ObjectiveCoefficients), list(p.SignRestrictions), cons)
        assert [[[float(v).hex() for v in c.Coefficients], c.Relation, float(c.RHS).hex()] for c in cons[before:]] == g["rows_added"]


def test_native_number_formatting_against_the_interpreted_formatter():
    """three independent statements of the .NET Framework rules -- csrc/host_io.cu, tests/net_reference.py and
    oracle/csharp/csrun.py (which ran the reference's NumFormat.N3 / TableIterationFormater.Format) -- agree"""
    import net_reference as R
    from lpr_381_group_v22_b200.utilities import NumFormat, TableIterationFormater
    g = GOLD["format"]
    vals = unhex(g["values"])
    assert [NumFormat.N3(float(v)) for v in vals] == g["n3"]
    assert [R.N3(float(v)) for v in vals] == g["n3"]
    k = len(vals) - len(vals) % 4
    tab = vals[:k].reshape(-1, 4)
    assert TableIterationFormater.Format(tab, 2, "T") == g["table"]
    assert R.format_table(tab.tolist(), 2, "T") == g["table"]


def test_native_snapshot_text_of_the_primal_solver():
    """IterationSnapshots of PrimalSimplexSolver as the reference built them, against the native formatter fed with
    the oracle's tableaux (the GPU path produces the same tableaux bit for bit, tests/test_reference_run_gpu.py)"""
    from lpr_381_group_v22_b200.utilities import TableIterationFormater
    for g in GOLD["primal"]:
        if "snapshots" not in g:
            continue
        n = len(g["objective"])
        T, basis = O.primal_build(g["objective"], cons_of(g), g["is_max"])
        texts = [TableIterationFormater.Format(T, n, "Initial Tableau")]
        for k, (pr, pc) in enumerate(g["pivots"]):
            O.lib().orc_primal_pivot(T.shape[0], T.shape[1], T.ctypes.data_as(O._dp), pr, pc, 1)
            texts.append(TableIterationFormater.Format(T, n, f"Iteration {k + 1} - After pivot"))
        assert texts == g["snapshots"][:-1]
        assert g["snapshots"][-1].startswith(TableIterationFormater.Format(T, n, "Final Tableau (Optimal)"))


def test_native_result_files_against_output_file_write(tmp_path):
    """OutputFileWrite.WriteFullResults / WriteSnapshotsOnly, CanonicalFormConverter.CanonicalFormForFile and
    Program.AddUpperBoundConstraints as the reference executed them, against lpr_out_* / lpr_model_* (csrc/host_io.cu)"""
    import lpr_381_group_v22_b200 as L
    from lpr_381_group_v22_b200.io import Model
    for k, g in enumerate(GOLD["output"]):
        m = Model.parse_text(g["text"])
        before = len(m.constraints())
        if g["add_upper_bound_rows"]:
            m.add_upper_bound_rows()
        added = m.constraints()[before:]
        assert len(added) == len(g["rows_added"])
        for c, (co, rel, rhs) in zip(added, g["rows_added"]):
            assert same_bits(c.Coefficients, unhex(co)) and c.Relation == rel and float(c.RHS).hex() == rhs
        assert m.canonical_form() == g["canonical_form"]
        path = str(tmp_path / f"out{k}" / "output_results.txt")
        x = unhex(g["x"]).tolist()
        z = float.fromhex(g["final_z"])
        L.io.OutputFileWrite.WriteFullResults(path, g["solver"], m, g["snapshots"], z, x, append=False,
                                              timestamp=g["timestamp"])
        assert open(path, "rb").read() == g["file_after_full_results"].encode("utf-8")
        L.io.OutputFileWrite.WriteSnapshotsOnly(path, g["solver"] + " (again)", g["snapshots"], z, x, append=True,
                                                timestamp=g["timestamp"])
        assert open(path, "rb").read() == g["file_after_append"].encode("utf-8")
        assert g["file_after_append"].count("\ufeff") == 1     # one byte-order mark, at the start


@pytest.mark.parametrize("i", range(len(GOLD["mid_size"])))
def test_mid_size_runs(i):
    """longer pivot sequences (12-15 pivots on 19x43 .. 41x69 tableaux): digests of what the reference computed"""
    g = GOLD["mid_size"][i]
    cons = cons_of(g)
    n = len(g["objective"])
    pg, rg = g["primal"], g["revised"]
    T0, b0 = O.primal_build(g["objective"], cons, True)
    assert hashlib.sha256(T0.tobytes()).hexdigest() == pg["initial_tableau_sha256"]
    r = O.primal_solve(T0, b0)
    assert r["log"].tolist() == pg["pivots"] and r["basis"].tolist() == pg["basis"]
    assert list(r["T"].shape) == pg["final_tableau_shape"]
    assert hashlib.sha256(r["T"].tobytes()).hexdigest() == pg["final_tableau_sha256"]
    assert float(r["T"][0, -1]).hex() == pg["final_z"] and same_bits(O.primal_extract(r["T"], n), unhex(pg["x"]))
    rr = O.rev_solve(np.array([c[0] for c in cons], dtype=float), [c[2] for c in cons], g["objective"], False, want_binv=True)
    labels = [f"x{e + 1}" if e < n else f"S{e - n + 1}" for e in rr["log"][:, 1].tolist()]
    assert labels == rg["entering_labels"] and rr["basis"].tolist() == rg["basis"]
    assert hashlib.sha256(rr["Binv"].tobytes()).hexdigest() == rg["binv_sha256"]
    assert float(rr["z"]).hex() == rg["final_z"] and same_bits(rr["x"], unhex(rg["x"])) and same_bits(rr["xB"], unhex(rg["xb"]))


def test_program_main_sessions():
    """the reference's whole application (Program.Main, scripted keyboard) on its shipped model: the answers of SURVEY
    Appendix C as the application itself printed and wrote them, and the state leak between menu options"""
    prog = {tuple(g["keys"]): g for g in GOLD["program"]}
    one = prog[("1", "13", "7")]
    assert one["unhandled_exception"] is None and one["keys_left"] == []
    assert "Optimal Solution Found!" in one["console"] and "Warning: x5 = 0.2 violates binary constraint." in one["console"]
    assert one["output_file"].endswith("=== Final Results ===\r\nZ* = 15.4\r\nx1 = 0\r\nx2 = 1\r\nx3 = 1\r\nx4 = 1\r\nx5 = 0.2\r\nx6 = 1\r\n")
    r = final_a_oracle()
    assert float(r["T"][0, -1]).hex() == float(15.4).hex() or abs(r["T"][0, -1] - 15.4) < 1e-12
    two = prog[("2", "7")]
    assert "Solver: Revised Primal Simplex Algorithm (T-*)" in two["output_file"] and "Z* = 15.4" in two["output_file"]
    three = prog[("3", "7")]
    assert "Total branchs processed: 20" in three["console_tail"] and "Z* = 15\r\nx1 = 0\r\nx2 = 1\r\nx3 = 1\r\nx4 = 1\r\nx5 = 0\r\nx6 = 1" in three["console_tail"]
    assert prog[("4", "7")]["output_file"] is None            # option 4 only prints the canonical form (Program.cs:420-429)
    leak = prog[("1", "13", "2")]
    assert leak["unhandled_exception"] == "System.ArgumentException: Constraint 2 has incorrect number of coefficients."


def test_native_cli_bound_rows_against_program_main():
    """the x_i <= 1 rows menu option 1 appends inline (Program.cs:114-124, one entry too long with a stray 1: SURVEY
    Q1), as Main itself built them and OutputFileWrite printed them, against lpr_model_add_cli_bound_rows"""
    from lpr_381_group_v22_b200.io import Model
    g = {tuple(r["keys"]): r for r in GOLD["program"]}[("1", "13", "7")]
    m = Model.parse_text(g["model"])
    m.add_cli_bound_rows()
    canon = m.canonical_form()
    assert "+ 1x1 + 0x2 + 0x3 + 0x4 + 0x5 + 0x6 + 0x7 + 1x8 + 0x9 + S2 = 1" in canon
    assert canon in g["output_file"]


def final_a_oracle():
    rows = []
    for i in range(6):
        co = [0.0] * 9
        co[i] = 1.0
        co[7] = 1.0
        rows.append((co, "<=", 1.0))
    T, b = O.primal_build([2, 3, 3, 5, 2, 4], [([11, 8, 6, 14, 10, 10], "<=", 40)] + rows)
    return O.primal_solve(T, b)


def test_longer_runs():
    """longer runs, 61 x 161 and 101 x 301 tableaux (tests/golden/reference_run_large.json, written once by
    `make_reference_run.py --large`): pivot sequence, basis, digests of the final tableau and of B^-1, z, x"""
    large = json.load(open(os.path.join(HERE, "golden", "reference_run_large.json")))
    assert large["meta"]["reference_sources_sha256"] == GOLD["meta"]["reference_sources_sha256"]
    for g in large["runs"]:
        cons = cons_of(g)
        n = len(g["objective"])
        pg, rg = g["primal"], g["revised"]
        T0, b0 = O.primal_build(g["objective"], cons, True)
        assert hashlib.sha256(T0.tobytes()).hexdigest() == pg["initial_tableau_sha256"]
        r = O.primal_solve(T0, b0)
        assert len(pg["pivots"]) >= 70 and r["log"].tolist() == pg["pivots"] and r["basis"].tolist() == pg["basis"]
        assert hashlib.sha256(r["T"].tobytes()).hexdigest() == pg["final_tableau_sha256"]
        assert float(r["T"][0, -1]).hex() == pg["final_z"] and same_bits(O.primal_extract(r["T"], n), unhex(pg["x"]))
        if rg is None:
            continue
        rr = O.rev_solve(np.array([c[0] for c in cons], dtype=float), [c[2] for c in cons], g["objective"], False, want_binv=True)
        labels = [f"x{e + 1}" if e < n else f"S{e - n + 1}" for e in rr["log"][:, 1].tolist()]
        assert labels == rg["entering_labels"] and rr["basis"].tolist() == rg["basis"]
        assert hashlib.sha256(rr["Binv"].tobytes()).hexdigest() == rg["binv_sha256"]
        assert float(rr["z"]).hex() == rg["final_z"] and same_bits(rr["x"], unhex(rg["x"]))


# ------------------------------------------------------------------------------------------- the interpreter itself
def _run(src, cls="T", method="F", *args):
    from csharp import Interpreter
    it = Interpreter()
    it.load_source(src)
    return it, it.call_static(cls, method, *args)


def test_interpreter_numeric_typing_rules():
    src = """
    using System; using System.Collections.Generic; using System.Linq;
    public static class T {
        public static double Half(int a, int b) { double s = 0; s += a / b; return s; }          // int / int truncates
        public static double Conv() { double x = 7; return x / 2; }                              // int -> double on store
        public static double Tern(bool c) { return (c ? 1 : 2.5) / 2; }                         // ?: unifies to double
        public static int Rem() { return (-7 / 2) * 10 + (-7 % 3); }
        public static double ListConv() { var l = new List<double> { 1, 2 }; l.Add(3); return l[2] / 2; }
        public static double NegZero() { double z = -0.0; return 1 / z; }
        public static int Cast(double d) { return (int)d; }
        public static double Prec() { return 1 + 2 * 3 - 4 / 8.0 % 3; }
        public static string Interp(double v, int w) { return $"[{v,8:F3}|{w:D3}|{v:0.##}|{(w > 2 ? "a:b" : "c")}]"; }
    }"""
    from csharp import Interpreter
    it = Interpreter()
    it.load_source(src)
    c = lambda m, *a: it.call_static("T", m, *a)
    assert c("Half", 7, 2) == 3.0 and c("Conv") == 3.5 and c("Tern", True) == 0.5 and c("Tern", False) == 1.25
    assert c("Rem") == -31 and c("ListConv") == 1.5 and c("NegZero") == float("-inf")
    assert c("Cast", -2.9) == -2 and c("Cast", float("nan")) == -2147483648 and c("Prec") == 6.5
    assert c("Interp", 3.14159, 7) == "[   3.142|007|3.14|a:b]"


def test_interpreter_bcl_rules():
    from csharp.csrun import CsException, format_double, introsort, math_round, math_round_digits
    # Math.Round: banker's at the integer, the Framework's scale / round / unscale at digits
    assert [math_round(v) for v in (0.5, 1.5, 2.5, -0.5, -1.5, 0.49999999999999994)] == [0, 2, 2, -0.0, -2, 0]
    assert math_round_digits(2.00005, 4) == 2.0 and math_round_digits(2.00015, 4) == 2.0002
    assert math_round_digits(1.0005, 3, away=True) == 1.001 and math_round_digits(-1.0005, 3, away=True) == -1.001
    # number formatting: 15 significant digits, then half up on the digit string
    assert format_double(0.1 + 0.2) == "0.3" and format_double(1e15) == "1E+15" and format_double(123456789012345.0) == "123456789012345"
    assert format_double(2.5, "F0") == "3" and format_double(-0.0004, "F3") == "0.000" and format_double(1.0005, "F3") == "1.001"
    assert format_double(0.00001) == "1E-05" and format_double(-1234.5, "N1") == "-1,234.5"
    assert format_double(12.3456, "0.###") == "12.346" and format_double(5.0, "0.###") == "5" and format_double(0.5, "0.00") == "0.50"
    assert format_double(0.1, "R") == "0.1" and format_double(1 / 3, "R") == "0.33333333333333331"
    # LINQ
    src = """
    using System; using System.Collections.Generic; using System.Linq;
    public static class T {
        public static string F() {
            var l = new List<double> { 3, double.NaN, 1 };
            var s = "";
            s += l.Min() + "|" + l.Max() + "|" + l.IndexOf(double.NaN) + "|" + l.Contains(-0.0 + 1);
            try { new List<double>().Min(); } catch (InvalidOperationException) { s += "|empty"; }
            try { var x = l[-1]; } catch (ArgumentOutOfRangeException) { s += "|range"; }
            try { double[] a = new double[2]; a[2] = 1; } catch (IndexOutOfRangeException) { s += "|index"; }
            try { List<double> n = null; n.ToList(); } catch (ArgumentNullException) { s += "|argnull"; }
            try { List<double> n = null; n.Add(1); } catch (NullReferenceException) { s += "|nullref"; }
            try { foreach (var v in l) l.Add(v); } catch (InvalidOperationException) { s += "|modified"; }
            int? q = null; s += "|" + (q + 1 == null) + (q > 0) + q.HasValue;
            var o = new[] { 5, 3, 9, 3 }.Select((v, i) => new { v, i }).OrderBy(p => p.v).Select(p => p.i);
            s += "|" + string.Join(",", o);
            return s;
        }
    }"""
    _, out = _run(src)
    assert out == "NaN|3|1|True|empty|range|index|argnull|nullref|modified|TrueFalseFalse|1,3,0,2"
    # List<T>.Sort is the Framework's introspective sort: unstable above 16 elements
    keys = [(i % 3, i) for i in range(40)]
    introsort(keys, lambda a, b: (a[0] > b[0]) - (a[0] < b[0]))
    assert [k[0] for k in keys] == sorted(k[0] for k in keys)
    assert [k[1] for k in keys if k[0] == 0] != sorted(k[1] for k in keys if k[0] == 0)
    small = [(i % 3, i) for i in range(16)]
    introsort(small, lambda a, b: (a[0] > b[0]) - (a[0] < b[0]))
    assert small == sorted(small)                     # insertion sort up to 16 elements: stable
    with pytest.raises(CsException):
        _run("public static class T { public static int F() { int z = 0; return 1 / z; } }")
