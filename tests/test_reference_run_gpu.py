"""GPU parity tests against OUTPUTS OF THE REFERENCE'S OWN C# SOURCES (tests/golden/reference_run.json, produced by
tests/golden/make_reference_run.py with the interpreter under oracle/csharp/): the CUDA path, driven through the
host mirrors of the reference's classes and the C ABI, must return what the reference's classes returned.

Bit-exact for the tableau, cutting-plane, B&B and sensitivity paths (every tableau element compared as uint64);
the revised path is held to the north star's 1e-9 with an identical pivot sequence.
The oracle is not involved here at all: golden file on one side, liblprb200.so on the other.
"""
import hashlib
import json
import os

import numpy as np
import pytest

import lpr_381_group_v22_b200 as L
from lpr_381_group_v22_b200 import _native as N

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "reference_run.json")))


def unmat(m):
    return np.array([float.fromhex(h) for h in m["hex"]], dtype=np.float64).reshape(m["shape"])


def unhex(v):
    return np.array([float.fromhex(h) for h in v], dtype=np.float64)


def assert_bits(a, b, what):
    a = np.ascontiguousarray(a, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    if not np.array_equal(a.view(np.uint64), b.view(np.uint64)):
        bad = np.argwhere(a.view(np.uint64) != b.view(np.uint64))
        i = tuple(bad[0])
        raise AssertionError(f"{what}: {len(bad)} elements differ, first at {i}: {a[i]!r} vs {b[i]!r}")


def constraints_of(g):
    return [L.Constraint(list(co), rel, float(rhs)) for co, rel, rhs in g["constraints"]]


def sha(text):
    return hashlib.sha256(text.encode("utf-8")).hexdigest()


@pytest.mark.parametrize("i", range(len(GOLD["primal"])))
def test_primal_simplex_solver(i):
    g = GOLD["primal"][i]
    trace = "snapshots" in g
    s = L.PrimalSimplexSolver(g["objective"], constraints_of(g), g["is_max"], trace=trace)
    assert_bits(s.DeviceTableau.read(), unmat(g["initial_tableau"]), "constructor")
    s.Solve()
    assert [list(p) for p in s.PivotLog] == g["pivots"]
    assert s.Status == (N.OPTIMAL if g["status"] == "optimal" else N.UNBOUNDED)
    assert_bits(s.FinalTableau, unmat(g["final_tableau"]), "FinalTableau")
    assert_bits(s.GetFinalTableau(), unmat(g["get_final_tableau"]), "GetFinalTableau()")
    assert s.BasicVariables == g["basis"]
    assert float(s.FinalZ).hex() == g["final_z"]
    if g["status"] == "optimal":
        assert_bits(s.SolutionVector, unhex(g["x"]), "SolutionVector")
    else:
        assert s.SolutionVector is None and g["x"] is None
    assert sha(s.FinalTable) == g["final_table_sha256"]
    if trace:
        assert s.IterationSnapshots == g["snapshots"]


@pytest.mark.parametrize("i", range(len(GOLD["primal2"])))
def test_primal_simplex_solver2(i):
    g = GOLD["primal2"][i]
    T = unmat(g["tableau"])
    s = L.PrimalSimplexSolver2(T[0].copy(), [r.copy() for r in T[1:]])
    ok = s.Solve(g["max_iters"], g["print_steps"])
    assert ok == g["returned"]
    obj, rows = s.GetRows(False)
    assert_bits(np.vstack([obj] + list(rows)), unmat(g["final_tableau"]), "GetRows(false)")
    assert [list(p) for p in s.PivotLog] == g["pivots"]
    if ok:
        assert float(s.FinalZ).hex() == g["final_z"]


@pytest.mark.parametrize("i", range(len(GOLD["dual"])))
def test_dual_simplex_solver(i):
    g = GOLD["dual"][i]
    T = unmat(g["tableau"])
    obj, rows = T[0].copy(), [r.copy() for r in T[1:]]
    d = L.DualSimplexSolver()
    if g["exception"] is not None:
        with pytest.raises(L.InvalidOperationException):
            d.Solve(obj, rows, g["max_iters"], g["print_steps"])
    else:
        assert d.Solve(obj, rows, g["max_iters"], g["print_steps"]) == g["returned"]
    assert_bits(np.vstack([obj] + rows), unmat(g["final_tableau"]), "rows after Solve")
    if g["print_steps"]:
        assert [list(p) for p in d.PivotLog] == g["printed_pivots"]


CUT_END = {"Displayed the Optimal Tableau.": (N.OPTIMAL,), "All RHS are integers. No Gomory cut needed.": (N.NO_CUT_NEEDED,),
           "No valid pivot column on the cut (need a negative cut coeff with non-zero obj coeff).": (N.NO_PIVOT_COL,),
           "Dual Simplex failed (infeasible or max iters).": (N.INFEASIBLE, N.ITER_LIMIT),
           "Pivot too small/zero.": (N.PIVOT_TOO_SMALL,),
           "Cutting-plane step finished (further steps may be required).": (N.CUT_STEP_DONE,)}


@pytest.mark.parametrize("i", range(len(GOLD["cutting_plane"])))
def test_cutting_plane_solver(i):
    g = GOLD["cutting_plane"][i]
    T = unmat(g["tableau"])
    obj, rows = T[0].copy(), [r.copy() for r in T[1:]]
    cp = L.CuttingPlaneSolver()
    cp.CuttingPlaneSolution(obj, rows)
    assert_bits(np.vstack([obj] + rows), unmat(g["final_tableau"]), "rows after CuttingPlaneSolution")
    assert cp.Status in CUT_END[g["last_console_line"]]
    k = len(g["cut_pivots_1based"])
    assert [[T.shape[0] + j, int(c) + 1] for j, c in enumerate(cp.CutLog[:k, 1])] == g["cut_pivots_1based"]


@pytest.mark.parametrize("i", range(len(GOLD["revised"])))
def test_revised_primal_simplex_solver(i):
    g = GOLD["revised"][i]
    n = len(g["c"])
    cons = [L.Constraint(list(a), rel, float(b)) for a, rel, b in zip(g["A"], g["relations"], g["b"])]
    s = L.RevisedPrimalSimplexSolver(g["c"], cons, g["is_min"], trace=True)
    try:
        msg = None
        try:
            s.Solve()
        except Exception as e:        # the reference throws System.Exception with these messages (:91, :179, :267)
            msg = str(e)
        assert msg == g["exception"]
        labels = [f"x{e + 1}" if e < n else f"S{e - n + 1}" for _, e, _ in s.PivotLog]
        assert labels == g["entering_labels"]
        assert s.BasicVariables == g["basis"]
        binv = unmat(g["binv"])
        assert np.allclose(s.BInverse, binv, rtol=1e-9, atol=1e-9 * max(1.0, float(np.abs(binv).max())))
        if msg is None:
            assert len(s.IterationSnapshots) == g["n_snapshots"]
            z, x, xb = float.fromhex(g["final_z"]), unhex(g["x"]), unhex(g["xb"])
            assert abs(s.FinalZ - z) <= 1e-9 * max(1.0, abs(z))
            assert np.allclose(s.SolutionVector, x, rtol=1e-9, atol=1e-9 * max(1.0, float(np.abs(x).max())))
            assert np.allclose(s.BasicValues, xb, rtol=1e-9, atol=1e-9 * max(1.0, float(np.abs(xb).max())))
    finally:
        s.close()


@pytest.mark.parametrize("i", range(len(GOLD["bb"])))
def test_branch_and_bound_adapter(i):
    g = GOLD["bb"][i]
    p = L.PrimalSimplexSolver(g["objective"], constraints_of(g), True, trace=False)
    p.Solve()
    assert_bits(p.FinalTableau, unmat(g["root_tableau"]), "root tableau")
    x, z = L.BranchAndBoundAdapter.SolveFromPrimal(p, g["enable_pruning"], g["is_min"])
    r = L.BranchAndBoundAdapter.LastRun
    nodes = g["nodes"]
    assert r["nodes"] == len(nodes)
    log = r["node_log"]
    assert log[:, 0].tolist() == [nd["depth"] for nd in nodes]
    assert log[:, 3].astype(bool).tolist() == [nd["pruned"] for nd in nodes]
    for k, nd in enumerate(nodes):
        if not nd["pruned"]:
            assert int(log[k, 1]) == nd["branch_var"], (k, nd)
            assert bool(log[k, 2]) == nd["integer"], (k, nd)
    assert_bits(np.array(x, dtype=np.float64), unhex(g["x"]), "x")
    assert float(z).hex() == g["z"]
    assert (r["status"] == N.NODE_LIMIT) == g["hit_node_cap"]


@pytest.mark.parametrize("i", range(len(GOLD["bb_formulate"])))
def test_dual_simplex_solver_bb_formulate_and_solve(i):
    g = GOLD["bb_formulate"][i]
    solver = L.BranchBoundSimplexSolver.DualSimplexSolverBB()
    rows = [list(r) for r in g["rows"]]
    T = solver.FormulateTableau(list(g["objective"]), rows)
    assert_bits(T, unmat(g["tableau"]), "FormulateTableau")
    assert [[float(v).hex() for v in r] for r in rows] == g["rows_after"]      # the caller's rows are mutated (:42-56)
    Td, opt, prow, pcol = solver.DoDualSimplex(list(g["objective"]), [list(r) for r in g["rows"]], g["is_min"])
    if g["optimal_value"] is None:
        assert opt is None
    else:
        assert_bits(Td, unmat(g["final_tableau"]), "DoDualSimplex")
        assert float(opt).hex() == g["optimal_value"]
        assert prow == g["pivot_rows"] and pcol == g["pivot_cols"]


@pytest.mark.parametrize("i", range(len(GOLD["sensitivity"])))
def test_sensitivity_add_constraint(i):
    g = GOLD["sensitivity"][i]
    s = L.PrimalSimplexSolver(g["objective"], constraints_of(g), True, trace=False)
    s.Solve()
    assert_bits(s.FinalTableau, unmat(g["final_tableau"]), "final tableau")
    with L.SensitivityAnalyzer(s.FinalTableau, s.SolutionVector, s.FinalZ, s.BasicVariables) as sa:
        assert sa.BasicVariables == g["basis_rebuilt"]
        if g["exception"] is None:
            sa.AddNewConstraintNonInteractive(g["tech"], g["rhs"])
            assert float(sa.CurrentZ).hex() == g["z_after"]
            assert_bits(np.array(sa.solutionVector), unhex(g["x_after"]), "solution")
        else:
            with pytest.raises(L.InvalidOperationException) as err:
                sa.AddNewConstraintNonInteractive(g["tech"], g["rhs"])
            assert str(err.value) == g["exception"]
        assert_bits(sa.CurrentTableau, unmat(g["tableau_after"]), "tableau after")
        assert sa.BasicVariables == g["basis_after"]


@pytest.mark.parametrize("i", range(len(GOLD["sensitivity_rhs"])))
def test_sensitivity_resolve_after_rhs_change(i):
    g = GOLD["sensitivity_rhs"][i]
    T = unmat(g["tableau_before_resolve"])
    with L.DeviceTableau.from_host(T) as t:
        t.sens_rebuild_basis()
        r = t.solve(L.RULE_SENS, max_pivots=10000)
        assert (r["status"] == N.OPTIMAL) == (g["exception"] is None)
        assert_bits(t.read(), unmat(g["tableau_after"]), "tableau after")
        assert t.basis.tolist() == g["basis_after"]
        if g["exception"] is None:
            assert float(t.objective()).hex() == g["z_after"]
            assert_bits(t.sens_solution(), unhex(g["x_after"]), "solution")


@pytest.mark.parametrize("i", range(len(GOLD["mid_size"])))
def test_mid_size_runs(i):
    """12-15 pivots on 19x43 .. 41x69 tableaux: digests of what the reference's classes computed"""
    g = GOLD["mid_size"][i]
    n = len(g["objective"])
    pg, rg = g["primal"], g["revised"]
    s = L.PrimalSimplexSolver(g["objective"], constraints_of(g), True, trace=False)
    assert hashlib.sha256(s.DeviceTableau.read().tobytes()).hexdigest() == pg["initial_tableau_sha256"]
    s.Solve()
    assert [list(p) for p in s.PivotLog] == pg["pivots"] and s.BasicVariables == pg["basis"]
    assert list(s.FinalTableau.shape) == pg["final_tableau_shape"]
    assert hashlib.sha256(np.ascontiguousarray(s.FinalTableau).tobytes()).hexdigest() == pg["final_tableau_sha256"]
    assert float(s.FinalZ).hex() == pg["final_z"]
    assert_bits(s.SolutionVector, unhex(pg["x"]), "SolutionVector")
    rv = L.RevisedPrimalSimplexSolver(g["objective"], constraints_of(g), False, trace=False)
    try:
        rv.Solve()
        labels = [f"x{e + 1}" if e < n else f"S{e - n + 1}" for _, e, _ in rv.PivotLog]
        assert labels == rg["entering_labels"] and rv.BasicVariables == rg["basis"]
        z, x = float.fromhex(rg["final_z"]), unhex(rg["x"])
        assert abs(rv.FinalZ - z) <= 1e-9 * abs(z)
        assert np.allclose(rv.SolutionVector, x, rtol=1e-9, atol=1e-9 * float(np.abs(x).max()))
    finally:
        rv.close()
