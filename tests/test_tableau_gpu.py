"""GPU parity tests for the dense tableau path (PrimalSimplexSolver / PrimalSimplexSolver2 /
DualSimplexSolver / cutting plane) through the C ABI, against the CPU oracle.  Bit-exact."""
import numpy as np
import pytest

import oracle_lib as O
import lpr_381_group_v22_b200 as L
from lpr_381_group_v22_b200 import _native as N

pytestmark = pytest.mark.gpu


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


def assert_bit_equal(a, b, what="tableau"):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    if not np.array_equal(bits(a), bits(b)):
        bad = np.argwhere(bits(a) != bits(b))
        i = tuple(bad[0])
        raise AssertionError(f"{what}: {len(bad)} elements differ, first at {i}: {a[i]!r} vs {b[i]!r}")


def cli_model(obj, cons):
    cons = [L.Constraint(*c) for c in cons]
    L.add_cli_bound_rows(len(obj), cons)
    return obj, cons


def oracle_cons(cons):
    return [(c.Coefficients, c.Relation, c.RHS) for c in cons]


MODEL_A = ([2, 3, 3, 5, 2, 4], [([11, 8, 6, 14, 10, 10], "<=", 40)])
MODEL_B = ([2, 3, 4], [([1, 2, 3], "<=", 10), ([3, 2, 1], ">=", 15)])


@pytest.mark.parametrize("model", [MODEL_A, MODEL_B])
@pytest.mark.parametrize("cli", [True, False])
@pytest.mark.parametrize("trace", [True, False])
def test_fixture_models_primal(model, cli, trace):
    obj, cons = model
    cons = [L.Constraint(*c) for c in cons]
    if cli:
        L.add_cli_bound_rows(len(obj), cons)
    s = L.PrimalSimplexSolver(obj, cons, trace=trace)
    s.Solve()
    T0, b0 = O.primal_build(obj, oracle_cons(cons))
    ref = O.primal_solve(T0, b0)
    assert s.Status == ref["status"]
    assert s.PivotLog == [tuple(x) for x in ref["log"].tolist()]
    assert s.BasicVariables == ref["basis"].tolist()
    assert_bit_equal(s.GetFinalTableau(), ref["T"])
    assert bits(np.array(s.SolutionVector)).tolist() == bits(O.primal_extract(ref["T"], len(obj))).tolist()
    assert s.FinalZ == ref["T"][0, -1]
    if trace:
        assert len(s.IterationSnapshots) == ref["n_pivots"] + 2


def test_appendix_c_known_answers():
    obj, cons = cli_model(*MODEL_A)
    s = L.PrimalSimplexSolver(obj, cons)
    s.Solve()
    assert s.PivotLog == [(5, 3), (7, 5), (3, 1), (4, 2), (1, 0), (1, 4)]
    assert s.BasicVariables == [4, 7, 1, 2, 3, 11, 5]
    assert s.FinalZ.hex() == "0x1.ecccccccccccdp+3"
    assert s.SolutionVector == [0.0, 1.0, 1.0, 1.0, 0.2, 1.0]
    obj, cons = cli_model(*MODEL_B)
    s = L.PrimalSimplexSolver(obj, cons)
    s.Solve()
    assert s.PivotLog == [(5, 2), (4, 1), (3, 0)]
    assert s.FinalZ == 9.0 and s.SolutionVector == [1.0, 1.0, 1.0]
    assert s.GetFinalTableau()[:, -1].tolist() == [9.0, 4.0, -9.0, 1.0, 1.0, 1.0]  # Q3: infeasible "optimum"


def test_build_matches_oracle_ctor():
    rng = np.random.default_rng(5)
    n, m = 7, 5
    obj = rng.normal(size=n).round(3).tolist()
    cons = []
    for i in range(m):
        k = [n, n + 3, n - 2, n, n + 1][i]  # ragged coefficient lists (only the first n are read)
        cons.append(L.Constraint(rng.normal(size=k).round(3).tolist(), ["<=", ">=", "=", ">=", "<="][i],
                                 float(rng.normal())))
    for is_max in (True, False):
        with L.DeviceTableau.from_model(obj, cons, is_max) as t:
            T0, b0 = O.primal_build(obj, oracle_cons(cons), is_max)
            assert_bit_equal(t.read(), T0)
            assert t.basis.tolist() == b0.tolist()


@pytest.mark.parametrize("m,n,seed", [(8, 16, 1), (33, 70, 2), (64, 128, 3), (100, 37, 4), (255, 513, 5)])
@pytest.mark.parametrize("fused", [True, False])
def test_random_dense_lp_bitexact(m, n, seed, fused):
    A, b, c = O.gen_dense_lp(seed, m, n)
    T0, b0 = O.primal_build(list(c), [(A[i], "<=", b[i]) for i in range(m)])
    ref = O.primal_solve(T0, b0)
    with L.DeviceTableau.from_host(T0) as t:
        r = t.solve(L.RULE_PRIMAL, fused=fused)
        assert r["status"] == ref["status"] == L.OPTIMAL
        assert r["n_pivots"] == ref["n_pivots"]
        assert r["log"].tolist() == ref["log"].tolist()
        assert t.basis.tolist() == ref["basis"].tolist()
        assert_bit_equal(t.read(), ref["T"])
        assert_bit_equal(t.extract_solution(n), O.primal_extract(ref["T"], n), "solution")


def test_device_generator_matches_oracle():
    m, n, seed = 40, 90, 381
    A, b, c = O.gen_dense_lp(seed, m, n)
    T0, b0 = O.primal_build(list(c), [(A[i], "<=", b[i]) for i in range(m)])
    with L.DeviceTableau.dense_lp(seed, m, n) as t:
        assert_bit_equal(t.read(), T0)
        assert t.basis.tolist() == b0.tolist()


def test_max_pivots_and_step():
    m, n, seed = 30, 60, 9
    A, b, c = O.gen_dense_lp(seed, m, n)
    T0, b0 = O.primal_build(list(c), [(A[i], "<=", b[i]) for i in range(m)])
    ref = O.primal_solve(T0, b0, max_pivots=5)
    assert ref["status"] == O.ITER_LIMIT
    with L.DeviceTableau.from_host(T0) as t:
        r = t.solve(L.RULE_PRIMAL, max_pivots=5)
        assert r["status"] == L.ITER_LIMIT and r["n_pivots"] == 5
        assert_bit_equal(t.read(), ref["T"])
    with L.DeviceTableau.from_host(T0) as t:
        for k in range(5):
            e, l, st = t.step(L.RULE_PRIMAL)
            assert (l, e) == tuple(ref["log"][k]) and st == L.RUNNING
        assert_bit_equal(t.read(), ref["T"])


def test_unbounded_and_degenerate():
    # unbounded: max x1 with x1 - x2 <= 1
    T = np.array([[-1.0, 0.0, 0.0, 0.0], [1.0, -1.0, 1.0, 1.0]])
    # second pivot: entering x2 has only a negative entry
    ref = O.primal_solve(T)
    with L.DeviceTableau.from_host(T) as t:
        r = t.solve(L.RULE_PRIMAL)
        assert r["status"] == ref["status"]
        assert r["log"].tolist() == ref["log"].tolist()
        assert_bit_equal(t.read(), ref["T"])
    s = L.PrimalSimplexSolver([1.0, 1.0], [L.Constraint([1.0, -1.0], "<=", 1.0)])
    s.Solve()
    assert s.Status == L.UNBOUNDED and s.SolutionVector is None and s.FinalZ == 0.0
    # ties: duplicated columns / rows => lowest index wins in both rules
    T = np.array([[-3.0, -3.0, -1.0, 0, 0, 0, 0.0], [1, 1, 1, 1, 0, 0, 4.0], [1, 1, 0, 0, 1, 0, 4.0],
                  [2, 2, 1, 0, 0, 1, 8.0]])
    ref = O.primal_solve(T)
    with L.DeviceTableau.from_host(T) as t:
        for fused in (True, False):
            t.upload(T)
            r = t.solve(L.RULE_PRIMAL, fused=fused)
            assert r["log"].tolist() == ref["log"].tolist()
            assert_bit_equal(t.read(), ref["T"])


def _random_tableau(rng, R, C, neg_rhs=False):
    T = rng.integers(-4, 9, size=(R, C)).astype(float)
    T[1:, C - 1 - (R - 1):C - 1] = np.eye(R - 1)
    T[0, C - 1 - (R - 1):] = 0.0
    T[1:, -1] = rng.integers(1, 20, size=R - 1)
    T[0, :C - R] = -rng.integers(1, 9, size=C - R)
    if neg_rhs:
        T[0, :C - R] = rng.integers(1, 9, size=C - R)
        T[1:, :C - R] = -np.abs(T[1:, :C - R]) - 1
        T[1:, -1] = -rng.integers(1, 20, size=R - 1)
    return T


@pytest.mark.parametrize("seed", range(6))
def test_primal2_rule(seed):
    rng = np.random.default_rng(seed)
    T = _random_tableau(rng, 6 + seed, 14 + 2 * seed)
    T[1:, :4] = np.abs(T[1:, :4]) + 1  # bounded
    for ps in (False, True):
        ref = O.primal2_solve(T, 10000, ps)
        s = L.PrimalSimplexSolver2(T[0], [r for r in T[1:]])
        ok = s.Solve(10000, ps)
        assert ok == (ref["status"] == O.OPTIMAL)
        assert s.PivotLog == [tuple(x) for x in ref["log"].tolist()]
        obj, rows = s.GetRows(False)
        assert_bit_equal(np.vstack([obj[None, :], np.array(rows)]), ref["T"])
    ref = O.primal2_solve(T, 2, True)  # iteration cap only live when printing (Q16)
    with L.DeviceTableau.from_host(T) as t:
        r = t.solve(L.RULE_PRIMAL2, max_pivots=2, print_steps=True)
        assert r["status"] == ref["status"] and r["n_pivots"] == ref["n_pivots"]
        assert_bit_equal(t.read(), ref["T"])


@pytest.mark.parametrize("seed", range(6))
def test_dual_rule(seed):
    rng = np.random.default_rng(100 + seed)
    T = _random_tableau(rng, 5 + seed, 12 + 2 * seed, neg_rhs=True)
    ref = O.dual_solve(T, 10000, True)
    obj = T[0].copy()
    rows = [r.copy() for r in T[1:]]
    d = L.DualSimplexSolver()
    ok = d.Solve(obj, rows, 10000, True)
    assert ok == (ref["status"] == O.OPTIMAL)
    assert d.PivotLog == [tuple(x) for x in ref["log"].tolist()]
    assert_bit_equal(np.vstack([obj[None, :], np.array(rows)]), ref["T"])


def test_hysteresis_fallback_paths():
    # near-ties inside the 1e-9/1e-10 windows force the literal sequential replay
    T = np.array([[-1.0, -1.0 - 5e-11, -1.0 - 1.2e-10, -0.5, 0, 0, 0.0],
                  [1, 1, 1, 1, 1, 0, 4.0], [1, 1, 1 + 5e-11, 2, 0, 1, 4.0]])
    ref = O.primal2_solve(T)
    with L.DeviceTableau.from_host(T) as t:
        r = t.solve(L.RULE_PRIMAL2)
        assert r["log"].tolist() == ref["log"].tolist()
        assert_bit_equal(t.read(), ref["T"])
    T = np.array([[1.0, 2.0, 3.0, 0, 0, 0, 0.0], [-1, -2, -1, 1, 0, 0, -2.0], [-1, -1, -3, 0, 1, 0, -2.0 - 5e-10],
                  [-2, -1, -1, 0, 0, 1, -2.0 - 1.2e-9]])
    ref = O.dual_solve(T)
    with L.DeviceTableau.from_host(T) as t:
        r = t.solve(L.RULE_DUAL, max_pivots=10000, print_steps=True)
        assert r["log"].tolist() == ref["log"].tolist()
        assert_bit_equal(t.read(), ref["T"])


def test_pivot_at_and_reads():
    rng = np.random.default_rng(3)
    T = rng.normal(size=(9, 21))
    ref = T.copy()
    O.lib().orc_primal_pivot(9, 21, ref.ctypes.data_as(O._dp), 4, 7, 1)
    with L.DeviceTableau.from_host(T) as t:
        t.pivot_at(4, 7)
        assert_bit_equal(t.read(), ref)
        assert_bit_equal(t.read_row(4), ref[4], "row")
        assert_bit_equal(t.read_col(7), ref[:, 7], "col")


def test_gomory_cut_and_cutting_plane_model_a():
    obj, cons = cli_model(*MODEL_A)
    s = L.PrimalSimplexSolver(obj, cons)
    s.Solve()
    Tf = s.GetFinalTableau()
    row, cut = O.gomory_cut(Tf)
    with L.DeviceTableau.from_host(Tf, row_cap=Tf.shape[0] + 16) as t:
        r2, c2 = t.gomory_cut()
        assert r2 == row == 0
        assert_bit_equal(c2, cut, "cut")
        ref = O.cutting_plane(Tf)
        res = t.cutting_plane()
        assert res["status"] == ref["status"] == L.OPTIMAL
        assert res["log"].tolist() == ref["log"].tolist() == [[0, 0, 1, 1]]
        assert_bit_equal(t.read(), ref["T"])
        assert t.read()[0, -1] == 15.0


@pytest.mark.parametrize("seed", range(5))
def test_cutting_plane_random_ip(seed):
    m, n = 4 + seed, 6 + seed
    A, b, c = O.gen_dense_ip(1000 + seed, m, n)
    T0, b0 = O.primal_build(list(c), [(A[i], "<=", b[i]) for i in range(m)])
    lp = O.primal_solve(T0, b0)
    ref = O.cutting_plane(lp["T"], max_cuts=12)
    with L.DeviceTableau.from_host(lp["T"], row_cap=lp["T"].shape[0] + 16) as t:
        res = t.cutting_plane(max_cuts=12)
        assert res["status"] == ref["status"]
        assert res["log"].tolist() == ref["log"].tolist()
        assert_bit_equal(t.read(), ref["T"])


def test_full_size_window_cfg2():
    """BASELINE cfg2 shape (4097 x 12289): the first pivots bit-exact against the oracle, generated
    on device, plus size-independent invariants afterwards."""
    m, n, seed, K = 4096, 8192, 383, 40  # 16 + 16 + 8: two full delayed-update groups and a partial one
    A, b, c = O.gen_dense_lp(seed, m, n)
    T0, b0 = O.primal_build(list(c), [(A[i], "<=", b[i]) for i in range(m)])
    del A
    ref = O.primal_solve(T0, b0, max_pivots=K, threads=8)
    with L.DeviceTableau.dense_lp(seed, m, n) as t:
        r = t.solve(L.RULE_PRIMAL, max_pivots=K)
        assert r["log"].tolist() == ref["log"].tolist()
        got = t.read()
        assert_bit_equal(got, ref["T"])
        # basic columns are unit vectors; objective row of basic columns is 0
        basis = t.basis
        for i in (0, 1, m // 2, m - 1):
            col = got[:, basis[i]]
            assert col[i + 1] == 1.0 and np.count_nonzero(col) == 1


def test_full_size_solve_cfg2_plans_agree_and_certify_optimality():
    """BASELINE cfg2 solved to optimality (14 866 pivots).  The oracle needs ~13 minutes for this, so the full-size
    check is (1) the per-pivot plan -- the one pinned bit-exactly to the oracle on windows -- and the default
    overlapped delayed-update plan give the same pivot log and the same final tableau, bit for bit, and (2) the
    result certifies itself: primal feasible, dual feasible, complementary, z = c.x = y.b (the LP optimality
    conditions, evaluated from the synthetic A, b, c on the host)."""
    m, n, seed = 4096, 8192, 383
    A, b, c = O.gen_dense_lp(seed, m, n)
    with L.DeviceTableau.dense_lp(seed, m, n) as t:
        r = t.solve(L.RULE_PRIMAL, log_cap=1 << 15)
        assert r["status"] == L.OPTIMAL and r["n_pivots"] == 14866
        T = t.read()
        z, x, basis = t.objective(), t.extract_solution(n), t.basis.copy()
    with L.DeviceTableau.dense_lp(seed, m, n) as t:
        r1 = t.solve(L.RULE_PRIMAL, log_cap=1 << 15, blocked=False)  # one sweep per pivot
        assert r1["status"] == L.OPTIMAL and r1["log"].tolist() == r["log"].tolist()
        assert_bit_equal(t.read(), T)
        assert t.basis.tolist() == basis.tolist()
    # optimality certificate
    assert np.all(T[0, :-1] >= 0.0) and np.all(T[1:, -1] >= 0.0)          # reduced costs / basic values
    y = T[0, n:n + m]                                                     # duals sit under the slack columns
    assert z == T[0, -1] and abs(z - c @ x) <= 1e-9 * abs(z) and abs(z - y @ b) <= 1e-9 * abs(z)
    assert np.all(A @ x <= b * (1 + 1e-9)) and np.all(x >= 0.0)
    assert np.all(A.T @ y >= c * (1 - 1e-9)) and np.all(y >= 0.0)
    slack = b - A @ x
    assert np.all(np.abs(slack * y) <= 1e-6 * np.abs(z))                  # complementary slackness
    for i in range(0, m, 257):                                            # basic columns are unit vectors
        col = T[:, basis[i]]
        assert col[i + 1] == 1.0 and np.count_nonzero(col) == 1


def test_full_size_solve_cfg2_matches_oracle_digest():
    """The WHOLE 14 866-pivot cfg2 solve against the oracle: SHA-256 of the pivot log, of the basis and of the final
    tableau, compared with the digests the threaded CPU oracle produced once (tests/golden/make_cfg2_full_hash.py,
    ~4 minutes on 8 cores; committed as tests/golden/cfg2_full_solve.json).  Default plan (overlapped delayed updates)."""
    import hashlib
    import json
    import os
    with open(os.path.join(os.path.dirname(__file__), "golden", "cfg2_full_solve.json")) as f:
        gold = json.load(f)
    m, n, seed = gold["m"], gold["n"], gold["seed"]
    with L.DeviceTableau.dense_lp(seed, m, n) as t:
        r = t.solve(L.RULE_PRIMAL, log_cap=1 << 15)
        assert r["status"] == L.OPTIMAL and r["n_pivots"] == gold["n_pivots"]
        log = np.ascontiguousarray(r["log"], dtype=np.int32)
        assert log[:8].tolist() == gold["first_pivots"] and log[-8:].tolist() == gold["last_pivots"]
        assert hashlib.sha256(log.tobytes()).hexdigest() == gold["pivot_log_sha256"]
        assert hashlib.sha256(np.ascontiguousarray(t.basis, dtype=np.int32).tobytes()).hexdigest() == gold["basis_sha256"]
        T = np.ascontiguousarray(t.read(), dtype=np.float64)
        assert float(T[0, -1]).hex() == gold["z_hex"]
        assert hashlib.sha256(T.tobytes()).hexdigest() == gold["final_tableau_sha256"]


@pytest.mark.parametrize("seed", range(5))
def test_sensitivity_resolve_rule(seed):
    """SensitivityAnalyzer.ResolveAll (dual phase then primal re-optimisation, SensitivityAnalyzer.cs:121-201):
    perturb the RHS / objective of an optimal tableau and re-solve on the device."""
    m, n = 6 + seed, 9 + seed
    A, b, c = O.gen_dense_lp(50 + seed, m, n)
    T0, b0 = O.primal_build(list(c), [(A[i], "<=", b[i]) for i in range(m)])
    opt = O.primal_solve(T0, b0)
    T = opt["T"].copy()
    rng = np.random.default_rng(seed)
    T[1 + rng.integers(0, m), -1] -= 5.0 * (1 + seed)        # RHS change => dual simplex needed
    nb = [j for j in range(n + m) if j not in opt["basis"].tolist()]
    T[0, nb[seed % len(nb)]] = -0.75                           # a reduced cost turns negative => primal phase
    ref = O.sens_resolve(T, opt["basis"])
    with L.DeviceTableau.from_host(T) as t:
        t.basis = opt["basis"]
        r = t.solve(L.RULE_SENS, max_pivots=10000)
        assert r["status"] == ref["status"]
        assert r["log"].tolist() == ref["log"].tolist()
        assert t.basis.tolist() == ref["basis"].tolist()
        assert_bit_equal(t.read(), ref["T"])


@pytest.mark.parametrize("m,n,seed", [(5, 7, 1), (9, 6, 2), (12, 20, 3), (40, 90, 4), (130, 70, 5)])
def test_sensitivity_add_constraint_on_device(m, n, seed):
    """SensitivityAnalyzer(finalTableau, ...).AddNewConstraintNonInteractive(tech, rhs) with the tableau kept in
    HBM (SensitivityAnalyzer.cs:22-41, :609-659, :706-723, ResolveAll :203-209) == the oracle, bit for bit,
    twice in a row (the second constraint is added to the re-solved tableau)."""
    A, b, c = O.gen_dense_lp(100 + seed, m, n)
    s = L.PrimalSimplexSolver(list(c), [L.Constraint(A[i], "<=", b[i]) for i in range(m)], trace=False)
    s.Solve()
    T = s.GetFinalTableau()
    with L.SensitivityAnalyzer(T, s.SolutionVector, s.FinalZ, s.BasicVariables) as sa:
        Tref = np.array(T, dtype=np.float64)
        Tref[0, -1] = s.FinalZ
        bref = O.sens_rebuild_basis(Tref)
        assert sa.BasicVariables == bref.tolist()
        sol = list(s.SolutionVector)
        for k in range(2):
            cols = Tref.shape[1]
            tech = np.zeros(cols - 1)
            tech[:n] = 1.0 + (np.arange(n) + k) % 3
            xfull = O.sens_solution(Tref)
            rhs = (0.6 - 0.2 * k) * float(tech @ xfull)
            ax = 0.0
            for j in range(min(len(tech), len(sol))):
                ax += tech[j] * sol[j]
            T1, _ = O.sens_add_constraint(Tref, bref, tech, rhs - ax)
            b1 = O.sens_rebuild_basis(T1)
            ref = O.sens_resolve(T1, b1)
            if ref["status"] != O.OPTIMAL:   # the reference throws (:151 / :194): so does the mirror
                with pytest.raises(L.InvalidOperationException):
                    sa.AddNewConstraintNonInteractive(tech, rhs)
                assert sa.LastPivotLog == [tuple(p) for p in ref["log"].tolist()]
                assert_bit_equal(sa.CurrentTableau, ref["T"])
                break
            r = sa.AddNewConstraintNonInteractive(tech, rhs)
            assert r["status"] == ref["status"] == O.OPTIMAL
            assert r["log"].tolist() == ref["log"].tolist()
            assert sa.BasicVariables == ref["basis"].tolist()
            assert_bit_equal(sa.CurrentTableau, ref["T"])
            assert sa.CurrentZ == ref["T"][0, -1]
            assert_bit_equal(np.array(sa.solutionVector), O.sens_solution(ref["T"]))
            mm = ref["T"].shape[0] - 1
            assert sa.ShadowPrices() == ref["T"][0, ref["T"].shape[1] - 1 - mm:-1].tolist()
            Tref, bref, sol = ref["T"], O.sens_rebuild_basis(ref["T"]), O.sens_solution(ref["T"]).tolist()
        with pytest.raises(L.LprError):   # headroom exhausted after 16 additions is an error, not a crash
            for _ in range(20):
                sa._tab.sens_add_constraint(np.zeros(sa._tab.shape[1] - 1), 1.0)


def test_edge_shapes_and_errors():
    # smallest legal tableau, single constraint / single variable
    T = np.array([[-1.0, 0.0, 0.0], [2.0, 1.0, 4.0]])
    ref = O.primal_solve(T)
    with L.DeviceTableau.from_host(T) as t:
        r = t.solve(L.RULE_PRIMAL)
        assert r["log"].tolist() == ref["log"].tolist() == [[1, 0]]
        assert_bit_equal(t.read(), ref["T"])
        assert t.objective() == 2.0
    # already optimal: zero pivots
    T = np.array([[1.0, 2.0, 0.0, 0.0], [1.0, 1.0, 1.0, 3.0]])
    with L.DeviceTableau.from_host(T) as t:
        r = t.solve(L.RULE_PRIMAL)
        assert r["status"] == L.OPTIMAL and r["n_pivots"] == 0
        assert_bit_equal(t.read(), T)
    # ragged width that is not a multiple of anything, rows > 256*16 are covered by the cfg2 test
    rng = np.random.default_rng(1)
    for (R, Cc) in ((2, 2), (3, 17), (5, 257), (40, 513)):
        T = rng.normal(size=(R, Cc))
        ref = T.copy()
        O.lib().orc_primal_pivot(R, Cc, ref.ctypes.data_as(O._dp), R - 1, Cc // 2, 1)
        with L.DeviceTableau.from_host(T) as t:
            t.pivot_at(R - 1, Cc // 2)
            assert_bit_equal(t.read(), ref)
    with pytest.raises(L.LprError):
        L.DeviceTableau.from_host(np.zeros((1, 1)))
    with pytest.raises(L.LprError):
        with L.DeviceTableau.from_host(np.zeros((2, 3))) as t:
            t.pivot_at(5, 0)
    with pytest.raises(L.LprError):
        with L.DeviceTableau.from_host(np.zeros((2, 3))) as t:
            t.append_row(np.zeros(3))  # no headroom


def test_nan_and_inf_do_not_hang():
    T = np.array([[-1.0, -2.0, 0.0, 0.0], [np.nan, 1.0, 1.0, 4.0], [1.0, np.inf, 0.0, 2.0]])
    ref = O.primal_solve(T, max_pivots=10)
    with L.DeviceTableau.from_host(T) as t:
        r = t.solve(L.RULE_PRIMAL, max_pivots=10)
        assert r["status"] == ref["status"] and r["log"].tolist() == ref["log"].tolist()
        got = t.read()
        assert np.array_equal(np.isnan(got), np.isnan(ref["T"]))
        mask = ~np.isnan(got)
        assert np.array_equal(got[mask], ref["T"][mask])


@pytest.mark.parametrize("m,n,seed", [(8, 16, 1), (33, 70, 2), (100, 37, 4), (255, 513, 5), (300, 300, 6),
                                      (21, 2500, 7), (700, 90, 8), (50, 850, 9)])
@pytest.mark.parametrize("block", [2, 3, 8, 16])
def test_blocked_delayed_update_is_bit_identical(m, n, seed, block, monkeypatch):
    """K pending pivots applied by one sweep == one sweep per pivot == the oracle, for the overlapped
    (tableau_pipelined.cu: select of group g+1 runs beside the out-of-place sweep of group g) and the
    single-stream (tableau_blocked.cu) plans."""
    import subprocess, sys, os, json
    # LPR_TAB_BLOCK is read once per process: run the comparison in a child process
    code = f"""
import sys, json, numpy as np
sys.path.insert(0, {os.path.dirname(os.path.abspath(__file__))!r}); sys.path.insert(0, {os.path.dirname(os.path.dirname(os.path.abspath(__file__)))!r})
import oracle_lib as O, lpr_381_group_v22_b200 as L
A, b, c = O.gen_dense_lp({seed}, {m}, {n})
T0, b0 = O.primal_build(list(c), [(A[i], "<=", b[i]) for i in range({m})])
out = {{}}
for cap in (-1, 7, 17, 40):
  ref = O.primal_solve(T0, b0, max_pivots=cap)
  for pipelined in (True, False):
    with L.DeviceTableau.from_host(T0) as t:
        r = t.solve(L.RULE_PRIMAL, max_pivots=cap, pipelined=pipelined)
        ok = (r["status"] == ref["status"] and r["n_pivots"] == ref["n_pivots"] and r["log"].tolist() == ref["log"].tolist()
              and t.basis.tolist() == ref["basis"].tolist()
              and np.array_equal(t.read().view(np.uint64), ref["T"].view(np.uint64)))
        # a second solve on the same handle continues from the (possibly swapped) current buffer
        if ok and cap >= 0 and r["status"] == L.ITER_LIMIT:
            ref2 = O.primal_solve(ref["T"], ref["basis"])
            r2 = t.solve(L.RULE_PRIMAL, pipelined=pipelined)
            ok = (r2["status"] == ref2["status"] and r2["log"].tolist() == ref2["log"].tolist()
                  and np.array_equal(t.read().view(np.uint64), ref2["T"].view(np.uint64)))
        out["%d:%d" % (cap, int(pipelined))] = bool(ok)
print(json.dumps(out))
"""
    env = dict(os.environ, LPR_TAB_BLOCK=str(block))
    if block == 3:   # also cover the fallback without the green-context SM partition (two priority streams)
        env["LPR_PIPE_GREEN"] = "0"
    res = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr[-2000:]
    out = json.loads(res.stdout.strip().splitlines()[-1])
    assert out and all(out.values()) and len(out) == 8, out


def test_build_paths_full_rows_and_cache_reuse():
    """the fast constructor path (coefficient block DMA'd straight into the tableau) and the buffer cache"""
    rng = np.random.default_rng(11)
    n, m = 40, 30
    obj = rng.normal(size=n).round(3).tolist()
    for trial in range(3):  # repeated create/destroy of same-size tableaux reuses cached device buffers
        cons = [L.Constraint(rng.normal(size=n).round(3).tolist(), ["<=", ">=", "="][i % 3], float(rng.normal()))
                for i in range(m)]
        for is_max in (True, False):
            with L.DeviceTableau.from_model(obj, cons, is_max) as t:
                T0, b0 = O.primal_build(obj, oracle_cons(cons), is_max)
                assert_bit_equal(t.read(), T0)
                assert t.basis.tolist() == b0.tolist()
    big = rng.normal(size=(600, 2200))  # > 8 MB: goes through the cache on destroy
    for _ in range(3):
        with L.DeviceTableau.from_host(big) as t:
            assert_bit_equal(t.read(), big)


# ---- SURVEY 8f rows 2-3 on the device: snapshot text and model input ---------------------------------------------
def test_device_snapshot_equals_host_formatter():
    """lpr_tab_format (row blocks streamed D2H + native formatting) == lpr_fmt_table on the downloaded tableau ==
    the independent restatement, for a single-block and a multi-block (> 8 MB) tableau"""
    import net_reference as R
    from lpr_381_group_v22_b200.utilities import TableIterationFormater as F
    rng = np.random.default_rng(21)
    for rows, cols, nv in ((7, 12, 4), (1300, 1111, 600)):
        T = rng.normal(size=(rows, cols)) * 10.0 ** rng.integers(-4, 5, (rows, cols))
        labels = [f"x{i}" for i in range(1, 6)]
        with L.DeviceTableau.from_host(T) as t:
            dev = F.FormatDevice(t, nv, "Iteration 2 - After pivot", labels)
        assert dev == F.Format(T, nv, "Iteration 2 - After pivot", labels)
        if rows < 100:
            assert dev == R.format_table(T.tolist(), nv, "Iteration 2 - After pivot", labels)


def test_traced_solve_snapshots_are_the_reference_text():
    """IterationSnapshots of a traced solve: initial tableau, one per pivot, final block -- each the formatter's
    text of the oracle's tableau at that point"""
    import net_reference as R
    obj = [3.0, 2.0, 4.0]
    rows = [([1.0, 1.0, 2.0], "<=", 4.0), ([2.0, 0.0, 3.0], "<=", 5.0), ([2.0, 1.0, 3.0], "<=", 7.0)]
    s = L.PrimalSimplexSolver(obj, [L.Constraint(*r) for r in rows], trace=True)
    s.Solve()
    T0, b0 = O.primal_build(obj, rows)
    assert s.IterationSnapshots[0] == R.format_table(T0.tolist(), 3, "Initial Tableau")
    for k in range(1, len(s.PivotLog) + 1):
        ref = O.primal_solve(T0.copy(), b0.copy(), max_pivots=k)
        assert s.IterationSnapshots[k] == R.format_table(ref["T"].tolist(), 3, f"Iteration {k} - After pivot")


def test_model_to_device_equals_list_constructor(tmp_path):
    """lpr_tab_create_from_model (parsed text, CLI rows, binary round trip) builds the tableau the List<Constraint>
    constructor builds"""
    from lpr_381_group_v22_b200.io import Model
    text = "max +2 +3 +3 +5 +2 +4\n+11 +8 +6 +14 +10 +10 <= 40\n1 0 2 0 1 0 >= 3\nbin bin bin bin bin bin"
    m = Model.parse_text(text).add_cli_bound_rows()
    cons = [L.Constraint([11, 8, 6, 14, 10, 10], "<=", 40), L.Constraint([1, 0, 2, 0, 1, 0], ">=", 3)]
    L.add_cli_bound_rows(6, cons)
    with L.DeviceTableau.from_model([2, 3, 3, 5, 2, 4], cons) as a, m.to_device() as b:
        assert np.array_equal(a.read().view(np.uint64), b.read().view(np.uint64)) and a.basis.tolist() == b.basis.tolist()
    path = str(tmp_path / "m.lprm")
    m.save_binary(path)
    with Model.load_binary(path).to_device(is_maximization=False) as c, \
            L.DeviceTableau.from_model([2, 3, 3, 5, 2, 4], cons, False) as d:
        assert np.array_equal(c.read().view(np.uint64), d.read().view(np.uint64))
    A, bb, cc = O.gen_dense_lp(5, 40, 70)
    with Model.from_dense(cc, A, bb).to_device() as e, \
            L.DeviceTableau.from_model(list(cc), [L.Constraint(A[i], "<=", bb[i]) for i in range(40)]) as f:
        assert np.array_equal(e.read().view(np.uint64), f.read().view(np.uint64))


@pytest.mark.parametrize("persist", ["0", "1", "1-no-row-prefetch"])
def test_generic_rules_persistent_and_two_kernel_paths_agree_with_oracle(persist, monkeypatch):
    """PrimalSimplexSolver2 / DualSimplexSolver / CuttingPlaneSolver loops run by default as ONE cooperative launch
    (tableau_persistent.cu: redundant selection per CTA, out-of-place update, one grid barrier per pivot);
    LPR_TAB_PERSIST=0 keeps the two-kernel path.  Both must match the oracle bit for bit on mid-size tableaux."""
    monkeypatch.setenv("LPR_TAB_PERSIST", persist[0])
    if persist.endswith("no-row-prefetch"):  # the path of tableaux too wide for the shared-memory row prefetch
        monkeypatch.setenv("LPR_PERSIST_XROWS", "0")
    rng = np.random.default_rng(77)
    for R, C in ((40, 100), (150, 333), (257, 700)):
        T = _random_tableau(rng, R, C)
        T[1:, :4] = np.abs(T[1:, :4]) + 1
        for ps in (False, True):
            ref = O.primal2_solve(T, 10000, ps)
            with L.DeviceTableau.from_host(T) as t:
                r = t.solve(L.RULE_PRIMAL2, max_pivots=10000, print_steps=ps)
                assert r["status"] == ref["status"] and r["log"].tolist() == ref["log"].tolist()
                assert_bit_equal(t.read(), ref["T"])
        ref = O.primal2_solve(T, 3, True)
        with L.DeviceTableau.from_host(T) as t:
            r = t.solve(L.RULE_PRIMAL2, max_pivots=3, print_steps=True)
            assert r["status"] == ref["status"] == O.ITER_LIMIT and r["n_pivots"] == ref["n_pivots"]
            assert_bit_equal(t.read(), ref["T"])
        T = _random_tableau(rng, R, C, neg_rhs=True)
        ref = O.dual_solve(T, 10000, True)
        with L.DeviceTableau.from_host(T) as t:
            r = t.solve(L.RULE_DUAL, max_pivots=10000, print_steps=True)
            assert r["status"] == ref["status"] and r["log"].tolist() == ref["log"].tolist()
            assert_bit_equal(t.read(), ref["T"])
    for seed in range(3):
        m, n = 20 + 7 * seed, 30 + 9 * seed
        A, b, c = O.gen_dense_ip(2000 + seed, m, n)
        T0, b0 = O.primal_build(list(c), [(A[i], "<=", b[i]) for i in range(m)])
        lp = O.primal_solve(T0, b0)
        ref = O.cutting_plane(lp["T"], max_cuts=10)
        with L.DeviceTableau.from_host(lp["T"], row_cap=lp["T"].shape[0] + 12) as t:
            res = t.cutting_plane(max_cuts=10)
            assert res["status"] == ref["status"] and res["log"].tolist() == ref["log"].tolist()
            assert_bit_equal(t.read(), ref["T"])


def test_cutting_plane_cfg5_size_matches_oracle():
    """BASELINE cfg5's root (513 x 1537 relaxation tableau): 32 Gomory cuts with their dual / primal clean-up pivots
    (what bench.py reports as bb.root_cuts) against orc_cutting_plane: cut log and final tableau bit for bit."""
    m, n, seed = 512, 1024, 385
    A, b, c = O.gen_dense_ip(seed, m, n)
    T0, b0 = O.primal_build(list(c), [(A[i], "<=", b[i]) for i in range(m)])
    lp = O.primal_solve(T0, b0, threads=8)
    assert lp["status"] == O.OPTIMAL
    ref = O.cutting_plane(lp["T"], max_cuts=32)
    with L.DeviceTableau.from_host(lp["T"], row_cap=lp["T"].shape[0] + 40) as t:
        res = t.cutting_plane(max_cuts=32)
        assert res["n_cuts"] == ref["n_cuts"] == 32 and res["status"] == ref["status"]
        assert res["log"].tolist() == ref["log"].tolist()
        assert_bit_equal(t.read(), ref["T"])
