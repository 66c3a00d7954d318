"""CPU tests (no GPU): the oracle against (1) the known answers of SURVEY.md Appendix C, derived
independently from the cited C# loops, (2) the committed golden vectors, (3) scipy HiGHS, (4) the
reference's own in-code check knapsack B&B == DP (Program.cs:467-470)."""
import json
import os

import numpy as np
import pytest

import oracle_lib as O

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "oracle_vectors.json")))


def hexs(a):
    return [float(x).hex() for x in np.asarray(a, dtype=float).ravel()]


def cli_rows(n):
    rows = []
    for i in range(n):
        co = [0.0] * (n + 3)
        co[i] = 1.0
        co[n + 1] = 1.0
        rows.append((co, "<=", 1.0))
    return rows


MODEL_A = ([2, 3, 3, 5, 2, 4], [([11, 8, 6, 14, 10, 10], "<=", 40)])
MODEL_B = ([2, 3, 4], [([1, 2, 3], "<=", 10), ([3, 2, 1], ">=", 15)])


def final_a():
    T, b = O.primal_build(MODEL_A[0], MODEL_A[1] + cli_rows(6))
    return O.primal_solve(T, b)


def test_appendix_c1_readme_model():
    T, b = O.primal_build(MODEL_B[0], MODEL_B[1] + cli_rows(3))
    assert T.shape == (6, 9)
    r = O.primal_solve(T, b)
    assert r["log"].tolist() == [[5, 2], [4, 1], [3, 0]] and r["basis"].tolist() == [3, 4, 0, 1, 2]
    assert float(r["T"][0, -1]).hex() == "0x1.2000000000000p+3"
    assert O.primal_extract(r["T"], 3).tolist() == [1, 1, 1]
    assert r["T"][:, -1].tolist() == [9, 4, -9, 1, 1, 1]  # Q3: the >= row stays violated
    T, b = O.primal_build(*MODEL_B)
    r = O.primal_solve(T, b)
    assert r["log"].tolist() == [[1, 2], [1, 0]] and r["basis"].tolist() == [0, 4]
    assert r["T"][0, -1] == 20.000000000000004 and O.primal_extract(r["T"], 3)[0] == 10.000000000000002
    rr = O.rev_solve(np.array([[1, 2, 3], [3, 2, 1.0]]), [10, 15], [2, 3, 4])
    assert rr["log"][:, :2].tolist() == [[0, 2], [1, 0]] and rr["basis"].tolist() == [2, 0]
    assert rr["z"] == 16.25 and rr["x"].tolist() == [4.375, 0, 1.875]


def test_appendix_c2_textfile_model():
    r = final_a()
    assert r["T"].shape == (8, 14)
    assert r["log"].tolist() == [[5, 3], [7, 5], [3, 1], [4, 2], [1, 0], [1, 4]]
    assert r["basis"].tolist() == [4, 7, 1, 2, 3, 11, 5]
    assert float(r["T"][0, -1]).hex() == "0x1.ecccccccccccdp+3"
    assert O.primal_extract(r["T"], 6).tolist() == [0, 1, 1, 1, 0.2, 1]
    assert r["T"][0].tolist() == [0.2000000000000001, 0, 0, 0, 0, 0, 0.2, 0, 1.4, 1.8, 2.2, 0, 1.9999999999999998, 15.4]
    assert r["T"][:, -1].tolist() == [15.4, 0.2, 1, 1, 1, 1, 0.8, 1]
    T, b = O.primal_build(*MODEL_A)
    r2 = O.primal_solve(T, b)
    assert r2["log"].tolist() == [[1, 3], [1, 2]] and r2["T"][0, -1] == 20 and O.primal_extract(r2["T"], 6)[2] == 6.666666666666667
    A = np.vstack([[11, 8, 6, 14, 10, 10], np.eye(6)])
    rr = O.rev_solve(A, [40] + [1] * 6, MODEL_A[0])
    assert rr["log"].tolist() == [[4, 3, 10], [6, 5, 12], [2, 1, 8], [3, 2, 9], [0, 0, 6], [0, 4, 0]]
    assert rr["basis"].tolist() == [4, 7, 1, 2, 3, 11, 5] and rr["z"] == 15.399999999999999


def test_appendix_c3_branch_and_bound():
    Tf = final_a()["T"]
    r = O.bb_solve(Tf, 6, prune=False, max_nodes=20)
    assert r["nodes"] == 20 and r["x"].tolist() == [0, 1, 1, 1, 0, 1] and r["z"] == 15.0
    # visit order 0,1,1.1,1.2,1.2.1,1.2.2,1.2.2.1,1.2.2.2,1.2.2.2.1,1.2.2.2.1.1,2,2.1,2.1.1,2.1.2,... as depths
    assert r["node_log"][:, 0].tolist() == [0, 1, 2, 2, 3, 3, 4, 4, 5, 6, 1, 2, 3, 3, 4, 4, 5, 5, 6, 2]
    assert r["node_log"][0].tolist() == [0, 4, 0, 0]   # root branches on x5 = 0.2
    assert r["node_log"][1].tolist()[:2] == [1, 0]     # child 1 branches on x1
    assert r["node_log"][10].tolist()[:2] == [1, 3]    # child 2 branches on x4
    z = r["node_z"].tolist()
    assert z[:3] == [15.4, 15.3636, 15.0] and z[10] == 14.1429
    for artefact in (11.9996, 13.0005, 11.0008, 11.9998, 10.9999):
        assert artefact in z
    rp = O.bb_solve(Tf, 6, prune=True, max_nodes=20)
    assert rp["nodes"] == 5 and rp["x"].tolist() == [0, 1, 1, 1, 0, 1] and rp["z"] == 15.0


def test_appendix_c4_gomory_cut():
    Tf = final_a()["T"]
    row, cut = O.gomory_cut(Tf)
    assert row == 0
    np.testing.assert_allclose(cut, -np.array([0.1, 0, 0, 0, 0, 0, 0.1, 0, 0.2, 0.4, 0.6, 0, 0, 0.2]), atol=1e-12)
    cp = O.cutting_plane(Tf)
    assert cp["status"] == O.OPTIMAL and cp["log"].tolist() == [[0, 0, 1, 1]]
    assert cp["T"].shape == (9, 14) and cp["T"][0, -1] == 15.0
    assert all(abs(v - round(v)) < 1e-9 for v in cp["T"][1:, -1])


def test_appendix_c5_knapsack_and_reference_check():
    dp, ch = O.knap_dp(40, [11, 8, 6, 14, 10, 10], [2, 3, 3, 5, 2, 4])
    kb = O.knap_bb(40, [11, 8, 6, 14, 10, 10], [2, 3, 3, 5, 2, 4])
    assert dp == 15 and ch.tolist() == [0, 1, 1, 1, 0, 1]
    assert abs(dp - kb["best"]) < 1e-6 and kb["chosen"].tolist() == [0, 1, 1, 1, 0, 1]  # Program.cs:467-470
    for seed in range(20):
        w, v, cap = O.gen_knapsack(1000 + seed, 30 + seed)
        kb = O.knap_bb(cap, w, v)
        dp, _ = O.knap_dp(int(cap), w.astype(int), v.astype(int))
        assert kb["best"] == dp
        assert float(np.dot(kb["chosen"], w)) <= cap and float(np.dot(kb["chosen"], v)) == dp


def test_golden_vectors_reproduced():
    for name, (obj, cons) in (("A", MODEL_A), ("B", MODEL_B)):
        for cli in (True, False):
            g = GOLD[f"primal_{name}_{'cli' if cli else 'raw'}"]
            T, b = O.primal_build(obj, cons + (cli_rows(len(obj)) if cli else []))
            r = O.primal_solve(T, b)
            assert r["log"].tolist() == g["log"] and r["basis"].tolist() == g["basis"]
            assert hexs(r["T"]) == g["final"] and float(r["T"][0, -1]).hex() == g["z"]
    Tf = final_a()["T"]
    for prune in (False, True):
        g = GOLD[f"bb_A_prune{int(prune)}"]
        r = O.bb_solve(Tf, 6, prune=prune, max_nodes=20)
        assert r["node_log"].tolist() == g["node_log"] and hexs(r["node_z"]) == g["node_z"] and r["pivots"] == g["pivots"]
    g = GOLD["cut_A"]
    cp = O.cutting_plane(Tf)
    assert hexs(cp["T"]) == g["final"] and cp["log"].tolist() == g["log"]
    for key in [k for k in GOLD if k.startswith("dense_lp_")]:
        _, _, seed, shape = key.split("_")
        m, n = map(int, shape.split("x"))
        A, b, c = O.gen_dense_lp(int(seed), m, n)
        assert float(O.lib().orc_u01(int(seed), 0)).hex() == GOLD[key]["u0"] and float(A[0, 0]).hex() == GOLD[key]["a00"]
        T0, b0 = O.primal_build(list(c), [(A[i], "<=", b[i]) for i in range(m)])
        r = O.primal_solve(T0, b0)
        assert r["log"].tolist() == GOLD[key]["log"] and float(r["T"][0, -1]).hex() == GOLD[key]["z"]
    g = GOLD["knap_385_64"]
    kb = O.knap_bb(g["cap"], g["w"], g["v"])
    assert kb["best"] == g["best"] == g["dp"] and kb["chosen"].tolist() == g["chosen"]


def test_generator_is_language_independent():
    # splitmix64 restated in numpy: bit-identical stream (the CUDA generator uses the same arithmetic)
    def sm64(x):
        x = (x + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
        x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
        x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
        return x ^ (x >> 31)
    A, b, c = O.gen_dense_lp(381, 5, 7)
    for i in range(5):
        for j in range(7):
            assert A[i, j] == 0.1 + (sm64(381 + i * 7 + j) >> 11) * 2.0 ** -53
        assert b[i] == (7 / 4.0) * (1.0 + (sm64(381 + (1 << 40) + i) >> 11) * 2.0 ** -53)
    assert c[3] == 1.0 + (sm64(381 + (2 << 40) + 3) >> 11) * 2.0 ** -53


@pytest.mark.parametrize("m,n,seed", [(10, 20, 1), (25, 40, 2), (40, 30, 3)])
def test_against_scipy_highs(m, n, seed):
    from scipy.optimize import linprog
    A, b, c = O.gen_dense_lp(seed, m, n)
    T0, b0 = O.primal_build(list(c), [(A[i], "<=", b[i]) for i in range(m)])
    r = O.primal_solve(T0, b0)
    rr = O.rev_solve(A, b, c)
    hs = linprog(-c, A_ub=A, b_ub=b, bounds=[(0, None)] * n, method="highs")
    assert hs.status == 0
    assert abs(r["T"][0, -1] + hs.fun) <= 1e-9 * abs(hs.fun)
    assert abs(rr["z"] + hs.fun) <= 1e-9 * abs(hs.fun)
    assert r["basis"].tolist() == rr["basis"].tolist() or abs(rr["z"] - r["T"][0, -1]) < 1e-9 * abs(hs.fun)


def test_threaded_pivot_is_bit_identical():
    A, b, c = O.gen_dense_lp(4, 60, 90)
    T0, b0 = O.primal_build(list(c), [(A[i], "<=", b[i]) for i in range(60)])
    a = O.primal_solve(T0, b0, threads=1)
    t = O.primal_solve(T0, b0, threads=4)
    assert a["log"].tolist() == t["log"].tolist()
    assert np.array_equal(a["T"].view(np.uint64), t["T"].view(np.uint64))


def test_bb_quirks():
    # Q11: a column summing to 1 counts as "basic" even when it is not a unit column
    T = np.array([[0.0, 0.5, 0.0, 3.0], [1.0, 0.25, 0.0, 2.0], [0.0, 0.25, 1.0, 1.0]])
    basic = O.bb_identify_basic(T).tolist()
    assert 1 in basic and 0 in basic and 2 in basic
    # rounding is banker's at 4 d.p.
    x = np.array([[0.00005, 0.00015, -0.00005, 2.5e-5, 12345.67895, -0.99995, 3.14159265]])
    assert np.array_equal(O.bb_round(x).view(np.uint64), (np.rint(x * 1e4) / 1e4).view(np.uint64))
    assert O.lib().orc_net_round(2.5) == 2.0 and O.lib().orc_net_round(3.5) == 4.0 and O.lib().orc_net_round(-0.5) == 0.0
    # the "drop last tableau" quirk: negative RHS after the primal phase
    T = np.array([[-1.0, 0.0, 0.0, 0.0], [1.0, 1.0, 0.0, -1e-12], [1.0, 0.0, 1.0, 2.0]])
    r = O.bb_node_solve(T)
    assert r["status"] in (O.OPTIMAL, O.INFEASIBLE)


def test_package_generators_match_oracle():
    """the numpy generators the bench uses for cfg4/cfg5 are bit-identical to the oracle's"""
    from lpr_381_group_v22_b200.bench_workloads import gen_dense_ip, gen_knapsack
    A, b, c = gen_dense_ip(385, 9, 14)
    A2, b2, c2 = O.gen_dense_ip(385, 9, 14)
    assert np.array_equal(A, A2) and np.array_equal(b, b2) and np.array_equal(c, c2)
    w, v, cap = gen_knapsack(384, 100)
    w2, v2, cap2 = O.gen_knapsack(384, 100)
    assert np.array_equal(w, w2) and np.array_equal(v, v2) and cap == cap2


@pytest.mark.parametrize("m,n,seed", [(5, 7, 1), (9, 6, 2), (12, 20, 3)])
def test_sensitivity_add_constraint_against_highs(m, n, seed):
    """AddNewConstraintNonInteractive + ResolveAll (SensitivityAnalyzer.cs:609-659, :98-209, :706-723) restated.
    Reference quirk: the new row is built as -tech_j + sum_B tech_B * T[B, j] with slack +1 and
    RHS rhs - a.x, i.e. its technical coefficients carry the opposite sign of the textbook row; the restatement
    keeps that.  Fed with -tech the same code yields the textbook row, and then the re-solved tableau must be
    optimal for the LP with the extra row (HiGHS objective, 1e-9 relative) -- which pins everything else."""
    from scipy.optimize import linprog
    A, b, c = O.gen_dense_lp(100 + seed, m, n)
    T0, b0 = O.primal_build(list(c), [(A[i], "<=", b[i]) for i in range(m)])
    opt = O.primal_solve(T0, b0)
    assert opt["status"] == O.OPTIMAL
    T = opt["T"]
    basis = O.sens_rebuild_basis(T)
    assert sorted(basis.tolist()) == sorted(opt["basis"].tolist())
    x = O.sens_solution(T)
    tech = np.zeros(n + m)
    tech[:n] = 1.0 + np.arange(n) % 3
    rhs = 0.6 * float(tech @ x)          # cuts the current optimum off
    ax = 0.0
    for j in range(n + m):
        ax += tech[j] * x[j]
    for sign in (+1.0, -1.0):
        T1, b1 = O.sens_add_constraint(T, basis, sign * tech, rhs - ax)
        assert T1.shape == (m + 2, n + m + 2) and T1[-1, -1] < 0      # new row infeasible => dual simplex runs
        assert T1[-1, n + m] == 1.0 and np.array_equal(T1[:-1, -1], T[:, -1])
        b1 = O.sens_rebuild_basis(T1)
        assert b1[-1] == n + m                                          # the new slack is basic in the new row
        res = O.sens_resolve(T1, b1)
        if sign > 0:
            # the reference's own row: whatever it converges to satisfies its stopping rules (:85-96, :173-181)
            if res["status"] == O.OPTIMAL:
                Tr = res["T"]
                nb = [j for j in range(Tr.shape[1] - 1) if j not in res["basis"].tolist()]
                assert np.all(Tr[1:, -1] >= -1e-9) and np.all(Tr[0, nb] >= -1e-9)
            continue
        assert res["status"] == O.OPTIMAL and res["n_pivots"] >= 1
        A2 = np.vstack([A, tech[:n]])
        b2 = np.concatenate([b, [rhs]])
        hs = linprog(-c, A_ub=A2, b_ub=b2, bounds=[(0, None)] * n, method="highs")
        assert hs.status == 0
        assert res["T"][0, -1] == pytest.approx(-hs.fun, rel=1e-9)
        x2 = O.sens_solution(res["T"])
        assert float(tech[:n] @ x2[:n]) == pytest.approx(rhs, rel=1e-9)  # the new row is tight


def test_run_branch_and_bound_entry_restated():
    """RunBranchAndBound (BranchBoundSimplexSolver.cs:1253-1298): ConfigureProblem :1233-1251 appends x_i <= 1 rows
    that are one entry longer than a model row, FormulateTableau :28-113 negates '>=' rows (flag 1) and puts the
    identity at column i+n-1, DoDualSimplex :289-468 solves (isMin picks the optimality test / entering rule)."""
    from scipy.optimize import linprog
    obj = [2.0, 3.0, 3.0, 5.0, 2.0, 4.0]                       # model A (data/TextFile.txt)
    cons = [[11.0, 8.0, 6.0, 14.0, 10.0, 10.0, 40.0, 0.0]]
    o2, c2 = O.bb_configure_problem(obj, cons)
    assert len(c2) == 7 and len(c2[1]) == 9 and c2[3][2] == 1.0 and c2[3][7] == 1.0 and c2[3][8] == 0.0
    T = O.bb_formulate(o2, c2)
    assert T.shape == (8, 14) and T[0, :6].tolist() == [-2, -3, -3, -5, -2, -4]
    assert T[1, :6].tolist() == [11, 8, 6, 14, 10, 10] and T[1, -1] == 40 and T[1, 6] == 1
    assert np.array_equal(T[2:, 6:13], np.eye(7)[1:]) and np.all(T[2:, -1] == 1)   # the spilled 0 lands on column 6
    r = O.bb_node_solve_ex(T, False)
    hs = linprog(-np.array(obj), A_ub=[cons[0][:6]], b_ub=[40.0], bounds=[(0, 1)] * 6, method="highs")
    assert r["status"] == O.OPTIMAL and r["T"][0, -1] == pytest.approx(-hs.fun, rel=1e-12)
    bb = O.bb_solve(O.bb_round(r["T"]), 6, prune=False, max_nodes=20)
    assert bb["x"].tolist() == [0, 1, 1, 1, 0, 1] and bb["z"] == 15.0          # Appendix C3's incumbent
    # a '>=' row (flag 1) is negated entirely, zeros become -0.0 like `-1 * x` in C#
    Tg = O.bb_formulate([1.0, 2.0], [[1.0, 0.0, 3.0, 1.0], [1.0, 1.0, 10.0, 0.0]])
    assert Tg[1].tolist() == [-1.0, -0.0, 1.0, 0.0, -3.0] and np.signbit(Tg[1, 1])
    assert Tg[2].tolist() == [1.0, 1.0, 0.0, 1.0, 10.0]
    # isMin: optimal iff the objective row is <= 0; entering = smallest positive entry (:209-213, :346-348)
    Tm = np.array([[3.0, 1.0, 0.0, 0.0, 0.0], [1.0, 1.0, 1.0, 0.0, 4.0], [1.0, 3.0, 0.0, 1.0, 6.0]])
    rm = O.bb_node_solve_ex(Tm, True)
    assert rm["log"][0].tolist()[1] == 1 and rm["status"] == O.OPTIMAL   # column 1 (value 1) before column 0 (value 3)
    r0 = O.bb_node_solve_ex(np.array([[-3.0, -1.0, 0.0, 0.0, 0.0], [1.0, 1.0, 1.0, 0.0, 4.0], [1.0, 3.0, 0.0, 1.0, 6.0]]), True)
    assert r0["n_pivots"] == 0 and r0["status"] == O.OPTIMAL             # all <= 0: already "optimal" for a min
