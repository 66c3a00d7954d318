"""GPU parity tests for RevisedPrimalSimplexSolver (B^-1 resident in HBM) against the oracle.
Pivot sequence / basis / status exact; z, x, y, x_B within 1e-9 relative (north-star tolerance)."""
import numpy as np
import pytest

import oracle_lib as O
import lpr_381_group_v22_b200 as L

pytestmark = pytest.mark.gpu
RTOL = 1e-9


def close(a, b, what):
    a = np.asarray(a, dtype=float)
    b = np.asarray(b, dtype=float)
    scale = max(1.0, float(np.max(np.abs(b))) if b.size else 1.0)
    err = float(np.max(np.abs(a - b))) if b.size else 0.0
    assert err <= RTOL * scale, f"{what}: max abs err {err:.3e} (scale {scale:.3e})"


def solve_both(A, b, c, is_min=False, **kw):
    m, n = A.shape
    ref = O.rev_solve(A, b, c, is_min, want_binv=True)
    s = L.RevisedPrimalSimplexSolver(list(c), [L.Constraint(A[i], "<=", b[i]) for i in range(m)], is_min, **kw)
    return ref, s


def test_fixture_models():
    # README model through CLI option 2 ('+ + +' => no bound rows, '>=' ignored: Q6)
    A = np.array([[1, 2, 3], [3, 2, 1.0]])
    ref, s = solve_both(A, [10, 15], [2, 3, 4])
    s.Solve()
    assert s.PivotLog == [(0, 2, 3), (1, 0, 4)] == [tuple(x) for x in ref["log"].tolist()]
    assert s.BasicVariables == [2, 0] and s.FinalZ == 16.25 and s.SolutionVector == [4.375, 0.0, 1.875]
    # data/TextFile.txt through CLI option 2 ('bin' => six x_j <= 1 rows)
    A = np.vstack([[11, 8, 6, 14, 10, 10], np.eye(6)])
    b = [40] + [1] * 6
    obj = [2, 3, 3, 5, 2, 4]
    ref, s = solve_both(A, b, obj)
    s.Solve()
    assert s.PivotLog == [(4, 3, 10), (6, 5, 12), (2, 1, 8), (3, 2, 9), (0, 0, 6), (0, 4, 0)]
    assert s.BasicVariables == [4, 7, 1, 2, 3, 11, 5] == ref["basis"].tolist()
    close(s.FinalZ, 15.4, "z")
    close(s.SolutionVector, [0, 1, 1, 1, 0.2, 1], "x")
    close(s.DualPrices, ref["y"], "y")
    close(s.BInverse, ref["Binv"], "Binv")


@pytest.mark.parametrize("m,n,seed", [(6, 12, 1), (31, 64, 2), (64, 129, 3), (120, 75, 4), (200, 400, 5)])
@pytest.mark.parametrize("is_min", [False])
def test_random_dense_lp(m, n, seed, is_min):
    A, b, c = O.gen_dense_lp(seed, m, n)
    ref, s = solve_both(A, b, c, is_min)
    s.Solve()
    assert ref["status"] == O.OPTIMAL and s.Status == L.OPTIMAL
    assert s.PivotLog == [tuple(x) for x in ref["log"].tolist()]
    assert s.BasicVariables == ref["basis"].tolist()
    close(s.FinalZ, ref["z"], "z")
    close(s.SolutionVector, ref["x"], "x")
    close(s.DualPrices, ref["y"], "y")
    close(s.BasicValues, ref["xB"], "xB")
    close(s.BInverse, ref["Binv"], "Binv")


def test_minimization_and_exceptions():
    # min problem: c negated (:51); optimal at origin => zero iterations
    A, b, c = O.gen_dense_lp(7, 10, 20)
    ref, s = solve_both(A, b, c, True)
    s.Solve()
    assert s.PivotLog == [] and s.FinalZ == 0.0 == ref["z"]
    # unbounded: max x1 s.t. -x1 + x2 <= 1
    s = L.RevisedPrimalSimplexSolver([1.0, 0.0], [L.Constraint([-1.0, 1.0], "<=", 1.0)], False)
    with pytest.raises(Exception, match="Unbounded problem"):
        s.Solve()
    # infeasible start: negative rhs (Relation ignored => basis value negative)
    s = L.RevisedPrimalSimplexSolver([1.0, 1.0], [L.Constraint([1.0, 1.0], "<=", -2.0)], False)
    with pytest.raises(Exception, match="Infeasible basis"):
        s.Solve()
    with pytest.raises(ValueError):
        L.RevisedPrimalSimplexSolver([1.0, 1.0], [L.Constraint([1.0], "<=", 1.0)], False)


def test_agrees_with_tableau_solver_objective():
    m, n, seed = 80, 160, 11
    A, b, c = O.gen_dense_lp(seed, m, n)
    cons = [L.Constraint(A[i], "<=", b[i]) for i in range(m)]
    r = L.RevisedPrimalSimplexSolver(list(c), cons, False)
    r.Solve()
    p = L.PrimalSimplexSolver(list(c), cons, trace=False)
    p.Solve()
    close(r.FinalZ, p.FinalZ, "z revised vs tableau")
    # duals of the tableau solver = row 0 under the slack columns (SensitivityAnalyzer.cs:212-222)
    close(r.DualPrices, p.GetFinalTableau()[0, n:n + m], "duals")


def test_refactorisation_fp64_dmma():
    """Periodic refactorisation (Newton-Schulz refresh on the FP64 tensor cores) must not move results
    beyond the 1e-9 tolerance and must not change the pivot sequence."""
    import ctypes as C
    from lpr_381_group_v22_b200 import _native as N
    m, n, seed = 150, 260, 21
    A, b, c = O.gen_dense_lp(seed, m, n)
    ref = O.rev_solve(A, b, c, want_binv=True)
    cons = [L.Constraint(A[i], "<=", b[i]) for i in range(m)]
    s = L.RevisedPrimalSimplexSolver(list(c), cons, False, refactor_every=25)
    s.Solve()
    assert s.PivotLog == [tuple(x) for x in ref["log"].tolist()]
    close(s.FinalZ, ref["z"], "z")
    close(s.SolutionVector, ref["x"], "x")
    close(s.DualPrices, ref["y"], "y")
    close(s.BInverse, ref["Binv"], "Binv")
    # explicit refresh at the optimum: residual |I - B X| is small before, B^-1 barely moves
    before = s.BInverse
    N.check(N.lib().lpr_rev_refactor(s._h))
    res, fl = C.c_double(), C.c_double()
    N.check(N.lib().lpr_rev_last_refactor_info(s._h, C.byref(res), C.byref(fl)))
    after = s.BInverse
    assert res.value < 1e-9 and fl.value == 4.0 * 192 ** 3
    close(after, before, "Binv after refresh")
    basis = s.BasicVariables
    Bm = np.column_stack([A[:, v] if v < n else np.eye(m)[:, v - n] for v in basis])
    assert np.max(np.abs(Bm @ after - np.eye(m))) <= np.max(np.abs(Bm @ before - np.eye(m))) + 1e-13
    close(s.DualPrices, ref["y"], "y after refresh")


def test_cfg3_size_window_matches_oracle():
    """BASELINE cfg3 shape (m = 8192, n = 16384, seed 384, generated in HBM): the first 32 iterations against the
    oracle run on the same input (one ~90 s single-thread oracle run serves both device runs) -- pivot log exact,
    z / x / y / x_B within 1e-9 relative -- as shipped and with a refactorisation of B^-1 every 12 iterations (which
    must not move anything beyond the tolerance)."""
    import ctypes as C
    from lpr_381_group_v22_b200 import _native as N
    m, n, seed, K = 8192, 16384, 384, 32
    A, b, c = O.gen_dense_lp(seed, m, n)
    ref = O.rev_solve(A, b, c, False, max_iter=K, log_cap=K)
    del A
    assert ref["status"] == O.ITER_LIMIT and ref["n_iter"] == K
    lib = N.lib()
    for refactor_every in (0, 12):
        h = N.vp()
        N.check(lib.lpr_rev_create_dense_lp(0, seed, m, n, C.byref(h)))
        try:
            st, nit = C.c_int(), C.c_int64()
            log = np.zeros((K, 3), dtype=np.int32)
            N.check(lib.lpr_rev_solve(h, K, refactor_every, C.byref(st), C.byref(nit), N.pi(log), K))
            assert st.value == L.ITER_LIMIT and nit.value == K
            assert log.tolist() == ref["log"].tolist()
            basis = np.zeros(m, dtype=np.int32); x = np.zeros(n); y = np.zeros(m); xb = np.zeros(m); z = C.c_double()
            N.check(lib.lpr_rev_read_basis(h, N.pi(basis)))
            N.check(lib.lpr_rev_read_x(h, N.pd(x)))
            N.check(lib.lpr_rev_read_y(h, N.pd(y)))
            N.check(lib.lpr_rev_read_xb(h, N.pd(xb)))
            N.check(lib.lpr_rev_read_z(h, C.byref(z)))
            assert basis.tolist() == ref["basis"].tolist()
            close(z.value, ref["z"], "z")
            close(x, ref["x"], "x")
            close(y, ref["y"], "y")
            close(xb, ref["xB"], "xB")
            if refactor_every:
                res, fl = C.c_double(), C.c_double()
                N.check(lib.lpr_rev_last_refactor_info(h, C.byref(res), C.byref(fl)))
                assert res.value < 1e-9
        finally:
            lib.lpr_rev_destroy(h)


def _models_for_snapshots():
    yield "README", [2, 3, 4], [[1, 2, 3], [3, 2, 1.0]], [10, 15], False
    yield "TextFile", [2, 3, 3, 5, 2, 4], np.vstack([[11, 8, 6, 14, 10, 10], np.eye(6)]).tolist(), [40] + [1] * 6, False
    A, b, c = O.gen_dense_lp(5, 7, 9)
    yield "random max", list(c), A.tolist(), list(b), False
    A, b, c = O.gen_dense_lp(6, 5, 8)
    yield "random min", list(-c), A.tolist(), list(b), True


@pytest.mark.parametrize("case", list(_models_for_snapshots()), ids=lambda c: c[0])
def test_iteration_snapshots_match_capture_snapshot_restatement(case):
    """one CaptureSnapshot text block per iteration plus the "Optimal" block
    (RevisedPrimalSimplexSolver.cs:124-146, :226-246, :294-387) through lpr_rev_begin / lpr_rev_step /
    lpr_rev_format_snapshot, against the loop-for-loop Python restatement in tests/net_reference.py."""
    import net_reference as R
    _, obj, A, b, is_min = case
    snaps, log, x, z, basis = R.revised_solve_with_snapshots(obj, A, b, is_min)
    s = L.RevisedPrimalSimplexSolver(list(obj), [L.Constraint(A[i], "<=", b[i]) for i in range(len(A))], is_min, trace=True)
    s.Solve()
    assert s.PivotLog == log and s.BasicVariables == basis
    assert len(s.IterationSnapshots) == len(snaps) == len(log) + 1
    for k, (mine, ref) in enumerate(zip(s.IterationSnapshots, snaps)):
        assert mine == ref, f"snapshot {k} differs:\n{mine}\n---\n{ref}"
    close(s.FinalZ, z, "z")
    close(s.SolutionVector, x, "x")
    # the untraced path gives the same numbers and records no text
    q = L.RevisedPrimalSimplexSolver(list(obj), [L.Constraint(A[i], "<=", b[i]) for i in range(len(A))], is_min, trace=False)
    q.Solve()
    assert q.PivotLog == log and q.IterationSnapshots == [] and q.FinalZ == s.FinalZ


def test_result_file_of_a_revised_run_matches_restatement(tmp_path):
    """Program.cs:329-347 -> OutputFileWrite.WriteFullResults with the revised solver's IterationSnapshots."""
    import net_reference as R
    obj, A, b = [2, 3, 4], [[1, 2, 3], [3, 2, 1.0]], [10, 15]
    s = L.RevisedPrimalSimplexSolver(obj, [L.Constraint(A[i], "<=", b[i]) for i in range(2)], False)
    s.Solve()
    snaps, _, x, z, _ = R.revised_solve_with_snapshots(obj, A, b, False)
    assert s.IterationSnapshots == snaps


@pytest.mark.parametrize("m,n,seed", [(40, 70, 31), (150, 260, 32), (333, 500, 33)])
def test_full_refactorisation_from_the_basis_columns(m, n, seed):
    """The full path (blocked Gauss-Jordan inversion with partial pivoting on the FP64 tensor cores) rebuilds B^-1 from
    the basis columns ALONE: damage the inverse (noise, then garbage), refactorise, compare with numpy's inverse of the
    gathered basis; the guard of the automatic mode must pick the full path when the refresh would diverge."""
    import ctypes as C
    from lpr_381_group_v22_b200 import _native as N
    A, b, c = O.gen_dense_lp(seed, m, n)
    ref = O.rev_solve(A, b, c, want_binv=True)
    s = L.RevisedPrimalSimplexSolver(list(c), [L.Constraint(A[i], "<=", b[i]) for i in range(m)], False, trace=False)
    s.Solve()
    lib = N.lib()
    basis = s.BasicVariables
    Bm = np.column_stack([A[:, v] if v < n else np.eye(m)[:, v - n] for v in basis])
    exact = np.linalg.inv(Bm)
    good = s.BInverse
    rng = np.random.default_rng(seed)
    path, after = C.c_int(), C.c_double()

    def err(X):
        return float(np.max(np.abs(Bm @ X - np.eye(m))))

    # (1) explicit full refactorisation of an intact inverse
    N.check(lib.lpr_rev_refactor_ex(s._h, 2))
    N.check(lib.lpr_rev_last_refactor_path(s._h, C.byref(path), C.byref(after)))
    assert path.value == 2 and after.value < 1e-9
    X = s.BInverse
    assert err(X) < 1e-11 and np.max(np.abs(X - exact)) <= 1e-9 * max(1.0, np.max(np.abs(exact)))
    # (2) multiplicative noise of 1e-3 on every entry: too far for a one-step refresh, the guard takes the full path
    noisy = good * (1.0 + 1e-3 * rng.standard_normal(good.shape))
    N.check(lib.lpr_rev_write_binv(s._h, N.pd(N.f64(noisy))))
    N.check(lib.lpr_rev_refactor_ex(s._h, 0))
    N.check(lib.lpr_rev_last_refactor_path(s._h, C.byref(path), C.byref(after)))
    assert path.value == 2 and err(s.BInverse) < 1e-11
    # (2b) noise of 1e-11: the cheap refresh is enough and is what the guard picks
    noisy = good * (1.0 + 1e-11 * rng.standard_normal(good.shape))
    N.check(lib.lpr_rev_write_binv(s._h, N.pd(N.f64(noisy))))
    N.check(lib.lpr_rev_refactor_ex(s._h, 0))
    N.check(lib.lpr_rev_last_refactor_path(s._h, C.byref(path), C.byref(after)))
    assert path.value == 1 and err(s.BInverse) < 1e-11
    # (3) garbage: the refresh would diverge (max |I - B X| >= 0.5), the guard must take the full path
    N.check(lib.lpr_rev_write_binv(s._h, N.pd(N.f64(rng.standard_normal(good.shape)))))
    N.check(lib.lpr_rev_refactor_ex(s._h, 0))
    N.check(lib.lpr_rev_last_refactor_path(s._h, C.byref(path), C.byref(after)))
    assert path.value == 2
    X = s.BInverse
    assert err(X) < 1e-11 and np.max(np.abs(X - exact)) <= 1e-9 * max(1.0, np.max(np.abs(exact)))
    close(s.DualPrices, ref["y"], "y after the full refactorisation")
