"""GPU parity tests for BranchBoundSimplexSolver (AddConstraint / DoDualSimplex / ExecuteBranchAndBound)
through the C ABI against the CPU oracle.  Bit-exact tableaux, identical visit order and incumbent."""
import numpy as np
import pytest

import oracle_lib as O
import lpr_381_group_v22_b200 as L

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


def model_a_final():
    obj = [2, 3, 3, 5, 2, 4]
    cons = [L.Constraint([11, 8, 6, 14, 10, 10], "<=", 40)]
    L.add_cli_bound_rows(6, cons)
    s = L.PrimalSimplexSolver(obj, cons)
    s.Solve()
    return s


def ip_final(seed, m, n):
    A, b, c = O.gen_dense_ip(seed, m, n)
    T0, b0 = O.primal_build(list(c), [(A[i], "<=", b[i]) for i in range(m)])
    return O.primal_solve(T0, b0)["T"]


def test_round4_matches_oracle():
    rng = np.random.default_rng(0)
    T = rng.normal(size=(13, 29)) * 10
    T[0, :5] = [0.00005, -0.00005, 0.00015, -0.00004, 2.5e-5]
    T[1, :4] = [1e17, -1e17, 0.49999999999999994e-4, 12345.67895]
    with L.DeviceTableau.from_host(T) as t:
        t.round4()
        assert_bit_equal(t.read(), O.bb_round(T))


def test_round4_division_free_range_boundary():
    """net_round4 computes k / 1e4 without a division for |x| < 2e5 (tools/div1e4_check.c) and literally above:
    both sides of the boundary, halfway cases and signed zeros against the oracle's literal form."""
    rng = np.random.default_rng(5)
    xs = np.concatenate([
        rng.uniform(-2.2e5, 2.2e5, 4000), rng.uniform(-3, 3, 4000), rng.integers(-2_000_000, 2_000_000, 2000) / 1e4,
        (rng.integers(-10**9, 10**9, 2000) + 0.5) / 1e4,
        [199999.99995, 2e5, -2e5, 200000.00005, 214748.3648, 214748.36485, -0.0, 0.0, -4e-5, 4e-5, 1.00005, 0.99995,
         np.nan, np.inf, -np.inf, 1e16, 9.99e15],
    ])
    T = np.zeros((41, 300))
    T.flat[:xs.size] = xs
    with L.DeviceTableau.from_host(T) as t:
        t.round4()
        assert_bit_equal(t.read(), O.bb_round(T))


@pytest.mark.parametrize("poison", [3.1e5, -7e8, 1e17])
def test_add_constraint_literal_path_for_large_entries(poison):
    """One entry outside the division-free range sends the whole job through the literal row-order column sums."""
    Tf = np.array(ip_final(31, 7, 9), copy=True)
    Tf[3, 11] = poison
    for var, typ in ((0, 0), (4, 1)):
        ref = O.bb_add_constraint(Tf, 9, var, 2.0, typ)
        with L.DeviceTableau.from_host(Tf) as t:
            with t.bb_add_constraint(9, var, 2.0, typ) as ch:
                assert_bit_equal(ch.read(), ref)


@pytest.mark.parametrize("seed,m,n", [(41, 9, 14), (42, 20, 33), (43, 33, 64)])
def test_add_constraint_random_ip(seed, m, n):
    """Row-segment / column-pair geometry of k_bb_addc_build on shapes that straddle its tile edges."""
    Tf = ip_final(seed, m, n)
    x = O.bb_extract(O.bb_round(Tf), n)
    for var in (0, n // 2, n - 1):
        for typ in (0, 1):
            bound = float(np.floor(x[var]) if typ == 0 else np.ceil(x[var]))
            ref = O.bb_add_constraint(Tf, n, var, bound, typ)
            with L.DeviceTableau.from_host(Tf) as t:
                with t.bb_add_constraint(n, var, bound, typ) as ch:
                    assert_bit_equal(ch.read(), ref)


def test_net_round_equals_rint_on_edges():
    xs = [0.5, 1.5, 2.5, -0.5, -1.5, 0.49999999999999994, -0.49999999999999994, 1e15 + 0.5, 4503599627370497.0,
          0.0, -0.0, 2.4999999999999996, 1e300, -3.5, 7.5e-5 * 1e4]
    for x in xs:
        a, b = O.lib().orc_net_round(x), float(np.rint(x))
        assert np.float64(a).view(np.uint64) == np.float64(b).view(np.uint64), x


@pytest.mark.parametrize("var,typ", [(4, 0), (4, 1), (0, 0), (2, 1)])
def test_add_constraint_model_a(var, typ):
    s = model_a_final()
    Tf = s.GetFinalTableau()
    val = O.bb_extract(O.bb_round(Tf), 6)[var]
    bound = np.floor(val) if typ == 0 else np.ceil(val)
    ref = O.bb_add_constraint(Tf, 6, var, bound, typ)
    with L.DeviceTableau.from_host(Tf) as t:
        with t.bb_add_constraint(6, var, bound, typ) as ch:
            assert_bit_equal(ch.read(), ref)


def test_branch_var_and_node_solve_model_a():
    s = model_a_final()
    Tf = O.bb_round(s.GetFinalTableau())
    with L.DeviceTableau.from_host(Tf) as t:
        var, val, x = t.bb_branch_var(6)
        rv, rval = O.bb_branch_var(Tf, 6)
        assert (var, val) == (rv, rval) == (4, 0.2)
        assert_bit_equal(x, O.bb_extract(Tf, 6), "x")
        for typ, bound in ((0, 0.0), (1, 1.0)):
            ref_child = O.bb_add_constraint(Tf, 6, var, bound, typ)
            ref = O.bb_node_solve(ref_child)
            with t.bb_add_constraint(6, var, bound, typ) as ch:
                r = ch.bb_node_solve()
                assert r["status"] == ref["status"] == L.OPTIMAL
                assert r["log"].tolist() == ref["log"].tolist()
                assert_bit_equal(ch.read(), ref["T"])


def test_bb_model_a_appendix_c3():
    s = model_a_final()
    x, z = L.BranchAndBoundAdapter.SolveFromPrimal(s, enablePruning=False, isMin=False)
    run = L.BranchAndBoundAdapter.LastRun
    ref = O.bb_solve(s.GetFinalTableau(), 6, prune=False, max_nodes=20)
    assert x == [0.0, 1.0, 1.0, 1.0, 0.0, 1.0] and z == 15.0
    assert run["nodes"] == ref["nodes"] == 20
    assert run["node_log"].tolist() == ref["node_log"].tolist()
    assert_bit_equal(run["node_z"], ref["node_z"], "node z")
    assert run["pivots"] == ref["pivots"]
    x, z = L.BranchAndBoundAdapter.SolveFromPrimal(s, enablePruning=True)
    assert L.BranchAndBoundAdapter.LastRun["nodes"] == 5 and z == 15.0


@pytest.mark.parametrize("seed,m,n", [(11, 3, 5), (12, 4, 6), (13, 5, 8), (14, 6, 6), (15, 8, 12)])
@pytest.mark.parametrize("prune", [False, True])
def test_bb_random_ip_matches_oracle(seed, m, n, prune):
    Tf = ip_final(seed, m, n)
    cap = 60
    ref = O.bb_solve(Tf, n, prune=prune, max_nodes=cap)
    bb = L.BranchBoundSimplexSolver.BranchAndBound()
    bb.SetNumVars(n)
    x, z = bb.ExecuteBranchAndBound([Tf], prune, max_nodes=cap)
    run = bb.LastRun
    assert run["nodes"] == ref["nodes"]
    assert run["status"] == ref["status"]
    assert run["node_log"].tolist() == ref["node_log"].tolist()
    assert_bit_equal(run["node_z"], ref["node_z"], "node z")
    assert run["has_solution"] == ref["has_solution"]
    if ref["has_solution"]:
        assert z == ref["z"]
        assert_bit_equal(np.array(x), ref["x"], "incumbent")


def test_bb_cfg5_size_first_nodes_match_oracle():
    """BASELINE cfg5 shape (root 513 x 1537): the first 48 nodes of ExecuteBranchAndBound -- node log, node objective
    values, pivot count -- against the oracle, so the node kernels (row-segmented AddConstraint build, segmented
    evaluation, machine-wide sweep grid) are checked at the size the bench runs them, not only on toy tableaux."""
    m, n, seed = 512, 1024, 385
    A, b, c = O.gen_dense_ip(seed, m, n)
    T0, b0 = O.primal_build(list(c), [(A[i], "<=", b[i]) for i in range(m)])
    lp = O.primal_solve(T0, b0, threads=8)
    assert lp["status"] == O.OPTIMAL
    Tf = lp["T"]
    ref = O.bb_solve(Tf, n, prune=True, max_nodes=48)
    bb = L.BranchBoundSimplexSolver.BranchAndBound()
    bb.SetNumVars(n)
    x, z = bb.ExecuteBranchAndBound([Tf], True, max_nodes=48)
    run = bb.LastRun
    assert run["nodes"] == ref["nodes"] == 48 and run["pivots"] == ref["pivots"]
    assert run["node_log"].tolist() == ref["node_log"].tolist()
    assert_bit_equal(run["node_z"], ref["node_z"], "node z")
    assert run["has_solution"] == ref["has_solution"]
    # one AddConstraint at this size, every element
    Tr = O.bb_round(Tf)
    val = O.bb_extract(Tr, n)
    var, _ = O.bb_branch_var(Tr, n)
    with L.DeviceTableau.from_host(Tr) as t:
        for typ in (0, 1):
            bound = float(np.floor(val[var]) if typ == 0 else np.ceil(val[var]))
            with t.bb_add_constraint(n, var, bound, typ) as ch:
                assert_bit_equal(ch.read(), O.bb_add_constraint(Tr, n, var, bound, typ))


@pytest.mark.parametrize("seed,m,n", [(21, 6, 9), (22, 8, 10), (23, 10, 14)])
def test_bb_pool_batched_equals_sequential(seed, m, n):
    """Throughput mode (batches of open nodes, pruning on): same incumbent as the sequential oracle."""
    import ctypes as C
    from lpr_381_group_v22_b200 import _native as N
    Tf = ip_final(seed, m, n)
    ref = O.bb_solve(Tf, n, prune=True, max_nodes=-1)
    h = N.vp()
    N.check(N.lib().lpr_bb_create(0, Tf.shape[0], Tf.shape[1], N.pd(N.f64(Tf)), n, 1, C.byref(h)))
    done, piv, left = C.c_int64(), C.c_int64(), C.c_int64()
    total = 0
    while True:
        N.check(N.lib().lpr_bb_run(h, 64, C.byref(done), C.byref(piv)))
        total += done.value
        N.check(N.lib().lpr_bb_open_count(h, C.byref(left)))
        if left.value == 0:
            break
    has, z, klen = C.c_int(), C.c_double(), C.c_int(0)
    x = np.zeros(n)
    N.check(N.lib().lpr_bb_get_incumbent(h, C.byref(has), C.byref(z), N.pd(x), None, C.byref(klen)))
    N.lib().lpr_bb_destroy(h)
    assert bool(has.value) == ref["has_solution"]
    if ref["has_solution"]:
        assert z.value == ref["z"]
        assert_bit_equal(x, ref["x"], "incumbent")
    assert total >= 1


@pytest.mark.parametrize("seed,m,n,world", [(21, 6, 9, 2), (23, 10, 14, 3)])
def test_bb_pool_replicated_root_partition_and_time_slices(seed, m, n, world):
    """The multi-GPU start-up without a transfer, on one device: `world` pools expand the same root identically
    (lpr_bb_run), each keeps every world-th open node (lpr_bb_keep_stride), then runs in time slices
    (lpr_bb_run_timed).  The best incumbent over the pools (value, then DFS key) is the sequential oracle's."""
    from lpr_381_group_v22_b200.distributed import BBPool, better
    Tf = ip_final(seed, m, n)
    ref = O.bb_solve(Tf, n, prune=True, max_nodes=-1)
    pools = [BBPool(Tf, n, prune=True) for _ in range(world)]
    try:
        for p in pools:
            guard = 0
            while 0 < p.open_count() < 3 * world and guard < 32:
                p.run(2)
                guard += 1
        counts = [p.open_count() for p in pools]
        assert len(set(counts)) == 1  # identical expansion
        for r, p in enumerate(pools):
            p.keep_stride(r, world)
        assert sum(p.open_count() for p in pools) == counts[0]
        best = None
        for p in pools:
            slices = 0
            while p.open_count() > 0:
                done = p.run(1 << 30, max_seconds=1e-4)  # a slice ends after the batch that crosses it
                assert done > 0
                slices += 1
                assert slices < 10000
            inc = p.get_incumbent()
            if better(inc, best):
                best = inc
        assert (best is not None) == ref["has_solution"]
        if best is not None:
            assert best[0] == ref["z"]
            assert_bit_equal(np.asarray(best[2]), ref["x"], "incumbent")
    finally:
        for p in pools:
            p.close()


@pytest.mark.parametrize("m,n,seed", [(1, 6, 0), (5, 8, 1), (12, 9, 2), (30, 40, 3)])
def test_formulate_and_run_branch_and_bound(m, n, seed):
    """DualSimplexSolverBB.FormulateTableau / PrepareInput / DoDualSimplex(isMin) and
    BranchAndBound.ConfigureProblem / RunBranchAndBound (BranchBoundSimplexSolver.cs:28-113, :281-468,
    :1233-1298) on the device == the oracle, bit for bit."""
    rng = np.random.default_rng(seed)
    if m == 1:
        obj = [2.0, 3.0, 3.0, 5.0, 2.0, 4.0]
        cons = [[11.0, 8.0, 6.0, 14.0, 10.0, 10.0, 40.0, 0.0]]
    else:
        A, b, c = O.gen_dense_ip(200 + seed, m, n)
        obj = list(c)
        cons = [list(A[i]) + [float(b[i]), 0.0] for i in range(m)]
        cons[1] = [1.0] * n + [1.0, 1.0]          # a '>=' row (flag 1): sum x >= 1
    ref_o, ref_c = O.bb_configure_problem(obj, cons)
    Tref = O.bb_formulate(ref_o, ref_c)
    solver = L.BranchBoundSimplexSolver.DualSimplexSolverBB()
    mine = [list(r) for r in ref_c]
    T = solver.FormulateTableau(list(ref_o), mine)
    assert_bit_equal(T, Tref, "FormulateTableau")
    assert all(len(a) == len(r) - 1 for a, r in zip(mine, ref_c))      # the caller's rows lost their flags (:54-56)
    if m > 1:
        assert mine[1][0] == -1.0                                        # and the '>=' row was negated in place
    _, is_min, sur, slk, nobj = solver.PrepareInput(list(ref_o), [list(r) for r in ref_c], False)
    assert (is_min, sur, slk, nobj) == (False, 0 if m == 1 else 1, len(ref_c) - (0 if m == 1 else 1), len(obj))
    for is_min in (False, True):
        ref = O.bb_node_solve_ex(Tref, is_min)
        Td, opt, prow, pcol = solver.DoDualSimplex(list(ref_o), [list(r) for r in ref_c], is_min)
        if ref["status"] == O.OPTIMAL:
            assert_bit_equal(Td, ref["T"], f"DoDualSimplex isMin={is_min}")
            assert opt == ref["T"][0, -1] and list(zip(prow, pcol)) == [tuple(p) for p in ref["log"].tolist()]
        else:
            assert opt is None
    # the whole entry point
    ref = O.bb_node_solve_ex(Tref, False)
    bb = L.BranchBoundSimplexSolver.BranchAndBound()
    if ref["status"] != O.OPTIMAL:
        with pytest.raises(L.InvalidOperationException):
            bb.RunBranchAndBound(obj, cons, False)
        return
    rbb = O.bb_solve(O.bb_round(ref["T"]), len(obj), prune=False, max_nodes=20)
    x, z = bb.RunBranchAndBound(obj, cons, False)
    assert bb.LastRun["nodes"] == rbb["nodes"]
    assert bb.LastRun["node_log"].tolist() == rbb["node_log"].tolist()
    if rbb["has_solution"]:
        assert x == rbb["x"].tolist() and z == rbb["z"]
    else:
        assert x is None
    if m == 1:
        assert x == [0, 1, 1, 1, 0, 1] and z == 15.0


def binary_ip_final(seed, m, n, div):
    """cfg5-family instance that terminates under the reference's semantics: the cfg5 generator's A and c, b = floor(row
    sum / div), plus one `x_j <= 1` row per variable (what menu option 3 appends, Program.cs:372-382)."""
    A, b, c = O.gen_dense_ip(seed, m, n)
    b = np.floor(A.sum(axis=1) / div)
    cons = [(A[i], "<=", b[i]) for i in range(m)] + [(np.eye(n)[j], "<=", 1.0) for j in range(n)]
    T0, b0 = O.primal_build(list(c), cons)
    lp = O.primal_solve(T0, b0)
    assert lp["status"] == O.OPTIMAL
    return lp["T"]


MID = [(31, 8, 24, 3.0), (32, 8, 26, 3.0), (42, 6, 32, 2.0), (43, 6, 30, 3.0)]  # 33..39 rows, 279 .. 1665 oracle nodes


@pytest.mark.parametrize("seed,m,n,div", MID)
def test_bb_pool_batch64_equals_sequential_oracle_mid_size(seed, m, n, div):
    """The batched pool the bench runs (64 nodes = 128 children per batch, pruning on) against the sequential oracle
    on mid-size instances whose trees close (hundreds to thousands of nodes, depth ~ 25): same incumbent, bit for bit."""
    import ctypes as C
    from lpr_381_group_v22_b200 import _native as N
    Tf = binary_ip_final(seed, m, n, div)
    ref = O.bb_solve(Tf, n, prune=True, max_nodes=-1, log_cap=1 << 16)
    assert ref["status"] == O.OPTIMAL and ref["has_solution"]
    h = N.vp()
    N.check(N.lib().lpr_bb_create(0, Tf.shape[0], Tf.shape[1], N.pd(N.f64(Tf)), n, 1, C.byref(h)))
    done, piv, left, ovf = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int64()
    total = 0
    while True:
        N.check(N.lib().lpr_bb_run(h, 1 << 20, C.byref(done), C.byref(piv)))
        total += done.value
        N.check(N.lib().lpr_bb_open_count(h, C.byref(left)))
        if left.value == 0:
            break
    N.check(N.lib().lpr_bb_stats(h, None, None, C.byref(ovf), None))
    has, z, klen = C.c_int(), C.c_double(), C.c_int(0)
    x = np.zeros(n)
    N.check(N.lib().lpr_bb_get_incumbent(h, C.byref(has), C.byref(z), N.pd(x), None, C.byref(klen)))
    N.lib().lpr_bb_destroy(h)
    assert ovf.value == 0 and has.value == 1
    assert z.value == ref["z"]
    assert_bit_equal(x, ref["x"], "incumbent")


@pytest.mark.parametrize("n_gpus", [1, 2, 3])
def test_bb_solve_mgpu_in_library_matches_oracle(n_gpus):
    """lpr_bb_solve_mgpu: host threads + NCCL inside the library.  With fewer physical GPUs than ranks the ranks
    share devices only when NCCL allows it, so on a 1-GPU box the N > 1 cases are skipped."""
    if n_gpus > L.device_count():
        pytest.skip("needs %d GPUs" % n_gpus)
    seed, m, n, div = MID[2]
    Tf = binary_ip_final(seed, m, n, div)
    ref = O.bb_solve(Tf, n, prune=True, max_nodes=-1, log_cap=1 << 16)
    r = L.solve_bb_mgpu(Tf, n, True, n_gpus=n_gpus, slice_seconds=2e-3)
    assert r["status"] == L.OPTIMAL and r["has_solution"] and r["stats"]["open_left"] == 0
    assert r["z"] == ref["z"]
    assert_bit_equal(r["x"], ref["x"], "incumbent")
    assert r["stats"]["n_gpus"] == n_gpus and (n_gpus == 1 or r["stats"]["nccl_version"] > 0)


def test_bb_depth_overflow_is_reported(monkeypatch):
    """A subtree cut at the slab depth headroom must not come back as OPTIMAL."""
    monkeypatch.setenv("LPR_BB_MAX_DEPTH", "3")
    seed, m, n, div = MID[0]
    Tf = binary_ip_final(seed, m, n, div)
    bb = L.BranchBoundSimplexSolver.BranchAndBound()
    bb.SetNumVars(n)
    bb.ExecuteBranchAndBound([Tf], True, max_nodes=-1)
    assert bb.LastRun["status"] == L.DEPTH_LIMIT


@pytest.mark.parametrize("rpg", ["1", "2", "3"])
def test_bb_solve_mgpu_several_pools_per_device(rpg, monkeypatch):
    """ranks_per_gpu > 1: several pools (host thread + stream each) share one device; their status vectors are combined
    in host memory and node records move with device-to-device copies (no NCCL inside a device).  Same incumbent."""
    monkeypatch.setenv("LPR_MG_RANKS_PER_GPU", rpg)
    seed, m, n, div = MID[3]
    Tf = binary_ip_final(seed, m, n, div)
    ref = O.bb_solve(Tf, n, prune=True, max_nodes=-1, log_cap=1 << 16)
    r = L.solve_bb_mgpu(Tf, n, True, n_gpus=1, slice_seconds=1e-3)
    assert r["status"] == L.OPTIMAL and r["stats"]["open_left"] == 0 and r["stats"]["ranks_per_gpu"] == int(rpg)
    assert r["z"] == ref["z"]
    assert_bit_equal(r["x"], ref["x"], "incumbent")
