"""GPU tests for the knapsack branch & bound and its DP arbiter (Program.cs:430-471) against the oracle.
Integer data: best value and selection bit-exact; B&B value == DP value (the reference's own check)."""
import ctypes as C

import numpy as np
import pytest

import oracle_lib as O
import lpr_381_group_v22_b200 as L
from lpr_381_group_v22_b200 import _native as N

pytestmark = pytest.mark.gpu


def test_program_cs_instance():
    cap, w, v = 40, [11, 8, 6, 14, 10, 10], [2, 3, 3, 5, 2, 4]  # Program.cs:433-435
    s = L.KnapsackBranchBoundSimplex(cap, [float(x) for x in w], [float(x) for x in v])
    best = s.Solve()
    chosen = s.GetSelectedItemsOriginal()
    dp = L.KnapsackBranchBoundSolver.Solve(cap, w, v)
    assert best == 15.0 and abs(dp - best) < 1e-6  # Program.cs:467-470
    assert [it.Id + 1 for it in chosen] == [2, 3, 4, 6]  # Appendix C5
    assert sum(int(it.Weight) for it in chosen) == 38
    ref = O.knap_bb(cap, w, v)
    assert s.chosen.tolist() == ref["chosen"].tolist()


@pytest.mark.parametrize("seed,n", [(1, 8), (2, 20), (3, 50), (4, 100), (5, 200), (6, 333)])
def test_random_instances_match_oracle_and_dp(seed, n):
    w, v, cap = O.gen_knapsack(seed, n)
    ref = O.knap_bb(cap, w, v)
    s = L.KnapsackBranchBoundSimplex(cap, w, v)
    best = s.Solve()
    assert best == ref["best"]
    assert s.chosen.tolist() == ref["chosen"].tolist()  # same DFS-first optimum
    dp, ch = L.KnapsackBranchBoundSolver.Solve(int(cap), w.astype(int), v.astype(int), return_chosen=True)
    odp, och = O.knap_dp(int(cap), w.astype(int), v.astype(int))
    assert dp == odp == best
    assert ch.tolist() == och.tolist()
    assert float(np.dot(ch, v)) == dp and float(np.dot(ch, w)) <= cap


@pytest.mark.parametrize("batch", [32, 1024])
def test_ties_are_deterministic_across_batch_sizes(batch, monkeypatch):
    # many equal-ratio items => many optimal selections; the DFS-first one must win for any batch
    rng = np.random.default_rng(7)
    n = 40
    w = rng.integers(1, 6, size=n).astype(float)
    v = w * 2.0
    cap = float(int(w.sum() // 3))
    ref = O.knap_bb(cap, w, v)
    monkeypatch.setenv("LPR_KNAP_BATCH", str(batch))
    s = L.KnapsackBranchBoundSimplex(cap, w, v)
    assert s.Solve() == ref["best"]
    assert s.chosen.tolist() == ref["chosen"].tolist()


def test_edge_cases():
    # nothing fits
    s = L.KnapsackBranchBoundSimplex(1.0, [5.0, 7.0], [10.0, 3.0])
    assert s.Solve() == 0.0 and s.chosen.tolist() == [0, 0]
    # everything fits
    s = L.KnapsackBranchBoundSimplex(100.0, [5.0, 7.0, 1.0], [10.0, 3.0, 4.0])
    assert s.Solve() == 17.0 and s.chosen.tolist() == [1, 1, 1]
    with pytest.raises(L.LprError):
        L.KnapsackBranchBoundSimplex(10.0, [0.0, 1.0], [1.0, 1.0]).Solve()


def test_export_import_roundtrip_keeps_answer():
    w, v, cap = O.gen_knapsack(11, 120)
    ref = O.knap_bb(cap, w, v)
    lib = N.lib()
    a, b = N.vp(), N.vp()
    N.check(lib.lpr_knap_create(0, cap, len(w), N.pd(w), N.pd(v), C.byref(a)))
    N.check(lib.lpr_knap_create(0, cap, len(w), N.pd(w), N.pd(v), C.byref(b)))
    done, st, left = C.c_int64(), C.c_int(), C.c_int64()
    N.check(lib.lpr_knap_run(a, 64, C.byref(done), C.byref(st)))  # expand a little on "rank 0"
    # "rank 1" starts empty: drop its root by exporting it
    buf = (C.c_uint8 * (1 << 22))()
    nbytes, nexp = C.c_int64(), C.c_int()
    N.check(lib.lpr_knap_export_nodes(b, 1, buf, len(buf), C.byref(nbytes), C.byref(nexp)))
    assert nexp.value == 1
    # steal half of rank 0's open nodes
    N.check(lib.lpr_knap_open_count(a, C.byref(left)))
    N.check(lib.lpr_knap_export_nodes(a, int(left.value // 2), buf, len(buf), C.byref(nbytes), C.byref(nexp)))
    N.check(lib.lpr_knap_import_nodes(b, buf, nbytes.value))
    results = []
    for h in (a, b):
        N.check(lib.lpr_knap_run(h, -1, C.byref(done), C.byref(st)))
        best, kb = C.c_double(), C.c_int()
        ch = np.zeros(len(w), dtype=np.uint8)
        key = np.zeros((len(w) + 63) // 64, dtype=np.uint64)
        N.check(lib.lpr_knap_get_incumbent(h, C.byref(best), ch.ctypes.data_as(N.bp), key.ctypes.data_as(N.u64p), C.byref(kb)))
        results.append((best.value, kb.value, key.copy(), ch.copy()))
        lib.lpr_knap_destroy(h)
    # reduce like the multi-GPU driver does: max value, then DFS-first key
    best = max(r[0] for r in results)
    assert best == ref["best"]


def test_cfg4_size_matches_oracle_and_dp():
    """BASELINE cfg4 (n = 10^4 weakly correlated items, seed 384): the reference's own check (Program.cs:467-470,
    B&B value == DP value within 1e-6) at the benchmark size, with the device DP, the oracle's DP and the oracle's
    depth-first B&B as arbiters; the selection must be the oracle's, bit for bit."""
    n, seed = 10000, 384
    w, v, cap = O.gen_knapsack(seed, n)
    ref = O.knap_bb(cap, w, v)                                   # ~10 s single thread
    odp = O.knap_dp_value(int(cap), w.astype(int), v.astype(int))  # ~20 s, O(capacity) memory
    s = L.KnapsackBranchBoundSimplex(cap, w, v)
    best = s.Solve()
    dp, ch = L.KnapsackBranchBoundSolver.Solve(int(cap), w.astype(int), v.astype(int), return_chosen=True)
    assert abs(best - dp) < 1e-6 and best == ref["best"] == odp == dp
    assert s.chosen.tolist() == ref["chosen"].tolist()
    assert float(np.dot(s.chosen, v)) == best and float(np.dot(s.chosen, w)) <= cap
    assert float(np.dot(ch, v)) == dp and float(np.dot(ch, w)) <= cap


@pytest.mark.parametrize("seed,n,batch", [(21, 120, 0), (22, 600, 0), (23, 2000, 0), (24, 600, 512), (384, 10000, 0)])
def test_cluster_kernel_and_pipeline_walk_the_same_tree(seed, n, batch, monkeypatch):
    """k_knap_narrow (one thread-block cluster looping over narrow levels) and the six-kernel pipeline share the stack
    discipline: same node count, same incumbent, same selection, whichever of them walks which levels (a small batch
    makes them alternate); node budgets that cut a level go to the pipeline."""
    from lpr_381_group_v22_b200.distributed import KnapPool
    w, v, cap = O.gen_knapsack(seed, n)
    if batch:
        monkeypatch.setenv("LPR_KNAP_BATCH", str(batch))
    got = {}
    for narrow in ("1", "0"):
        monkeypatch.setenv("LPR_KNAP_NARROW", narrow)
        p = KnapPool(cap, w, v)
        try:
            nodes = p.run(777)  # a budget in the middle of a level
            assert nodes == 777 or p.open_count() == 0
            while p.open_count() > 0:
                nodes += p.run(1 << 40)
            inc = p.get_incumbent()
            got[narrow] = (nodes, inc[0], inc[2].astype(np.uint8).tolist())
        finally:
            p.close()
    assert got["1"] == got["0"]
    if n <= 2000:
        ref = O.knap_bb(cap, w, v)
        assert got["1"][1] == ref["best"] and got["1"][2] == ref["chosen"].tolist()


@pytest.mark.parametrize("n_gpus", [1, 2, 4])
def test_knap_solve_mgpu_in_library(n_gpus):
    """lpr_knap_solve_mgpu (host threads + NCCL inside the library): value and selection are the oracle's for any
    GPU count."""
    if n_gpus > L.device_count():
        pytest.skip("needs %d GPUs" % n_gpus)
    w, v, cap = O.gen_knapsack(13, 400)
    ref = O.knap_bb(cap, w, v)
    s = L.KnapsackBranchBoundSimplex(cap, w, v, n_gpus=n_gpus)
    # n_gpus = 1 also goes through the multi-GPU driver here (no NCCL needed)
    import ctypes as C
    best, nodes, st = C.c_double(), C.c_int64(), C.c_int()
    ch = np.zeros(len(w), dtype=np.uint8)
    stats = N.MgpuStats()
    N.check(N.lib().lpr_knap_solve_mgpu(n_gpus, None, cap, len(w), N.pd(N.f64(w)), N.pd(N.f64(v)), -1, -1, 1e-3,
                                        C.byref(best), ch.ctypes.data_as(N.bp), C.byref(nodes), C.byref(st),
                                        C.byref(stats)))
    assert st.value == L.OPTIMAL and stats.open_left == 0
    assert best.value == ref["best"] and ch.tolist() == ref["chosen"].tolist()
    assert s.Solve() == ref["best"] and s.chosen.tolist() == ref["chosen"].tolist()


@pytest.mark.parametrize("batch", [64, 4096])
def test_node_budget_and_time_slices(batch, monkeypatch):
    """lpr_knap_run with a node budget stops on the budget (device-side accounting) and continues where it stopped;
    lpr_knap_run_timed returns between groups of levels; the answer does not depend on where the runs are cut."""
    monkeypatch.setenv("LPR_KNAP_BATCH", str(batch))
    w, v, cap = O.gen_knapsack(17, 300)
    ref = O.knap_bb(cap, w, v)
    from lpr_381_group_v22_b200.distributed import KnapPool
    p = KnapPool(cap, w, v)
    try:
        assert p.run(100) == 100  # exactly the budget while nodes are left
        total = 100
        guard = 0
        while p.open_count() > 0:
            total += p.run(1 << 40, max_seconds=1e-4)
            guard += 1
            assert guard < 100000
        inc = p.get_incumbent()
        assert inc[0] == ref["best"] and inc[2].astype(np.uint8).tolist() == ref["chosen"].tolist()
    finally:
        p.close()


@pytest.mark.parametrize("rpg", ["2", "4"])
def test_knap_mgpu_several_pools_per_device_steal_inside_the_device(rpg, monkeypatch):
    monkeypatch.setenv("LPR_MG_KNAP_RANKS_PER_GPU", rpg)
    w, v, cap = O.gen_knapsack(19, 500)
    ref = O.knap_bb(cap, w, v)
    s = L.KnapsackBranchBoundSimplex(cap, w, v, n_gpus=2)  # n_gpus > 1 selects the driver; ...
    import ctypes as C
    best, nodes, st = C.c_double(), C.c_int64(), C.c_int()
    ch = np.zeros(len(w), dtype=np.uint8)
    stats = N.MgpuStats()
    N.check(N.lib().lpr_knap_solve_mgpu(1, None, cap, len(w), N.pd(N.f64(w)), N.pd(N.f64(v)), -1, -1, 2e-4,
                                        C.byref(best), ch.ctypes.data_as(N.bp), C.byref(nodes), C.byref(st),
                                        C.byref(stats)))
    assert st.value == L.OPTIMAL and stats.open_left == 0 and stats.ranks_per_gpu == int(rpg)
    assert best.value == ref["best"] and ch.tolist() == ref["chosen"].tolist()
    del s


def test_knap_keep_stride_partitions_identical_pools():
    """Start-up partition of the multi-GPU driver: pools that expanded the same root by the same number of nodes are
    identical, each keeps every world-th record, and the union of the shares solves to the oracle's answer."""
    from lpr_381_group_v22_b200.distributed import KnapPool, better
    import ctypes as C
    w, v, cap = O.gen_knapsack(23, 400)
    ref = O.knap_bb(cap, w, v)
    world = 3
    pools = [KnapPool(cap, w, v) for _ in range(world)]
    try:
        for p in pools:
            while 0 < p.open_count() < 64 * world:
                p.run(64)
        counts = [p.open_count() for p in pools]
        assert len(set(counts)) == 1
        for r, p in enumerate(pools):
            N.check(N.lib().lpr_knap_keep_stride(p._h, r, world))
        assert sum(p.open_count() for p in pools) == counts[0]
        best = None
        for p in pools:
            while p.open_count() > 0:
                p.run(1 << 40)
            inc = p.get_incumbent()
            if better(inc, best):
                best = inc
        assert best[0] == ref["best"] and np.asarray(best[2]).astype(np.uint8).tolist() == ref["chosen"].tolist()
    finally:
        for p in pools:
            p.close()
