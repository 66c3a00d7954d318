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
