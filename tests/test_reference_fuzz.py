"""Differential fuzzing, CPU only and only where the reference tree exists (the build container): the reference's own
C# sources EXECUTED by oracle/csharp against the oracle on a few hundred more seeded random models than the committed
golden file holds -- degenerate ties, unbounded and infeasible models, '>=' and '=' rows, ragged coefficient lists,
iteration caps with and without printing.  Every comparison is bit for bit.  (The GPU path is compared with the
oracle on random models of the same families by the -m gpu tests.)
"""
import importlib.util
import os
import random

import numpy as np
import pytest

import oracle_lib as O

HERE = os.path.dirname(os.path.abspath(__file__))
pytestmark = pytest.mark.skipif(not os.path.isdir("/root/reference/LPR_381_Group_V22"),
                                reason="the reference tree is only in the build container")


@pytest.fixture(scope="module")
def run():
    spec = importlib.util.spec_from_file_location("make_reference_run", os.path.join(HERE, "golden", "make_reference_run.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    return gen.Runner()


def same(a, b):
    a = np.ascontiguousarray(a, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    return a.shape == b.shape and np.array_equal(a.view(np.uint64), b.view(np.uint64))


def unmat(m):
    return np.array([float.fromhex(h) for h in m["hex"]], dtype=np.float64).reshape(m["shape"])


def small(rng, lo, hi, frac=0.3):
    v = rng.randint(lo, hi)
    return v + rng.randint(0, 9) / 10 if rng.random() < frac else float(v)


@pytest.mark.parametrize("seed", range(4))
def test_primal_tableau_simplex(run, seed):
    rng = random.Random(1000 + seed)
    for _ in range(25):
        n, m = rng.randint(1, 10), rng.randint(1, 9)
        obj = [small(rng, -3, 9, 0.2) for _ in range(n)]
        cons = []
        for _ in range(m):
            k = n if rng.random() < 0.8 else rng.randint(1, n + 3)
            # few distinct values => many exact ties in both ratio tests
            co = [float(rng.choice([0, 0, 1, 1, 2, 3, -1])) if seed % 2 else small(rng, -2, 9) for _ in range(k)]
            cons.append((co, rng.choice(["<=", "<=", "<=", ">=", "="]), float(rng.choice([0, 0, 4, 6, 12, rng.randint(0, 40)]))))
        is_max = rng.random() < 0.8
        g, _ = run.primal(obj, cons, is_max)
        T0, b0 = O.primal_build(obj, cons, is_max)
        r = O.primal_solve(T0, b0)
        assert same(T0, unmat(g["initial_tableau"]))
        assert r["log"].tolist() == g["pivots"] and r["basis"].tolist() == g["basis"]
        assert same(r["T"], unmat(g["final_tableau"]))
        assert (r["status"] == O.OPTIMAL) == (g["status"] == "optimal")
        if g["status"] == "optimal":
            assert same(O.primal_extract(r["T"], n), [float.fromhex(h) for h in g["x"]])


@pytest.mark.parametrize("seed", range(4))
def test_primal2_and_dual_rules(run, seed):
    rng = random.Random(2000 + seed)
    for _ in range(20):
        n, m = rng.randint(1, 9), rng.randint(1, 8)
        T = [[0.0] * (n + m + 1) for _ in range(m + 1)]
        dual = rng.random() < 0.5
        for j in range(n):
            T[0][j] = float(rng.randint(0, 9)) if dual else -float(rng.randint(-2, 9))
        for i in range(m):
            for j in range(n):
                T[i + 1][j] = float(rng.choice([0, 1, 2, -1, -2, 3])) if seed % 2 else small(rng, -6, 6, 0.2)
            T[i + 1][n + i] = 1.0
            T[i + 1][-1] = float(rng.randint(-20, 10)) if dual else float(rng.choice([0, 0, 3, 6, rng.randint(0, 30)]))
        mi, ps = rng.choice([10000, 10000, 3, 1, 0]), rng.random() < 0.5
        if dual:
            g = run.dual(T, mi, ps)
            r = O.dual_solve(np.array(T), mi, ps)
            assert same(r["T"], unmat(g["final_tableau"]))
            if g["exception"] is None:
                assert (r["status"] == O.OPTIMAL) == g["returned"]
            else:
                assert r["status"] == O.PIVOT_TOO_SMALL
        else:
            g = run.primal2(T, mi, ps)
            r = O.primal2_solve(np.array(T), mi, ps)
            assert same(r["T"], unmat(g["final_tableau"])) and (r["status"] == O.OPTIMAL) == g["returned"]
            assert r["log"].tolist() == g["pivots"]


@pytest.mark.parametrize("seed", range(4))
def test_revised_simplex(run, seed):
    rng = random.Random(3000 + seed)
    for _ in range(20):
        n, m = rng.randint(1, 9), rng.randint(1, 8)
        c = [small(rng, -2, 9, 0.2) for _ in range(n)]
        A = [[float(rng.choice([0, 1, 1, 2, 3, -1])) if seed % 2 else small(rng, -2, 9) for _ in range(n)] for _ in range(m)]
        b = [float(rng.choice([0, 0, 5, 10, rng.randint(0, 40), -rng.randint(1, 3) if seed == 3 else 7])) for _ in range(m)]
        is_min = rng.random() < 0.3
        g = run.revised(c, A, b, ["<="] * m, is_min)
        r = O.rev_solve(np.array(A), b, c, is_min, want_binv=True)
        labels = [f"x{e + 1}" if e < n else f"S{e - n + 1}" for e in r["log"][:, 1].tolist()]
        assert labels == g["entering_labels"] and r["basis"].tolist() == g["basis"]
        assert same(r["Binv"], unmat(g["binv"]))
        if g["exception"] is None:
            assert r["status"] == O.OPTIMAL and float(r["z"]).hex() == g["final_z"]
            assert same(r["x"], [float.fromhex(h) for h in g["x"]])
        else:
            want = {"U": O.UNBOUNDED, "I": O.INFEASIBLE, "P": O.PIVOT_TOO_SMALL}[g["exception"][0]]
            assert r["status"] == want


@pytest.mark.parametrize("seed", range(3))
def test_cutting_plane_and_branch_and_bound(run, seed):
    rng = random.Random(4000 + seed)
    done = 0
    while done < 4:
        n, m = rng.randint(2, 5), rng.randint(1, 4)
        obj = [float(rng.randint(1, 9)) for _ in range(n)]
        cons = [([float(rng.randint(0, 9)) for _ in range(n)], "<=", float(rng.randint(5, 40))) for _ in range(m)]
        if rng.random() < 0.8:
            cons += [([1.0 if j == i else 0.0 for j in range(n)], "<=", float(rng.randint(1, 4))) for i in range(n)]
        prec, p = run.primal(obj, cons, True)
        if prec["status"] != "optimal":
            continue
        done += 1
        T = unmat(prec["final_tableau"])
        run.it.max_steps = run.it.steps + 4_000_000
        try:
            g = run.cutting_plane(T.tolist())
            r = O.cutting_plane(T, extra_rows=64)
            assert same(r["T"], unmat(g["final_tableau"]))
        except RuntimeError:
            pass                                    # a cut sequence longer than the step budget
        finally:
            run.it.max_steps = None
        prune = rng.random() < 0.5
        g = run.bb(obj, cons, prune)
        r = O.bb_solve(T, g["n_vars"], prune=prune, max_nodes=20)
        assert r["nodes"] == len(g["nodes"])
        assert r["node_log"][:, 0].tolist() == [nd["depth"] for nd in g["nodes"]]
        x = [float.fromhex(h) for h in g["x"]]
        if r["has_solution"]:
            assert same(r["x"], x) and float(r["z"]).hex() == g["z"]
        else:
            assert x == [] and g["z"] == "-inf"
