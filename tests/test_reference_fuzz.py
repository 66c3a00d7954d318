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
            r = O.cutting_plane(T, extra_rows=64, literal_sort=True)
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


# ------------------------------------------------------------------------------------------- native host code
TOKENS = ["1", "2", "-3", "+4", "0", "2.5", "-0.5", "1e1", "1,000", ".5", "5.", "x", "", "<=", ">=", "=", "max", "min", "MAX",
          "bin", "+", "-", "urs", "int", "1e", "--1", "0x10", "NaN", "1e999", "\t7"]


def _random_model_text(rng):
    n = rng.randint(1, 4)
    lines = []
    head = [rng.choice(["max", "min", "MAX", "Min", "maximize", "x"])] + [rng.choice(TOKENS[:12]) for _ in range(n)]
    lines.append((" " if rng.random() < 0.1 else "") + " ".join(head) + (" " if rng.random() < 0.2 else ""))
    for _ in range(rng.randint(0, 4)):
        k = n + 2 + rng.choice([0, 0, 0, 0, -1, -2, 1, 2])
        toks = [rng.choice(TOKENS[:10]) for _ in range(max(0, n))] + [rng.choice(["<=", ">=", "=", "<", "5"])] + [rng.choice(TOKENS[:11])]
        toks = toks[:max(0, k)] if k < len(toks) else toks + [rng.choice(TOKENS) for _ in range(k - len(toks))]
        if rng.random() < 0.15:
            toks[rng.randrange(len(toks))] = rng.choice(TOKENS) if toks else "1"
        sep = "  " if rng.random() < 0.2 else " "
        lines.append(("  " if rng.random() < 0.2 else "") + sep.join(toks))
    if rng.random() < 0.9:
        lines.append(" ".join(rng.choice(["+", "-", "urs", "bin", "int"]) for _ in range(rng.randint(0, n + 1))))
    eol = rng.choice(["\n", "\r\n", "\n", "\r"])
    return eol.join(lines) + (eol if rng.random() < 0.5 else "")


def test_native_parser_against_the_executed_reference_parser(run):
    """InputFileParser.ReadInputFile as the reference executes it, against lpr_model_parse_text (csrc/host_io.cu), on
    random texts: which exception comes first, the two silent early returns, every parsed value"""
    from lpr_381_group_v22_b200.io import Model
    rng = random.Random(5000)
    seen = {"ok": 0, "FormatException": 0, "IndexOutOfRangeException": 0, "early": 0}
    for _ in range(400):
        text = _random_model_text(rng)
        g = run.parse(text)
        if g["exception"] == "FormatException":
            with pytest.raises(ValueError, match="FormatException"):
                Model.parse_text(text)
            seen["FormatException"] += 1
            continue
        if g["exception"] == "IndexOutOfRangeException":
            with pytest.raises(IndexError, match="IndexOutOfRangeException"):
                Model.parse_text(text)
            seen["IndexOutOfRangeException"] += 1
            continue
        if g["exception"] == "OverflowException":          # 1e999: double.Parse overflows on the Framework
            with pytest.raises(ValueError):
                Model.parse_text(text)
            continue
        assert g["exception"] is None, (text, g["exception"])
        m = Model.parse_text(text)
        assert m.message == g["console"].strip(), text
        if g["problem_type"] is None:
            assert m.info()[0] is False
            seen["early"] += 1
            continue
        seen["ok"] += 1
        assert m.problem_type == g["problem_type"], text
        assert same(m.objective(), [float.fromhex(h) for h in g["objective"]]), text
        cons = m.constraints()
        assert len(cons) == len(g["constraints"]), text
        for c, (co, rel, rhs) in zip(cons, g["constraints"]):
            assert same(c.Coefficients, [float.fromhex(h) for h in co]) and c.Relation == rel and float(c.RHS).hex() == rhs, text
        assert m.signs() == g["signs"], text
    assert min(seen.values()) >= 5, seen


def test_native_result_files_against_the_executed_output_file_write(run, tmp_path):
    """OutputFileWrite / CanonicalFormForFile executed by the interpreter against lpr_out_* on random models, snapshot
    strings and special values"""
    import lpr_381_group_v22_b200 as L
    from lpr_381_group_v22_b200.io import Model
    from csharp.csrun import CsList, to_list
    rng = random.Random(6000)
    it = run.it
    specials = [0.0, -0.0, 1e-13, -1e-13, 0.0005, 2.5, 1e15, 1e16, 123456789.123456, -7.9995, 1 / 3, 100000.0]
    for k in range(25):
        n, m = rng.randint(1, 4), rng.randint(1, 3)
        obj = [rng.choice([1.0, -2.0, 0.5, 0.0, 12.25, -0.125, 1000.0, 1e-5]) for _ in range(n)]
        rows = [([rng.choice([0.0, 1.0, -1.5, 2.0, 0.1, 1e7]) for _ in range(n)], rng.choice(["<=", ">=", "="]),
                 rng.choice([4.0, 0.0, -3.5, 1e6, 0.3])) for _ in range(m)]
        signs = [rng.choice(["+", "-", "urs", "bin", "int"]) for _ in range(n)]
        ptype = rng.choice(["max", "min"])

        def tok(v):
            return repr(float(v)) if v != int(v) or abs(v) >= 1e15 else str(int(v))
        text = "\n".join([ptype + " " + " ".join(tok(v) for v in obj)] +
                         [" ".join(tok(v) for v in co) + f" {rel} {tok(rhs)}" for co, rel, rhs in rows] + [" ".join(signs)])
        snaps = [rng.choice(["", "one line", "ends with a newline\n", "a\r\nb\r\n", "tab\there", "unicode ≤ ∞"]) for _ in range(rng.randint(0, 3))]
        z = rng.choice(specials)
        x = [rng.choice(specials) for _ in range(rng.randint(0, n))]
        # the reference, executed
        it.files.clear()
        it.files["model.txt"] = text
        p = it.new("InputFileParser")
        it.call(p, "ReadInputFile", "model.txt")
        args = (it.get(p, "ProblemType"), it.get(p, "ObjectiveCoefficients"), it.get(p, "Constraints"), it.get(p, "SignRestrictions"))
        canon = it.call_static("CanonicalFormConverter", "CanonicalFormForFile", *args)
        it.call_static("OutputFileWrite", "WriteFullResults", "out.txt", "Solver X", *args, CsList(list(snaps), "string"),
                       float(z), to_list(x))
        it.call_static("OutputFileWrite", "WriteSnapshotsOnly", "out.txt", "Solver Y", CsList(list(snaps), "string"), float(z),
                       to_list(x) if x else None)
        want = it.files["out.txt"]
        # the native writer
        mdl = Model.parse_text(text)
        assert mdl.canonical_form() == canon, text
        path = str(tmp_path / f"o{k}.txt")
        L.io.OutputFileWrite.WriteFullResults(path, "Solver X", mdl, snaps, z, x, append=False, timestamp=it.now)
        L.io.OutputFileWrite.WriteSnapshotsOnly(path, "Solver Y", snaps, z, x if x else None, append=True, timestamp=it.now)
        assert open(path, "rb").read() == want.encode("utf-8"), (text, snaps, z, x)


@pytest.mark.parametrize("seed", range(3))
def test_sensitivity_reoptimisation(run, seed):
    """SensitivityAnalyzer(finalTableau, ...).AddNewConstraintNonInteractive and ResolveAll after an RHS change,
    executed, against orc_sens_*"""
    rng = random.Random(7000 + seed)
    done = 0
    while done < 12:
        n, m = rng.randint(2, 6), rng.randint(2, 6)
        obj = [float(rng.randint(1, 9)) for _ in range(n)]
        cons = [([float(rng.randint(0, 9)) for _ in range(n)], "<=", float(rng.randint(5, 40))) for _ in range(m)]
        T0, b0 = O.primal_build(obj, cons, True)
        opt = O.primal_solve(T0, b0)
        if opt["status"] != O.OPTIMAL:
            continue
        done += 1
        tech = [float(rng.randint(-2, 5)) for _ in range(n)] + [0.0] * m
        rhs = float(rng.randint(-10, 20))
        g = run.sensitivity(obj, cons, tech, rhs)
        T = opt["T"]
        basis = O.sens_rebuild_basis(T)
        assert basis.tolist() == g["basis_rebuilt"]
        x = O.primal_extract(T, n)
        ax = 0.0
        for j in range(n):
            ax += tech[j] * x[j]
        T1, _ = O.sens_add_constraint(T, basis, tech, rhs - ax)
        res = O.sens_resolve(T1, O.sens_rebuild_basis(T1))
        assert same(res["T"], unmat(g["tableau_after"])) and res["basis"].tolist() == g["basis_after"]
        assert (res["status"] == O.OPTIMAL) == (g["exception"] is None)
        if g["exception"] is None:
            assert float(res["T"][0, -1]).hex() == g["z_after"]
            assert same(O.sens_solution(res["T"]), [float.fromhex(h) for h in g["x_after"]])
        g2 = run.sensitivity_rhs(obj, cons, rng.randint(1, m), float(rng.choice([-30, -12, -5, 4, 9])))
        Tb = unmat(g2["tableau_before_resolve"])
        res = O.sens_resolve(Tb, O.sens_rebuild_basis(Tb))
        assert same(res["T"], unmat(g2["tableau_after"])) and res["basis"].tolist() == g2["basis_after"]
        assert (res["status"] == O.OPTIMAL) == (g2["exception"] is None)
