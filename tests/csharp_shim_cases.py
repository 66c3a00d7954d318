"""The C# P/Invoke shims under csharp/ EXECUTED (oracle/csharp interpreter, P/Invoke through ctypes) and compared with
what the reference's own classes returned on the same inputs (tests/golden/reference_run.json).

Shared by tests/test_csharp_shims.py (CPU: host-only entry points hit the real liblprb200.so, compute entry points the
oracle-backed test double tests/lprb200_double.py) and tests/test_csharp_shims_gpu.py (B200: everything hits the real
library).  The shims are the drop-in a maintainer adds to the reference; the image has no .NET, so
this is the only way they ever run here.
"""
import glob
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

from csharp import CsException, Interpreter  # noqa: E402
from csharp.csrun import CsList, from_cs, to_array, to_array2d, to_list  # noqa: E402
from csharp.pinvoke import NativeLibrary  # noqa: E402

GOLD = json.load(open(os.path.join(HERE, "golden", "reference_run.json")))
SHIMS = sorted(glob.glob(os.path.join(ROOT, "csharp", "*.cs")))


def unmat(m):
    return np.array([float.fromhex(h) for h in m["hex"]], dtype=np.float64).reshape(m["shape"])


def unhex(v):
    return np.array([float.fromhex(h) for h in v], dtype=np.float64)


def bits_equal(a, b):
    a = np.ascontiguousarray(a, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    return a.shape == b.shape and np.array_equal(a.view(np.uint64), b.view(np.uint64))


def sha(text):
    return hashlib.sha256(text.encode("utf-8")).hexdigest()


def real_library():
    """its own CDLL object: pinvoke sets restype / argtypes per call and must not touch the product binding's prototypes"""
    import ctypes
    from lpr_381_group_v22_b200 import _native as N
    N.lib()                                   # fails loudly when the library is missing
    return NativeLibrary(ctypes.CDLL(N.LIB_PATH))


class Shims:
    """one interpreter with every file of csharp/ loaded and `native` bound"""

    def __init__(self, native):
        self.it = Interpreter()
        self.it.native = native
        for f in SHIMS:
            self.it.load_file(f)

    def constraints(self, cons):
        it = self.it
        return CsList([it.new("Constraint", to_list(co), rel, float(rhs)) for co, rel, rhs in cons], None)

    @staticmethod
    def rows(T):
        return CsList([to_array(r) for r in T[1:]], None)

    # ---- each returns the list of mismatches (empty = parity)
    def primal(self, g, exact=True):
        it, bad = self.it, []
        it.console.clear()
        s = it.new("PrimalSimplexSolver", to_list(g["objective"]), self.constraints(g["constraints"]), g["is_max"])
        if not bits_equal(from_cs(it.call(s, "GetFinalTableau")), unmat(g["initial_tableau"])):
            bad.append("initial tableau")
        it.call(s, "Solve")
        if not bits_equal(from_cs(it.get(s, "FinalTableau")), unmat(g["final_tableau"])):
            bad.append("FinalTableau")
        if from_cs(it.get(s, "BasicVariables")) != g["basis"]:
            bad.append("BasicVariables")
        if float(it.get(s, "FinalZ")).hex() != g["final_z"]:
            bad.append("FinalZ")
        x = it.get(s, "SolutionVector")
        if g["x"] is None:
            if x is not None:
                bad.append("SolutionVector should stay null")
        elif x is None or not bits_equal(from_cs(x), unhex(g["x"])):
            bad.append("SolutionVector")
        if sha(it.get(s, "FinalTable")) != g["final_table_sha256"]:
            bad.append("FinalTable")
        snaps = from_cs(it.get(s, "IterationSnapshots"))
        if "snapshots" in g and snaps != g["snapshots"]:
            bad.append("IterationSnapshots")
        elif len(snaps) != g["n_snapshots"] or sha("".join(snaps)) != g["snapshots_sha256"]:
            bad.append("IterationSnapshots (digest)")
        it.call(s, "Dispose")
        return bad, s

    def primal2(self, g):
        it, bad = self.it, []
        T = unmat(g["tableau"]).tolist()
        s = it.new("PrimalSimplexSolver2", to_array(T[0]), self.rows(T))
        ok = it.call(s, "Solve", g["max_iters"], g["print_steps"])
        if bool(ok) != g["returned"]:
            bad.append(f"Solve returned {ok}")
        rows = it.call(s, "GetRows", False)
        if not bits_equal([from_cs(rows.vals[0])] + from_cs(rows.vals[1]), unmat(g["final_tableau"])):
            bad.append("GetRows(false)")
        if g["returned"] and float(it.get(s, "FinalZ")).hex() != g["final_z"]:
            bad.append("FinalZ")
        snaps = from_cs(it.get(s, "IterationSnapshots"))
        if len(snaps) != g["n_snapshots"] or sha("".join(snaps)) != g["snapshots_sha256"]:
            bad.append("IterationSnapshots")
        return bad

    def dual(self, g):
        it, bad = self.it, []
        T = unmat(g["tableau"]).tolist()
        obj, rows = to_array(T[0]), self.rows(T)
        s = it.new("DualSimplexSolver")
        err = None
        try:
            ok = it.call(s, "Solve", obj, rows, g["max_iters"], g["print_steps"])
        except CsException as e:
            ok, err = None, e.tname
        if err != g["exception"]:
            bad.append(f"exception {err}")
        if err is None and bool(ok) != g["returned"]:
            bad.append(f"Solve returned {ok}")
        if not bits_equal([from_cs(obj)] + from_cs(rows), unmat(g["final_tableau"])):
            bad.append("rows after Solve")
        return bad

    def accessors(self, g):
        it, bad = self.it, []
        T = unmat(g["tableau"]).tolist()
        if g["dual"]:
            obj, rows = to_array(T[0]), self.rows(T)
            s = it.new("DualSimplexSolver")
            if bool(it.call_static("DualSimplexSolver", "AnyNegativeRhs", rows)) != g["any_negative_rhs"]:
                bad.append("AnyNegativeRhs")
            err = None
            try:
                r = it.call(s, "GetRows", obj, rows)
                if not bits_equal([from_cs(r.vals[0])] + from_cs(r.vals[1]), unmat(g["rows"])):
                    bad.append("GetRows")
                o2, c2 = it.call(s, "GetObjectiveRow", obj, rows), it.call(s, "GetConstraintRows", obj, rows)
                if not bits_equal([from_cs(o2)] + from_cs(c2), unmat(g["rows_again"])):
                    bad.append("GetObjectiveRow / GetConstraintRows")
            except CsException as e:
                err = [e.tname, e.message]
            if err != g["exception"]:
                bad.append(f"exception {err}")
            if not bits_equal([from_cs(obj)] + from_cs(rows), unmat(g["inputs_after"])):
                bad.append("caller's rows after GetRows")
            it.console.clear()
            it.call_static("DualSimplexSolver", "PrintTableau", obj, rows, len(T[0]) - len(T), None)
            it.call_static("DualSimplexSolver", "PrintTableau", obj, rows, 1, "Custom title")
            if it.console_text() != g["print_tableau"]:
                bad.append("PrintTableau")
            return bad
        s = it.new("PrimalSimplexSolver2", to_array(T[0]), self.rows(T))
        err = None
        try:
            r = it.call(s, "GetRows")
            if not bits_equal([from_cs(r.vals[0])] + from_cs(r.vals[1]), unmat(g["rows"])):
                bad.append("GetRows")
            if not bits_equal(from_cs(it.call(s, "GetObjectiveRow")), unhex(g["objective_row"])):
                bad.append("GetObjectiveRow")
            if len(it.call(s, "GetConstraintRows", False).items) != g["n_constraint_rows"]:
                bad.append("GetConstraintRows")
            if float(it.get(s, "FinalZ")).hex() != g["final_z"]:
                bad.append("FinalZ")
        except CsException as e:
            err = [e.tname, e.message]
        if err != g["exception"]:
            bad.append(f"exception {err}")
        return bad

    def cutting_plane(self, g):
        it, bad = self.it, []
        T = unmat(g["tableau"]).tolist()
        obj, rows = to_array(T[0]), self.rows(T)
        s = it.new("CuttingPlaneSolver")
        it.call(s, "CuttingPlaneSolution", obj, rows)
        if not bits_equal([from_cs(obj)] + from_cs(rows), unmat(g["final_tableau"])):
            bad.append("rows after CuttingPlaneSolution")
        log = from_cs(it.get(s, "CutLog"))
        k = len(g["cut_pivots_1based"])
        if [[len(T) + j, c[1] + 1] for j, c in enumerate(log[:k])] != g["cut_pivots_1based"]:
            bad.append("CutLog")
        return bad

    def revised(self, g, text=True):
        it, bad = self.it, []
        cons = [(g["A"][i], g["relations"][i], g["b"][i]) for i in range(len(g["A"]))]
        s = it.new("RevisedPrimalSimplexSolver", to_list(g["c"]), self.constraints(cons), g["is_min"])
        err = None
        try:
            it.call(s, "Solve")
        except CsException as e:
            err = e.message
        if err != g["exception"]:
            bad.append(f"exception {err!r}")
        if from_cs(it.get(s, "BasicVariables")) != g["basis"]:
            bad.append("BasicVariables")
        if err is None:
            z, x = float.fromhex(g["final_z"]), unhex(g["x"])
            if abs(it.get(s, "FinalZ") - z) > 1e-9 * max(1.0, abs(z)):
                bad.append("FinalZ")
            if not np.allclose(from_cs(it.get(s, "SolutionVector")), x, rtol=1e-9, atol=1e-9 * max(1.0, float(np.abs(x).max()))):
                bad.append("SolutionVector")
        snaps = from_cs(it.get(s, "IterationSnapshots"))      # also when Solve threw: the iterations completed before it
        if len(snaps) != g["n_snapshots"]:
            bad.append(f"{len(snaps)} snapshots")
        elif text and sha("".join(snaps)) != g["snapshots_sha256"]:
            bad.append("IterationSnapshots text")
        it.call(s, "Dispose")
        return bad

    def bb(self, g):
        it, bad = self.it, []
        pg = {"objective": g["objective"], "constraints": g["constraints"], "is_max": True}
        it.console.clear()
        p = it.new("PrimalSimplexSolver", to_list(pg["objective"]), self.constraints(pg["constraints"]), True)
        it.call(p, "Solve")
        if not bits_equal(from_cs(it.get(p, "FinalTableau")), unmat(g["root_tableau"])):
            bad.append("root tableau")
        res = it.call_static("BranchAndBoundAdapter", "SolveFromPrimal", p, g["enable_pruning"], g["is_min"])
        x, z = from_cs(res.vals[0]), res.vals[1]
        if not bits_equal(np.array(x, dtype=np.float64), unhex(g["x"])):
            bad.append("x")
        if float(z).hex() != g["z"]:
            bad.append("z")
        it.call(p, "Dispose")
        return bad

    def bb_formulate(self, g):
        it, bad = self.it, []
        s = it.new("DualSimplexSolverBB")
        rows = CsList([to_list(r) for r in g["rows"]], None)
        T = it.call(s, "FormulateTableau", to_list(g["objective"]), rows)
        if not bits_equal(from_cs(T), unmat(g["tableau"])):
            bad.append("FormulateTableau")
        if [[float(v).hex() for v in r] for r in from_cs(rows)] != g["rows_after"]:
            bad.append("caller's rows after FormulateTableau")
        res = it.call(it.new("DualSimplexSolverBB"), "DoDualSimplex", to_list(g["objective"]),
                      CsList([to_list(r) for r in g["rows"]], None), g["is_min"])
        tabs, _dec, opt, pc, pr, _hdr = res.vals
        if (opt is None) != (g["optimal_value"] is None):
            bad.append("optimalValue null-ness")
        elif opt is not None:
            if float(opt).hex() != g["optimal_value"]:
                bad.append("optimalValue")
            if not bits_equal(from_cs(tabs.items[-1]), unmat(g["final_tableau"])):
                bad.append("final tableau")
            if from_cs(pc) != g["pivot_cols"] or from_cs(pr) != g["pivot_rows"]:
                bad.append("pivot lists")
        return bad

    def bb_parts(self, g):
        """BranchAndBound.RoundTableau / AddConstraint and DualSimplexSolverBB.PerformDualPivot / PerformPrimalPivot.  The
        two pivot members are one step of the native DoDualSimplex state machine (the C ABI has no single-pivot entry for
        this rule): compared in the states DoDualSimplex calls them from -- a dual pivot while a RHS is negative, a
        primal one when none is and none becomes negative (such a pivot is dropped, :392-400) -- and up to the
        -0.0 -> 0.0 clean-up DoDualSimplex applies to every tableau (:307-313)"""
        it, bad = self.it, []
        T = unmat(g["tableau"]).tolist()

        def ll(rows):
            return CsList([to_list(r) for r in rows], None)
        bbo = it.new("BranchAndBound")
        it.call(bbo, "SetNumVars", g["n_vars"])
        if not bits_equal(from_cs(it.call(bbo, "RoundTableau", ll(T))), unmat(g["rounded"])):
            bad.append("RoundTableau")
        row = [1.0 if j == g["var"] else 0.0 for j in range(g["n_vars"])] + [float(g["bound"]), float(g["type"])]
        out = it.call(bbo, "AddConstraint", CsList([to_list(row)], None), ll(T))
        if not bits_equal(from_cs(out.vals[0]), unmat(g["add_constraint"])):
            bad.append("AddConstraint")
        s = it.new("DualSimplexSolverBB")
        Tr = unmat(g["rounded"])
        if g["dual_pivot"] is not None:
            dp = it.call(s, "PerformDualPivot", ll(Tr.tolist()))
            if dp.vals[1] is None or not bits_equal(np.array(from_cs(dp.vals[0])) + 0.0, unmat(g["dual_pivot"]) + 0.0):
                bad.append("PerformDualPivot")
        if g["primal_pivot"] is not None and not (Tr[:, -1] < 0).any() and not (unmat(g["primal_pivot"])[:, -1] < 0).any():
            pp = it.call(s, "PerformPrimalPivot", ll(Tr.tolist()), False)
            if pp.vals[1] is None or not bits_equal(np.array(from_cs(pp.vals[0])) + 0.0, unmat(g["primal_pivot"]) + 0.0):
                bad.append("PerformPrimalPivot")
        return bad

    def sensitivity(self, g):
        it, bad = self.it, []
        p = it.new("PrimalSimplexSolver", to_list(g["objective"]), self.constraints(g["constraints"]), True)
        it.call(p, "Solve")
        s = it.new("SensitivityAnalyzer", it.get(p, "FinalTableau"), it.get(p, "SolutionVector"), it.get(p, "FinalZ"),
                   it.get(p, "BasicVariables"))
        err = None
        try:
            it.call(s, "AddNewConstraintNonInteractive", to_array(g["tech"]), float(g["rhs"]))
        except CsException as e:
            err = e.message
        if err != g["exception"]:
            bad.append(f"exception {err!r}")
        if not bits_equal(from_cs(it.get(s, "CurrentTableau")), unmat(g["tableau_after"])):
            bad.append("CurrentTableau")
        if float(it.get(s, "CurrentZ")).hex() != g["z_after"]:
            bad.append("CurrentZ")
        it.call(s, "Dispose")
        it.call(p, "Dispose")
        return bad

    def parser(self, g, tmp_path):
        it, bad = self.it, []
        path = os.path.join(str(tmp_path), "model.txt")
        with open(path, "w", encoding="utf-8", newline="") as f:
            f.write(g["text"])
        it.console.clear()
        p = it.new("InputFileParser")
        err = None
        try:
            it.call(p, "ReadInputFile", path)
        except CsException as e:
            err = e.tname
        if err != g["exception"]:
            bad.append(f"exception {err}")
            return bad
        if err is not None:
            return bad
        if it.console_text() != g["console"]:
            bad.append("console message")
        if it.get(p, "ProblemType") != g["problem_type"]:
            bad.append("ProblemType")
        if not bits_equal(from_cs(it.get(p, "ObjectiveCoefficients")), unhex(g["objective"])):
            bad.append("ObjectiveCoefficients")
        cons = it.get(p, "Constraints").items
        if len(cons) != len(g["constraints"]):
            bad.append("Constraints.Count")
        for c, (co, rel, rhs) in zip(cons, g["constraints"]):
            if not bits_equal(from_cs(it.get(c, "Coefficients")), unhex(co)) or it.get(c, "Relation") != rel \
                    or float(it.get(c, "RHS")).hex() != rhs:
                bad.append("a constraint")
        if from_cs(it.get(p, "SignRestrictions")) != g["signs"]:
            bad.append("SignRestrictions")
        return bad

    def format_table(self):
        g = GOLD["format"]
        vals = unhex(g["values"])
        k = len(vals) - len(vals) % 4
        text = self.it.call_static("TableIterationFormater", "Format", to_array2d(vals[:k].reshape(-1, 4).tolist()), 2, "T")
        return [] if text == g["table"] else ["TableIterationFormater.Format"]

    def knapsack(self, capacity, weights, values):
        it = self.it
        s = it.new("KnapsackBranchBoundSimplex", int(capacity), to_array(weights), to_array(values))
        best = it.call(s, "Solve")
        items = it.call(s, "GetSelectedItemsOriginal")
        dp = it.call_static("KnapsackBranchBoundSolver", "Solve", int(capacity), to_array([int(w) for w in weights], "int"),
                            to_array([int(v) for v in values], "int"))
        chosen = [it.get(i, "Id") for i in items.items]
        return best, dp, chosen


def run_all_against_real_library():
    """every compute case against the real liblprb200.so; {category: {index: [mismatches]}} -- run in a child process by
    tests/test_csharp_shims_gpu.py so that nothing the interop layer does can take the pytest process down"""
    sh = Shims(real_library())
    out = {}
    plan = [("primal", lambda g: sh.primal(g)[0]), ("primal2", sh.primal2), ("dual", sh.dual), ("accessors", sh.accessors),
            ("cutting_plane", sh.cutting_plane), ("revised", lambda g: sh.revised(g, text=False)), ("bb", sh.bb),
            ("bb_formulate", sh.bb_formulate), ("bb_parts", sh.bb_parts), ("sensitivity", sh.sensitivity)]
    for name, fn in plan:
        res = {}
        for i, g in enumerate(GOLD[name]):
            try:
                res[str(i)] = fn(g)
            except Exception as e:      # an interpreter / interop error is a finding too, not a crash of the run
                res[str(i)] = [f"{type(e).__name__}: {e}"[:300]]
        out[name] = res
    from lpr_381_group_v22_b200.bench_workloads import gen_knapsack
    w, v, cap = gen_knapsack(384, 40)
    try:
        best, dp, chosen = sh.knapsack(cap, w.tolist(), v.tolist())
        ok = best == dp and sum(v[i] for i in chosen) == best and sum(w[i] for i in chosen) <= cap
        out["knapsack"] = {"0": [] if ok else [f"B&B {best} vs DP {dp}"]}
    except Exception as e:
        out["knapsack"] = {"0": [f"{type(e).__name__}: {e}"[:300]]}
    out["native_calls"] = sorted(set(sh.it.native.calls))
    return out


if __name__ == "__main__":
    sys.path.insert(0, ROOT)
    print("SHIM_RESULTS " + json.dumps(run_all_against_real_library()))
