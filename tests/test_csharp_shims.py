"""CPU tests that EXECUTE the C# P/Invoke shims under csharp/ (the drop-in a maintainer adds to the reference,
INTEGRATION.md) and hold them to what the reference's own classes returned (tests/golden/reference_run.json).

The image has no .NET toolchain, so the shims were source-only until oracle/csharp/ could run them: the interpreter
executes the shim classes, `[DllImport] static extern` calls go through ctypes with the CLR's default marshalling
(oracle/csharp/pinvoke.py).  Here host-only entry points (parser, formatters) hit the REAL liblprb200.so and the compute
entry points an oracle-backed test double (tests/lprb200_double.py); tests/test_csharp_shims_gpu.py runs the same cases
against the real library for everything.  Also checked statically: every DllImport declaration against the prototype in
include/lprb200.h, and that the shims release every native handle they create.
"""
import os
import re

import pytest

import csharp_shim_cases as S
from lprb200_double import OracleBackedDouble

GOLD = S.GOLD
ROOT = S.ROOT


@pytest.fixture(scope="module")
def shims():
    double = OracleBackedDouble(S.real_library())
    sh = S.Shims(double)
    yield sh
    assert double.handles == {}, "the shims leaked native handles"


def test_every_shim_file_parses_and_declares_the_reference_class_surface():
    from csharp.csparse import parse_source
    classes = {}
    for f in S.SHIMS:
        unit = parse_source(open(f, encoding="utf-8-sig").read(), os.path.basename(f))
        for c in unit[2]:
            classes[c[1]] = c
    for name in ("PrimalSimplexSolver", "PrimalSimplexSolver2", "DualSimplexSolver", "RevisedPrimalSimplexSolver",
                 "CuttingPlaneSolver", "BranchBoundSimplexSolver", "BranchAndBoundAdapter", "SensitivityAnalyzer",
                 "InputFileParser", "TableIterationFormater", "KnapsackBranchBoundSimplex", "KnapsackBranchBoundSolver"):
        assert name in classes, name
    nested = {m[1] for m in classes["BranchBoundSimplexSolver"][4] if m[0] == "class"}
    assert {"DualSimplexSolverBB", "BranchAndBound", "TeeTextWriter"} <= nested


C_KIND = {"int": "int", "int64_t": "long", "double": "double", "uint64_t": "ulong", "float": "float"}


def _header_prototypes():
    text = open(os.path.join(ROOT, "include", "lprb200.h")).read()
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    text = re.sub(r"//[^\n]*", " ", text)
    protos = {}
    for ret, name, params in re.findall(r"\b(int|int64_t|const char\*|void)\s+(lpr_\w+)\s*\(([^;{]*?)\)\s*;", text, flags=re.S):
        plist = [] if params.strip() in ("", "void") else [" ".join(p.split()) for p in params.split(",")]
        kinds = []
        for p in plist:
            if "*" in p:
                kinds.append("pointer")
            else:
                kinds.append(C_KIND.get(p.split()[-2] if len(p.split()) > 1 else p.split()[0], p))
        protos[name] = (ret, kinds)
    return protos


def test_dllimport_declarations_match_the_c_header():
    from csharp.csparse import parse_source
    unit = parse_source(open(os.path.join(ROOT, "csharp", "LprNative.cs"), encoding="utf-8-sig").read(), "LprNative.cs")
    lpr = [c for c in unit[2] if c[1] == "Lpr"][0]
    protos = _header_prototypes()
    seen = 0
    for m in lpr[4]:
        if m[0] != "method" or "extern" not in m[1]:
            continue
        name, params, ret = m[3], m[4], m[2]
        assert name in protos, f"{name} is not declared in include/lprb200.h"
        cret, kinds = protos[name]
        assert len(params) == len(kinds), f"{name}: {len(params)} managed parameters, {len(kinds)} in the header"
        assert (ret[1], cret) in (("int", "int"), ("IntPtr", "const char*"), ("long", "int64_t")), (name, ret[1], cret)
        for (ty, pname, _d, mod), kind in zip(params, kinds):
            managed_pointer = bool(ty[3]) or mod in ("out", "ref") or ty[1] in ("IntPtr", "string")
            if kind == "pointer":
                assert managed_pointer, f"{name}.{pname}: the header takes a pointer, the shim passes {ty[1]} by value"
            else:
                assert not managed_pointer and ty[1] == kind, f"{name}.{pname}: header {kind}, shim {ty[1]}"
        seen += 1
    assert seen >= 70


@pytest.mark.parametrize("i", range(len(GOLD["primal"])))
def test_primal_simplex_solver_shim(shims, i):
    bad, _ = shims.primal(GOLD["primal"][i])
    assert bad == []


@pytest.mark.parametrize("i", range(len(GOLD["primal2"])))
def test_primal_simplex_solver2_shim(shims, i):
    assert shims.primal2(GOLD["primal2"][i]) == []


@pytest.mark.parametrize("i", range(len(GOLD["dual"])))
def test_dual_simplex_solver_shim(shims, i):
    assert shims.dual(GOLD["dual"][i]) == []


@pytest.mark.parametrize("i", range(len(GOLD["accessors"])))
def test_solve_on_demand_accessors_of_the_shims(shims, i):
    """GetRows / GetObjectiveRow / GetConstraintRows / AnyNegativeRhs and their InvalidOperationException messages"""
    assert shims.accessors(GOLD["accessors"][i]) == []


@pytest.mark.parametrize("i", range(len(GOLD["cutting_plane"])))
def test_cutting_plane_solver_shim(shims, i):
    assert shims.cutting_plane(GOLD["cutting_plane"][i]) == []


@pytest.mark.parametrize("i", range(len(GOLD["revised"])))
def test_revised_primal_simplex_solver_shim(shims, i):
    assert shims.revised(GOLD["revised"][i]) == []


@pytest.mark.parametrize("i", range(len(GOLD["bb"])))
def test_branch_and_bound_adapter_shim(shims, i):
    assert shims.bb(GOLD["bb"][i]) == []


@pytest.mark.parametrize("i", range(len(GOLD["bb_formulate"])))
def test_dual_simplex_solver_bb_shim(shims, i):
    assert shims.bb_formulate(GOLD["bb_formulate"][i]) == []


@pytest.mark.parametrize("i", range(len(GOLD["bb_parts"])))
def test_branch_and_bound_members_of_the_shim(shims, i):
    assert shims.bb_parts(GOLD["bb_parts"][i]) == []


@pytest.mark.parametrize("i", range(len(GOLD["sensitivity"])))
def test_sensitivity_analyzer_shim(shims, i):
    assert shims.sensitivity(GOLD["sensitivity"][i]) == []


def test_input_file_parser_and_formatter_shims_on_the_real_library(shims, tmp_path):
    """no double involved: lpr_model_* and lpr_fmt_table are host code"""
    before = len(shims.it.native.calls)
    for k, g in enumerate(GOLD["parser"]):
        d = tmp_path / f"p{k}"
        d.mkdir()
        assert shims.parser(g, d) == [], g["text"]
    assert shims.format_table() == []
    assert len(shims.it.native.calls) == before          # nothing was answered by the double
    assert "lpr_model_parse_file" in shims.it.native.real.calls and "lpr_fmt_table" in shims.it.native.real.calls


def test_run_branch_and_bound_entry_shim(shims):
    """BranchAndBound.RunBranchAndBound (:1253-1298) on model A == the reference's menu path result (Appendix C3)"""
    from csharp.csrun import CsList, to_list
    it = shims.it
    bb = it.new("BranchAndBound")
    it.console.clear()
    it.call(bb, "RunBranchAndBound", to_list([2, 3, 3, 5, 2, 4]), CsList([to_list([11, 8, 6, 14, 10, 10, 40, 0])], None), False)
    assert "Best integer solution: z = 15" in it.console_text()
    # AddConstraint on a rounded tableau == the oracle's (the double answers, so this checks the shim's marshalling)
    g = GOLD["bb"][0]
    root = S.unmat(g["root_tableau"])
    it.call(bb, "SetNumVars", g["n_vars"])
    first = g["nodes"][0]
    row = [1.0 if j == first["branch_var"] else 0.0 for j in range(g["n_vars"])] + [0.0, 0.0]
    out = it.call(bb, "AddConstraint", CsList([to_list(row)], None), CsList([to_list(r) for r in root.tolist()], None))
    import oracle_lib as O
    assert S.bits_equal(S.from_cs(out.vals[0]), O.bb_add_constraint(O.bb_round(root), g["n_vars"], first["branch_var"], 0.0, 0))


def test_knapsack_shims(shims):
    import oracle_lib as O
    w, v, cap = O.gen_knapsack(384, 40)
    best, dp, chosen = shims.knapsack(cap, w.tolist(), v.tolist())
    assert best == dp                                   # the reference's own check, Program.cs:467-470
    assert sum(v[i] for i in chosen) == best and sum(w[i] for i in chosen) <= cap


def test_compute_shims_fail_loudly_without_a_gpu():
    """with the REAL library and no CUDA device every compute shim throws InvalidOperationException("liblprb200: ...")
    from its first native call: the managed arguments were marshalled, the library answered, nothing fell back"""
    import lpr_381_group_v22_b200 as L
    if L.device_count() > 0:
        pytest.skip("a CUDA device is present")
    from csharp import CsException
    from csharp.csrun import CsList, to_array, to_list
    sh = S.Shims(S.real_library())
    it = sh.it
    g = GOLD["primal"][0]
    T = S.unmat(GOLD["primal2"][0]["tableau"]).tolist()
    attempts = {
        "PrimalSimplexSolver": lambda: it.new("PrimalSimplexSolver", to_list(g["objective"]), sh.constraints(g["constraints"]), True),
        "PrimalSimplexSolver2": lambda: it.call(it.new("PrimalSimplexSolver2", to_array(T[0]), sh.rows(T)), "Solve"),
        "DualSimplexSolver": lambda: it.call(it.new("DualSimplexSolver"), "Solve", to_array(T[0]), sh.rows(T), 10, False),
        "CuttingPlaneSolver": lambda: it.call(it.new("CuttingPlaneSolver"), "CuttingPlaneSolution", to_array(T[0]), sh.rows(T)),
        "RevisedPrimalSimplexSolver": lambda: it.new("RevisedPrimalSimplexSolver", to_list([1, 2]),
                                                     sh.constraints([([1, 1], "<=", 4)]), False),
        "DualSimplexSolverBB": lambda: it.call(it.new("DualSimplexSolverBB"), "FormulateTableau", to_list([1, 2]),
                                               CsList([to_list([1, 1, 4, 0])], None)),
        "KnapsackBranchBoundSolver": lambda: it.call_static("KnapsackBranchBoundSolver", "Solve", 10, to_array([3, 4], "int"),
                                                            to_array([5, 6], "int")),
    }
    for name, attempt in attempts.items():
        with pytest.raises(CsException) as err:
            attempt()
        assert err.value.tname == "InvalidOperationException" and err.value.message.startswith("liblprb200: "), name


@pytest.mark.skipif(not os.path.isdir("/root/reference/LPR_381_Group_V22"), reason="needs the reference's own .cs files (build container only)")
def test_reference_files_run_unchanged_on_top_of_the_shims():
    """the deployment INTEGRATION.md describes: the solver classes are replaced by the shims, everything else stays the
    reference's file -- here OutputFileWrite.cs, CanonicalFormConverter.cs, NumFormat and Program.AddUpperBoundConstraints,
    unmodified, consume the shim classes' members and must write the very files the all-reference run wrote"""
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_reference_run", os.path.join(S.HERE, "golden", "make_reference_run.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    double = OracleBackedDouble(S.real_library())
    sh = S.Shims(double)                                  # shim classes are registered first and win name lookups
    for f in ("IO/OutputFileWrite.cs", "Utilities/CanonicalFormConverter.cs", "Simplex/RevisedPrimalSimplexSolver.cs", "Program.cs"):
        sh.it.load_file(gen.REF + f)
    assert sh.it.find_class("PrimalSimplexSolver").methods["Solve"][0][5] is not None
    assert "lpr_tab_create_primal" in open(os.path.join(ROOT, "csharp", "PrimalSimplexSolverShim.cs")).read()
    run = gen.Runner.__new__(gen.Runner)
    run.it = sh.it
    import tempfile
    for g in GOLD["output"]:
        with tempfile.TemporaryDirectory() as d:        # the shim parser is native code: it reads a real file
            path = os.path.join(d, "model.txt")
            with open(path, "w", encoding="utf-8", newline="") as f:
                f.write(g["text"])
            rec = _output_through_shims(run, path, g)
        for key in ("rows_added", "canonical_form", "snapshots", "final_z", "x", "file_after_full_results", "file_after_append"):
            assert rec[key] == g[key], key
    assert "lpr_tab_step" in double.calls and "lpr_model_parse_file" in double.real.calls


def _output_through_shims(run, model_path, g):
    """Runner.output of the generator with the model read from a real file (the shim parser is native code)"""
    it = run.it
    orig_call = it.call

    def call(obj, method, *args, **named):
        if method == "ReadInputFile":
            args = (model_path,)
        return orig_call(obj, method, *args, **named)
    it.call = call
    try:
        return run.output(g["text"], g["solver"], g["add_upper_bound_rows"])
    finally:
        it.call = orig_call


@pytest.mark.skipif(not os.path.isdir("/root/reference/LPR_381_Group_V22"), reason="needs the reference's own Program.cs (build container only)")
def test_reference_program_main_runs_on_the_shims():
    """THE drop-in test: the reference's unmodified Program.cs (Main, the whole menu), OutputFileWrite.cs,
    CanonicalFormConverter.cs and NumFormat, with every solver class replaced by its shim from csharp/, driven by a
    scripted keyboard -- against the console transcript and data/output_results.txt of the all-reference run of the same
    session (tests/golden/reference_run.json "program").  Option 5 (knapsack) only exists with the shims: upstream
    Program.cs:444,468 refer to classes the reference does not contain."""
    import importlib.util
    import re as _re
    import tempfile
    spec = importlib.util.spec_from_file_location("make_reference_run", os.path.join(S.HERE, "golden", "make_reference_run.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    with tempfile.TemporaryDirectory() as d:
        double = OracleBackedDouble(S.real_library())
        sh = S.Shims(double)
        for f in ("IO/OutputFileWrite.cs", "Utilities/CanonicalFormConverter.cs", "Simplex/RevisedPrimalSimplexSolver.cs", "Program.cs"):
            sh.it.load_file(gen.REF + f)
        it = sh.it
        # the shim parser is native code and reads a real file: a POSIX project tree under d
        os.makedirs(os.path.join(d, "proj", "bin", "Debug"))
        os.makedirs(os.path.join(d, "proj", "data"))
        it.base_dir = os.path.join(d, "proj", "bin", "Debug") + "/"
        it.path_sep = "/"

        def session(model, keys):
            with open(os.path.join(d, "proj", "data", "model.txt"), "w", encoding="utf-8", newline="") as f:
                f.write(model)
            it.console.clear()
            it.files.clear()
            it.out_writer = None
            it.stdin = ["model.txt"] + list(keys)
            it.call_static("Program", "Main", None)
            return it.console_text(), it.files.get("data/output_results.txt")

        prog = {tuple(g["keys"]): g for g in GOLD["program"]}
        for keys in (("1", "13", "7"), ("2", "7"), ("4", "7"), ("9", "7")):
            g = prog[keys]
            console, out = session(g["model"], keys)
            assert console == g["console"], keys
            assert out == g["output_file"], keys
        # option 3: the reference prints every tableau of every node into the tee'd snapshot; the shim prints the result
        g = prog[("3", "7")]
        console, out = session(g["model"], ("3", "7"))
        block = _re.search(r"=== Branch & Bound Result ===.*?(?=-{20})", console, _re.S).group(0)
        assert block in g["console_tail"] and "Z* = 15" in block
        assert out.startswith("﻿") and out.endswith(g["output_file_tail"][g["output_file_tail"].index("=== Final Results ==="):])
        # option 5 does not compile upstream; with the shims it runs and the reference's own check passes
        console, _ = session(prog[("1", "13", "7")]["model"], ("5", "7"))
        assert "Branch & Bound Best Value Z* = 15" in console and "Dynamic Programming Result: 15" in console
        assert "Results Match: True" in console
        # Program.cs never disposes its solvers (the reference's classes are not IDisposable): the three solver objects and
        # the SensitivityAnalyzer of these sessions keep their handles until the SafeHandle finalizers run under a real GC
        assert len(double.handles) == 4


def test_untraced_paths_and_small_members_of_the_shims(shims):
    """what the recorded cases do not reach: the one-call paths above TraceMaxElements (no snapshots), DualPrices,
    PrepareInput, RoundVector / RoundNumber / IsInteger / RoundAllTableaux"""
    from csharp.csrun import CsList, from_cs, to_array, to_list
    import numpy as np
    it = shims.it
    p2 = it.find_class("PrimalSimplexSolver2")
    rv = it.find_class("RevisedPrimalSimplexSolver")
    it.ensure_static(p2)
    it.ensure_static(rv)
    saved = p2.statics["TraceMaxElements"], rv.statics["TraceMaxElements"]
    p2.statics["TraceMaxElements"] = rv.statics["TraceMaxElements"] = 0
    try:
        for g in GOLD["primal2"]:
            T = S.unmat(g["tableau"]).tolist()
            s = it.new("PrimalSimplexSolver2", to_array(T[0]), shims.rows(T))
            assert bool(it.call(s, "Solve", g["max_iters"], g["print_steps"])) == g["returned"]
            rows = it.call(s, "GetRows", False)
            assert S.bits_equal([from_cs(rows.vals[0])] + from_cs(rows.vals[1]), S.unmat(g["final_tableau"]))
            assert from_cs(it.get(s, "IterationSnapshots")) == []
        for g in GOLD["revised"]:
            if g["exception"] is not None:
                continue
            cons = [(g["A"][i], g["relations"][i], g["b"][i]) for i in range(len(g["A"]))]
            s = it.new("RevisedPrimalSimplexSolver", to_list(g["c"]), shims.constraints(cons), g["is_min"])
            it.call(s, "Solve")
            assert float(it.get(s, "FinalZ")).hex() == g["final_z"] and from_cs(it.get(s, "BasicVariables")) == g["basis"]
            assert from_cs(it.get(s, "IterationSnapshots")) == []
            y = np.array(from_cs(it.get(s, "DualPrices")))
            binv, basis = S.unmat(g["binv"]), g["basis"]
            c = np.array(g["c"], dtype=float) * (-1.0 if g["is_min"] else 1.0)
            cb = np.array([c[b] if b < len(c) else 0.0 for b in basis])
            assert np.allclose(y, cb @ binv, rtol=1e-12, atol=1e-12)          # y = c_B B^-1
            it.call(s, "Dispose")
    finally:
        p2.statics["TraceMaxElements"], rv.statics["TraceMaxElements"] = saved
    bbo = it.new("BranchAndBound")
    assert it.call(bbo, "RoundNumber", 1.00005) == 1.0001 and it.call(bbo, "IsInteger", 2.0000004) and not it.call(bbo, "IsInteger", 2.001)
    r = GOLD["rounding"]
    vals = [float.fromhex(h) for h in r["values"]]
    assert [float(it.call(bbo, "RoundNumber", v)).hex() for v in vals] == r["round4"]
    assert [bool(it.call(bbo, "IsInteger", v)) for v in vals] == r["is_integer"]
    import oracle_lib as O
    vec = [0.12345, 2.5, -0.00005, 7.00005]
    assert from_cs(it.call(bbo, "RoundVector", to_list(vec))) == [O.lib().orc_net_round4(v) for v in vec]
    g = GOLD["bb_parts"][0]
    both = it.call(bbo, "RoundAllTableaux", CsList([CsList([to_list(r) for r in S.unmat(g["tableau"]).tolist()], None)] * 2, None))
    assert all(S.bits_equal(from_cs(t), S.unmat(g["rounded"])) for t in both.items)
    f = GOLD["bb_formulate"][0]
    prep = it.call(it.new("DualSimplexSolverBB"), "PrepareInput", to_list(f["objective"]),
                   CsList([to_list(r) for r in f["rows"]], None), f["is_min"])
    assert S.bits_equal(from_cs(prep.vals[0]), S.unmat(f["tableau"])) and prep.vals[1] == f["is_min"]
    assert prep.vals[2] == sum(1 for r in f["rows"] if r[-1] in (1, 2)) and prep.vals[3] == len(f["rows"]) - prep.vals[2]
    assert prep.vals[4] == len(f["objective"])
