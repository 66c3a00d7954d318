"""The C++ host mirror (host/lpr_solvers.hpp) compiles and links against liblprb200.so everywhere; with a GPU
its fixture program must pass (same Appendix C answers as the Python mirror)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "lpr_381_group_v22_b200")
EXE = os.path.join(ROOT, "tests", "cpp", "test_host")


def build_exe():
    src = os.path.join(ROOT, "tests", "cpp", "test_host.cpp")
    cmd = ["g++", "-std=c++17", "-O1", "-ffp-contract=off", src, "-o", EXE, "-L", PKG, "-llprb200",
           f"-Wl,-rpath,{PKG}"]
    subprocess.check_call(cmd)
    return EXE


def test_cpp_host_mirror_compiles_and_links():
    exe = build_exe()
    assert os.path.exists(exe)


@pytest.mark.gpu
def test_cpp_host_mirror_fixtures():
    exe = build_exe()
    res = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "all checks passed" in res.stdout


def test_cpp_host_mirror_io_against_the_executed_reference(tmp_path):
    """the host-code half of host/lpr_solvers.hpp (IO::InputFileParser, AddUpperBoundConstraints, IO::OutputFileWrite,
    Utilities::TableIterationFormater / NumFormat) on CPU, against what the reference's own classes returned for the
    same inputs (tests/golden/reference_run.json "output" and "parser": outputs of the executed reference)"""
    import json
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_run.json")))
    src = os.path.join(ROOT, "tests", "cpp", "test_host_io.cpp")
    exe = os.path.join(ROOT, "tests", "cpp", "test_host_io")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-ffp-contract=off", src, "-o", exe, "-L", PKG, "-llprb200",
                           f"-Wl,-rpath,{PKG}"])

    def run(text, add, out, solver, stamp, z_hex, x_hex, snaps):
        model = tmp_path / "model.txt"
        model.write_bytes(text.encode("utf-8"))
        files = []
        for k, s in enumerate(snaps):
            f = tmp_path / f"snap{k}.txt"
            f.write_bytes(s.encode("utf-8"))
            files.append(str(f))
        cmd = [exe, str(model), "1" if add else "0", str(out), solver, stamp, z_hex, str(len(x_hex)), *x_hex,
               str(len(files)), *files]
        r = subprocess.run(cmd, capture_output=True, timeout=120)
        assert r.returncode == 0, r.stderr.decode()
        return r.stdout.decode("utf-8")

    for k, g in enumerate(gold["output"]):
        out = tmp_path / f"out{k}.txt"
        text = run(g["text"], g["add_upper_bound_rows"], out, g["solver"], g["timestamp"], g["final_z"], g["x"], g["snapshots"])
        added = [ln for ln in text.splitlines() if ln.startswith("ADDED")]
        assert len(added) == len(g["rows_added"])
        for ln, (co, rel, rhs) in zip(added, g["rows_added"]):
            body, r, b = ln[len("ADDED"):].split("|")
            assert [float.fromhex(t) for t in body.split()] == [float.fromhex(h) for h in co]
            assert r.strip() == rel and float.fromhex(b.strip()) == float.fromhex(rhs)
        assert out.read_bytes() == g["file_after_append"].encode("utf-8")
        assert "N3 2.001 0 1E+15" in text
        table = text.split("TABLE_BEGIN\n")[1].split("TABLE_END")[0]
        assert "Z\t0.000\t0.000\t1.000\t-1.000\t\r\n1\t0.500\t-0.500\t0.001\t-0.001\t\r\n2\t0.002\t2.001\t1234.568\t0.000\t" in table
    for g in gold["parser"]:
        text = run(g["text"], False, tmp_path / "unused.txt", "s", "2025-01-01 00:00:00", "0x0p+0", [], [])
        if g["exception"] is not None:
            assert text.startswith("EXCEPTION") and g["exception"] in text
            continue
        assert text.splitlines()[0] == "MESSAGE " + g["console"].strip()
        if g["problem_type"] is None:
            assert "TYPE" not in text
            continue
        assert f"TYPE {g['problem_type']}\n" in text
        obj = [ln for ln in text.splitlines() if ln.startswith("OBJ")][0].split()[1:]
        assert [float.fromhex(t) for t in obj] == [float.fromhex(h) for h in g["objective"]]
        signs = [ln for ln in text.splitlines() if ln.startswith("SIGNS")][0]
        assert signs == "SIGNS" + "".join(f" [{s}]" for s in g["signs"])
