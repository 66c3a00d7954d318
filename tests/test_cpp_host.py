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
