"""The cooperative chain kernel of the B&B batches (k_bb_chains, LPR_BB_PERSIST=1) against the sequential oracle.  The
switch is read once per process, so the check runs in a subprocess."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CODE = r'''
import sys
sys.path.insert(0, "tests")
import numpy as np
import oracle_lib as O
import lpr_381_group_v22_b200 as L
from test_bb_gpu import MID, binary_ip_final
for seed, m, n, div in MID[:2]:
    Tf = binary_ip_final(seed, m, n, div)
    ref = O.bb_solve(Tf, n, prune=True, max_nodes=-1, log_cap=1 << 16)
    r = L.solve_bb_mgpu(Tf, n, True, n_gpus=1, slice_seconds=1e-3)
    assert r["status"] == L.OPTIMAL and r["z"] == ref["z"], (r["z"], ref["z"])
    assert np.array_equal(np.asarray(r["x"]).view(np.uint64), ref["x"].view(np.uint64))
    seq = O.bb_solve(Tf, n, prune=False, max_nodes=40)
    bb = L.BranchBoundSimplexSolver.BranchAndBound()
    bb.SetNumVars(n)
    bb.ExecuteBranchAndBound([Tf], False, max_nodes=40)
    assert bb.LastRun["node_log"].tolist() == seq["node_log"].tolist()
    assert np.array_equal(bb.LastRun["node_z"].view(np.uint64), seq["node_z"].view(np.uint64))
print("OK")
'''


def test_cooperative_chain_kernel_matches_oracle():
    env = dict(os.environ, LPR_BB_PERSIST="1")
    out = subprocess.run([sys.executable, "-c", CODE], cwd=ROOT, env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
