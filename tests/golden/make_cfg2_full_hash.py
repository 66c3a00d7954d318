"""Generate tests/golden/cfg2_full_solve.json: SHA-256 digests of the FULL BASELINE cfg2 solve
(m=4096, n=8192, seed 383; 14 866 pivots to optimality) as computed by the CPU oracle
(oracle/lpr_oracle.cpp, orc_primal_solve = PrimalSimplexSolver.cs:102-211 restated), row loop of
Pivot split over all host threads (bit-identical to one thread: element-wise work, no reductions).

    python tests/golden/make_cfg2_full_hash.py            # several minutes on 8 cores

The GPU test (tests/test_tableau_gpu.py::test_full_size_solve_cfg2_matches_oracle_digest) hashes the
pivot log, the basis and the final tableau of the device solve and compares them with these digests,
so the whole 14 866-pivot sequence is pinned to the oracle, not just a window of it.
"""
import hashlib
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

import numpy as np  # noqa: E402
import oracle_lib as O  # noqa: E402

M, N, SEED = 4096, 8192, 383


def digests(log, basis, T):
    """the exact byte strings the GPU test hashes: int32 (row, col) pairs, int32 basis, float64 row-major tableau"""
    return dict(
        pivot_log_sha256=hashlib.sha256(np.ascontiguousarray(log, dtype=np.int32).tobytes()).hexdigest(),
        basis_sha256=hashlib.sha256(np.ascontiguousarray(basis, dtype=np.int32).tobytes()).hexdigest(),
        final_tableau_sha256=hashlib.sha256(np.ascontiguousarray(T, dtype=np.float64).tobytes()).hexdigest())


def main():
    threads = os.cpu_count() or 1
    A, b, c = O.gen_dense_lp(SEED, M, N)
    T0, b0 = O.primal_build(list(c), [(A[i], "<=", b[i]) for i in range(M)])
    del A
    t = time.perf_counter()
    r = O.primal_solve(T0, b0, threads=threads, log_cap=1 << 15)
    dt = time.perf_counter() - t
    assert r["status"] == O.OPTIMAL, r["status"]
    out = dict(m=M, n=N, seed=SEED, n_pivots=int(r["n_pivots"]), status="OPTIMAL",
               z_hex=float(r["T"][0, -1]).hex(), first_pivots=r["log"][:8].tolist(), last_pivots=r["log"][-8:].tolist(),
               oracle_seconds=round(dt, 1), oracle_threads=threads,
               generator="tests/golden/make_cfg2_full_hash.py (CPU oracle, C++ restatement of PrimalSimplexSolver.cs)")
    out.update(digests(r["log"], r["basis"], r["T"]))
    with open(os.path.join(HERE, "cfg2_full_solve.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
