"""Regenerates tests/golden/oracle_vectors.json from the CPU oracle.

The reference ships no golden vectors and cannot be executed here (no .NET), so the
committed vectors are the oracle's outputs on the reference's own fixtures (data/TextFile.txt = model A,
README model = model B, Program.cs:433-435 knapsack) plus seeded synthetic cases.  They are pinned to
the independently derived known answers of SURVEY.md Appendix C by tests/test_oracle_golden.py.
Run:  python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib as O  # noqa: E402


def hexs(a):
    return [float(x).hex() for x in np.asarray(a, dtype=float).ravel()]


def cli_rows(n):
    rows = []
    for i in range(n):
        co = [0.0] * (n + 3)
        co[i] = 1.0
        co[n + 1] = 1.0
        rows.append((co, "<=", 1.0))
    return rows


def main():
    out = {}
    models = {"A": ([2, 3, 3, 5, 2, 4], [([11, 8, 6, 14, 10, 10], "<=", 40)]),
              "B": ([2, 3, 4], [([1, 2, 3], "<=", 10), ([3, 2, 1], ">=", 15)])}
    for name, (obj, cons) in models.items():
        for cli in (True, False):
            c = cons + (cli_rows(len(obj)) if cli else [])
            T, b = O.primal_build(obj, c)
            r = O.primal_solve(T, b)
            out[f"primal_{name}_{'cli' if cli else 'raw'}"] = dict(
                shape=list(T.shape), status=r["status"], log=r["log"].tolist(), basis=r["basis"].tolist(),
                z=float(r["T"][0, -1]).hex(), x=hexs(O.primal_extract(r["T"], len(obj))), final=hexs(r["T"]))
    T, b = O.primal_build(*[models["A"][0], models["A"][1] + cli_rows(6)])
    Tf = O.primal_solve(T, b)["T"]
    for prune in (False, True):
        r = O.bb_solve(Tf, 6, prune=prune, max_nodes=20)
        out[f"bb_A_prune{int(prune)}"] = dict(nodes=r["nodes"], x=hexs(r["x"]), z=float(r["z"]).hex(),
                                              node_log=r["node_log"].tolist(), node_z=hexs(r["node_z"]),
                                              pivots=r["pivots"])
    row, cut = O.gomory_cut(Tf)
    cp = O.cutting_plane(Tf)
    out["cut_A"] = dict(row=row, cut=hexs(cut), status=cp["status"], log=cp["log"].tolist(), final=hexs(cp["T"]),
                        shape=list(cp["T"].shape))
    A = np.vstack([[11, 8, 6, 14, 10, 10], np.eye(6)])
    rr = O.rev_solve(A, [40] + [1] * 6, models["A"][0], want_binv=True)
    out["rev_A"] = dict(log=rr["log"].tolist(), basis=rr["basis"].tolist(), z=float(rr["z"]).hex(), x=hexs(rr["x"]),
                        y=hexs(rr["y"]))
    rr = O.rev_solve(np.array([[1, 2, 3], [3, 2, 1.0]]), [10, 15], [2, 3, 4])
    out["rev_B"] = dict(log=rr["log"].tolist(), basis=rr["basis"].tolist(), z=float(rr["z"]).hex(), x=hexs(rr["x"]))
    dp, ch = O.knap_dp(40, [11, 8, 6, 14, 10, 10], [2, 3, 3, 5, 2, 4])
    kb = O.knap_bb(40, [11, 8, 6, 14, 10, 10], [2, 3, 3, 5, 2, 4])
    out["knap_program_cs"] = dict(dp=dp, dp_chosen=ch.tolist(), bb=kb["best"], bb_chosen=kb["chosen"].tolist(),
                                  nodes=kb["nodes"])
    # seeded synthetic cases (generator of SURVEY 8d)
    for seed, m, n in ((381, 12, 24), (382, 20, 31)):
        A, b, c = O.gen_dense_lp(seed, m, n)
        T0, b0 = O.primal_build(list(c), [(A[i], "<=", b[i]) for i in range(m)])
        r = O.primal_solve(T0, b0)
        out[f"dense_lp_{seed}_{m}x{n}"] = dict(u0=float(O.lib().orc_u01(seed, 0)).hex(), a00=float(A[0, 0]).hex(),
                                                log=r["log"].tolist(), z=float(r["T"][0, -1]).hex(),
                                                basis=r["basis"].tolist())
    w, v, cap = O.gen_knapsack(385, 64)
    kb = O.knap_bb(cap, w, v)
    out["knap_385_64"] = dict(cap=cap, w=w.tolist(), v=v.tolist(), best=kb["best"], chosen=kb["chosen"].tolist(),
                              nodes=kb["nodes"], dp=O.knap_dp(int(cap), w.astype(int), v.astype(int))[0])
    with open(os.path.join(HERE, "oracle_vectors.json"), "w") as f:
        json.dump(out, f, indent=0, sort_keys=True)
    print("wrote", len(out), "cases")


if __name__ == "__main__":
    main()
