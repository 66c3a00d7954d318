"""Probe knapsack instance families on the GPU (node counts, nodes/s) -- used to choose the 'cfg4-hard' family of
bench.py: python tools/knap_probe.py [n_gpus]"""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lpr_381_group_v22_b200 import _native as N  # noqa: E402
from lpr_381_group_v22_b200.bench_workloads import u01  # noqa: E402


def family(kind, seed, n, R=1000):
    idx = np.arange(n, dtype=np.uint64)
    w = 1.0 + np.floor(R * u01(seed, 0, idx))
    if kind == "weak":
        v = np.maximum(1.0, w + np.floor(R / 5 * u01(seed, 1, idx)) - R / 10)
    elif kind == "strong":
        v = w + R / 10
    elif kind == "almost_strong":
        v = w + R / 10 + np.floor(R / 250 * u01(seed, 1, idx)) - R / 500
    elif kind == "uncorr":
        v = 1.0 + np.floor(R * u01(seed, 1, idx))
    elif kind == "subset":
        v = w.copy()
    else:
        raise ValueError(kind)
    return w, v, float(np.floor(w.sum() / 2.0))


def run(kind, seed, n, n_gpus=1, cap_nodes=200_000_000, R=1000):
    w, v, cap = family(kind, seed, n, R)
    best, nodes, st = C.c_double(), C.c_int64(), C.c_int()
    ch = np.zeros(n, dtype=np.uint8)
    stats = N.MgpuStats()
    t = time.perf_counter()
    N.check(N.lib().lpr_knap_solve_mgpu(n_gpus, None, cap, n, N.pd(w), N.pd(v), cap_nodes, -1, 2e-3, C.byref(best),
                                        ch.ctypes.data_as(N.bp), C.byref(nodes), C.byref(st), C.byref(stats)))
    dt = time.perf_counter() - t
    d = stats.as_dict()
    print(dict(kind=kind, seed=seed, n=n, R=R, n_gpus=n_gpus, best=best.value, nodes=nodes.value, status=st.value,
               wall=round(dt, 3), loop_s=round(d["seconds"], 4), mnodes_per_s=round(nodes.value / max(d["seconds"], 1e-9) / 1e6, 2),
               rounds=d["rounds"], steals=d["steals"], moved=d["nodes_moved"], open_left=d["open_left"],
               per_gpu=d["nodes_per_gpu"]), flush=True)


if __name__ == "__main__":
    ng = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    which = sys.argv[2] if len(sys.argv) > 2 else "first"
    os.environ.setdefault("LPR_KNAP_POOL_MB", "32768")
    first = [("weak", 384, 10000, 1000), ("weak", 384, 10000, 1000), ("weak", 385, 100000, 1000),
             ("weak", 386, 10000, 100000), ("uncorr", 387, 10000, 1000), ("almost_strong", 388, 2000, 1000),
             ("almost_strong", 388, 10000, 1000), ("strong", 389, 200, 1000), ("strong", 389, 1000, 1000),
             ("subset", 390, 1000, 100000)]
    fourth = [("strong", 395, 115, 1000), ("strong", 395, 120, 1000), ("strong", 397, 110, 1000), ("strong", 398, 110, 1000), ("strong", 399, 112, 1000)]
    third = [("strong", 389, 110, 1000), ("strong", 389, 120, 1000), ("strong", 395, 110, 1000), ("almost_strong", 388, 250, 1000), ("almost_strong", 388, 300, 1000), ("subset", 392, 3000, 1000000), ("subset", 396, 5000, 1000000)]
    second = [("subset", 390, 2000, 100000), ("subset", 390, 4000, 100000), ("subset", 391, 1000, 1000000),
              ("subset", 392, 3000, 1000000), ("almost_strong", 388, 100, 1000), ("almost_strong", 388, 200, 1000),
              ("almost_strong", 388, 400, 1000), ("strong", 389, 50, 1000), ("strong", 389, 100, 1000),
              ("weak", 393, 10000, 100), ("weak", 394, 30000, 300)]
    cap = int(os.environ.get("KNAP_PROBE_CAP", "1500000000"))
    lists = {"first": first, "second": second, "third": third, "cfg4": first[:1], "hard": [("strong", 395, 110, 1000)], "fourth": fourth}
    for kind, seed, n, R in lists[which]:
        try:
            run(kind, seed, n, ng, cap_nodes=cap, R=R)
        except Exception as ex:
            print("FAIL", kind, seed, n, repr(ex), flush=True)
