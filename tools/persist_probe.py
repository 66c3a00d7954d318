"""Time the generic-rule pivot loops on the cfg5 relaxation tableau (513 x 1537): cutting plane (32 cuts) and a dual
simplex chain, persistent cooperative kernel (default) vs the two-kernel path (LPR_TAB_PERSIST=0).  Development tool."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lpr_381_group_v22_b200 as L  # noqa: E402
from lpr_381_group_v22_b200 import bench_workloads as W  # noqa: E402

A, b, c = W.gen_dense_ip(385, 512, 1024)
final, lp, ms = W.lp_relaxation(A, b, c, 0)
print("relaxation", final.shape, lp["n_pivots"], "pivots", round(ms, 2), "ms", flush=True)
for persist in ("1", "0", "1", "0"):
    os.environ["LPR_TAB_PERSIST"] = persist
    with L.DeviceTableau.from_host(final, row_cap=final.shape[0] + 40) as t:
        t0 = time.perf_counter()
        r = t.cutting_plane(max_cuts=32)
        dt = time.perf_counter() - t0
        piv = int(r["log"][:, 2].sum() + r["log"][:, 3].sum()) + int(r["n_cuts"])
        print(dict(persist=persist, cuts=int(r["n_cuts"]), pivots=piv, host_ms=round(dt * 1e3, 3),
                   device_ms=round(t.last_solve_ms, 3), us_per_pivot=round(dt * 1e6 / piv, 2)), flush=True)
