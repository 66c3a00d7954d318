"""Quick device-side timing of the fused primal pivot loop (development tool, not the bench)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lpr_381_group_v22_b200 as L

ap = argparse.ArgumentParser()
ap.add_argument("--m", type=int, default=4096)
ap.add_argument("--n", type=int, default=8192)
ap.add_argument("--pivots", type=int, default=256)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--seed", type=int, default=383)
ap.add_argument("--fused", type=int, default=1)
ap.add_argument("--time-sweeps", type=int, default=0)
a = ap.parse_args()
R, C = a.m + 1, a.n + a.m + 1
byts = 16.0 * R * C
tag = " ".join(f"{k}={v}" for k, v in os.environ.items() if k.startswith("LPR_"))
for rep in range(a.reps):
    with L.DeviceTableau.dense_lp(a.seed, a.m, a.n) as t:
        l0 = L.launch_count()
        r = t.solve(L.RULE_PRIMAL, max_pivots=a.pivots, log_cap=0, fused=bool(a.fused), time_sweeps=bool(a.time_sweeps))
        ms = t.last_solve_ms
        npv = r["n_pivots"]
        print(f"[{tag}] rep{rep} {R}x{C} pivots={npv} status={r['status']} {ms:.3f} ms "
              f"{ms * 1e3 / max(1, npv):.2f} us/pivot {npv / ms * 1e3:.0f} pivots/s "
              f"{byts * npv / ms / 1e6:.0f} GB/s launches={L.launch_count() - l0}"
              + (f" sweep_kernel={t.last_sweep_us:.1f} us" if a.time_sweeps else ""), flush=True)
