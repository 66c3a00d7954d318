"""Device-side timing of the revised simplex iteration (development tool)."""
import argparse, ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lpr_381_group_v22_b200 as L
from lpr_381_group_v22_b200 import _native as N
ap = argparse.ArgumentParser()
ap.add_argument("--m", type=int, default=8192); ap.add_argument("--n", type=int, default=16384)
ap.add_argument("--iters", type=int, default=64); ap.add_argument("--reps", type=int, default=2)
a = ap.parse_args()
tag = " ".join(f"{k}={v}" for k, v in os.environ.items() if k.startswith("LPR_"))
for rep in range(a.reps):
    h = N.vp(); N.check(N.lib().lpr_rev_create_dense_lp(0, 384, a.m, a.n, C.byref(h)))
    st = C.c_int(); nit = C.c_int64(); ms = C.c_float()
    N.check(N.lib().lpr_rev_solve(h, a.iters, 0, C.byref(st), C.byref(nit), None, 0))
    N.check(N.lib().lpr_rev_last_solve_ms(h, C.byref(ms)))
    byts = 24.0 * a.m * a.m + 8.0 * a.m * a.n
    print(f"[{tag}] rep{rep} m={a.m} n={a.n} iters={nit.value} status={st.value} {ms.value:.2f} ms "
          f"{ms.value*1e3/max(1,nit.value):.1f} us/iter {nit.value/ms.value*1e3:.0f} it/s "
          f"{byts*nit.value/ms.value/1e6:.0f} GB/s(24m2+8mn)", flush=True)
    N.lib().lpr_rev_destroy(h)
