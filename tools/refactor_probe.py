"""Time the B^-1 refactorisation at BASELINE cfg3 (m = 8192): refresh (mode 1) and full blocked Gauss-Jordan (mode 2),
twice each (the first call of a handle allocates the workspace).  Development tool."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lpr_381_group_v22_b200 import _native as N  # noqa: E402

m, n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192, int(sys.argv[2]) if len(sys.argv) > 2 else 16384
lib = N.lib()
h = N.vp()
N.check(lib.lpr_rev_create_dense_lp(0, 384, m, n, C.byref(h)))
st, nit = C.c_int(), C.c_int64()
N.check(lib.lpr_rev_solve(h, 200, 0, C.byref(st), C.byref(nit), None, 0))
for mode in (1, 1, 2, 2, 0):
    N.check(lib.lpr_rev_refactor_ex(h, mode))
    ms, res, fl, path, after = C.c_float(), C.c_double(), C.c_double(), C.c_int(), C.c_double()
    N.check(lib.lpr_rev_last_refactor_ms(h, C.byref(ms)))
    N.check(lib.lpr_rev_last_refactor_info(h, C.byref(res), C.byref(fl)))
    N.check(lib.lpr_rev_last_refactor_path(h, C.byref(path), C.byref(after)))
    print(dict(mode=mode, path=path.value, ms=round(ms.value, 2), tflops=round(fl.value / ms.value / 1e9, 2),
               residual_before=res.value, residual_after=after.value), flush=True)
lib.lpr_rev_destroy(h)
