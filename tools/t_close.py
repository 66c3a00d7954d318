import time, sys, os
sys.path.insert(0, os.getcwd())
import lpr_381_group_v22_b200 as L
for rep in range(4):
    t0 = time.perf_counter()
    t = L.DeviceTableau.dense_lp(383, 4096, 8192)
    t1 = time.perf_counter()
    r = t.solve(L.RULE_PRIMAL, max_pivots=512, log_cap=0)
    t2 = time.perf_counter()
    z = t.objective()
    t3 = time.perf_counter()
    t.close()
    t4 = time.perf_counter()
    print(f"rep{rep} create {1e3*(t1-t0):.1f} solve {1e3*(t2-t1):.1f} (dev {0:.1f}) obj {1e3*(t3-t2):.1f} close {1e3*(t4-t3):.1f} ms", flush=True)
