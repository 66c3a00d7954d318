import os, sys
sys.path.insert(0, os.getcwd())
from lpr_381_group_v22_b200 import bench_workloads as W
import lpr_381_group_v22_b200 as L
A, b, c = W.gen_dense_ip(385, 512, 1024)
final, lp, ms = W.lp_relaxation(A, b, c, 0)
def bb(tag):
    r = W.bb_mgpu(final, 1024, 1, max_rounds=12, slice_seconds=0.01)
    print(tag, dict(nodes=r["nodes"], nps=round(r["nodes_per_s"]), run=r.get("run_seconds_per_gpu")), flush=True)
bb("before cuts")
bb("before cuts 2")
with L.DeviceTableau.from_host(final, row_cap=final.shape[0] + 40) as t:
    r = t.cutting_plane(max_cuts=32)
    print("cuts", t.last_solve_ms, flush=True)
bb("after cuts")
bb("after cuts 2")
