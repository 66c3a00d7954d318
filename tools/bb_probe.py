"""B&B simplex pool throughput on the cfg5 root for several (batch, pools per device) settings -- development tool:
python tools/bb_probe.py [n_gpus]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lpr_381_group_v22_b200 import bench_workloads as W  # noqa: E402

ng = int(sys.argv[1]) if len(sys.argv) > 1 else 1
os.environ.setdefault("LPR_BB_PREALLOC_MB", "65536")
os.environ.setdefault("LPR_BB_MAX_DEPTH", "192")
A, b, c = W.gen_dense_ip(385, 512, 1024)
final, lp, ms = W.lp_relaxation(A, b, c, 0)
bytes_rc = 16.0 * 514 * 1538
configs = ((64, 1), (64, 2), (128, 1), (128, 2), (64, 3), (256, 1), (32, 4))
if os.environ.get("BB_PROBE_CONFIGS"):  # e.g. "128:1,256:1"
    configs = tuple(tuple(int(x) for x in c.split(":")) for c in os.environ["BB_PROBE_CONFIGS"].split(","))
for batch, rpg in configs:
    os.environ["LPR_BB_BATCH"] = str(batch)
    os.environ["LPR_MG_RANKS_PER_GPU"] = str(rpg)
    if os.environ.get("LPR_BB_PROFILE"):
        print(f"--- batch {batch} rpg {rpg}", flush=True)
    r = W.bb_mgpu(final, 1024, ng, max_rounds=12, slice_seconds=0.01)
    gbs = bytes_rc * (2.0 + r["pivots_per_node"]) * r["nodes_per_s"] / 1e9
    print(dict(batch=batch, rpg=rpg, nodes=r["nodes"], nodes_per_s=round(r["nodes_per_s"]), ppn=round(r["pivots_per_node"], 3),
               gbs=round(gbs), frac=round(gbs / 6553.9 / ng, 3), per_gpu=r["nodes_per_gpu"], steals=r["steals"]), flush=True)
