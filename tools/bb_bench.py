"""B&B throughput (cfg5 dense IP / cfg4 knapsack) on 1..N GPUs (development tool; bench.py reports the
same numbers under "bb" / "knapsack").  Launch with torchrun for N > 1."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--what", default="bb", choices=["bb", "knap"])
    ap.add_argument("--m", type=int, default=512)
    ap.add_argument("--n", type=int, default=1024)
    ap.add_argument("--items", type=int, default=10000)
    ap.add_argument("--seed", type=int, default=385)
    ap.add_argument("--max-nodes", type=int, default=4000)
    ap.add_argument("--chunk", type=int, default=256)
    ap.add_argument("--slice-ms", type=float, default=0.0, help="time-sliced rounds (B&B simplex pool)")
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch, torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import lpr_381_group_v22_b200 as L
    from lpr_381_group_v22_b200.bench_workloads import run_bb_cfg5, run_knap_cfg4
    if a.what == "bb":
        r = run_bb_cfg5(a.m, a.n, a.seed, local, dist, a.max_nodes, a.chunk, a.slice_ms)
    else:
        r = run_knap_cfg4(a.items, a.seed, local, dist, a.max_nodes, a.chunk)
    if rank == 0:
        print(json.dumps(r), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
