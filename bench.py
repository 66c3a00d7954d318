#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native dense simplex pivot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (BASELINE.json configs[1]): synthetic dense LP m=4096 n=8192 fp64, primal tableau simplex
(reference layout 4097 x 12289, SURVEY.md 8d generator, seed 383).  One "step" = one full Solve()
from the slack basis to optimality.  Metric = simplex pivots/s (whole job, all N GPUs).

  value     : pivots/s with the tableau already resident in HBM (generated on the device), timed
              with CUDA events on the library's stream, max over ranks.
  e2e       : the same solve through the C ABI with HOST (pinned) model arrays: H2D of A/b/c,
              device-side tableau build, Solve(), D2H of the final tableau + basis + x + z.
  roofline  : HBM roofline of the dominant kernel (the rank-1 sweep) and of the whole pivot.
  cpu_baseline : the CPU oracle (C++ restatement of the reference's C# loops, single thread like the
              reference) on a bounded sample of the same workload, on this box's host cores.

With N > 1 (torchrun, one process per GPU) the tableau path runs as independent replicas (it does
not shard: DESIGN.md "multi-GPU"); B&B nodes/s at N GPUs is reported under "bb".
`--impl reference` times the CPU oracle with all host threads (the reference arm of the contract).
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

M, NV, SEED = 4096, 8192, 383
R, CC = M + 1, NV + M + 1
BYTES_PER_PIVOT = 16.0 * R * CC + 8.0 * (CC - 1) + 16.0 * (R - 1) + 8.0 * (R + CC)
BYTES_SWEEP = 16.0 * R * CC + 8.0 * (R + CC)  # tableau read+write once, staged column and row read
NOMINAL_HBM_GBS = 8000.0


def env_flag(name):
    """LPR_BENCH_SKIP_*: set and not "0" / empty"""
    return os.environ.get(name, "") not in ("", "0")


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def stop(self, t0, t1):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.lines:
            if ts < t0 or ts > t1 + 0.2:
                continue
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        if sm:
            out["sm_mhz"] = statistics.median(sm)
            out["sm_max_mhz"] = max(mx)
            out["samples"] = len(sm)
        out["reasons"] = sorted(reasons)
        return out


def dist_setup(n_gpus):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist_mod
        torch.cuda.set_device(local)
        dist_mod.init_process_group("nccl", device_id=torch.device("cuda", local))
        dist = dist_mod
    return rank, world, local, dist


def barrier_max(dist, local, value):
    if dist is None:
        return value
    import torch
    t = torch.tensor([value], dtype=torch.float64, device=f"cuda:{local}")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier(dist, local):
    if dist is not None:
        import torch
        dist.barrier(device_ids=[local])
        torch.cuda.synchronize(local)


def pinned(shape):
    import numpy as np
    try:
        import torch
        return torch.empty(shape, dtype=torch.float64, pin_memory=True).numpy()
    except Exception:
        return np.empty(shape)


# ---------------------------------------------------------------------------------------------
def run_reference_arm(args, rank, world):
    """CPU oracle (C++ restatement of PrimalSimplexSolver.cs:102-211; the image has no .NET) with all
    host threads, on a bounded sample of the cfg2 workload.  Rank 0 only."""
    if rank != 0:
        return
    import oracle_lib as O
    threads = os.cpu_count() or 1
    sample = int(os.environ.get("LPR_REF_SAMPLE_PIVOTS", "16"))
    A, b, c = O.gen_dense_lp(SEED, M, NV)
    T0, b0 = O.primal_build(list(c), [(A[i], "<=", b[i]) for i in range(M)])
    del A
    total = args.warmup + args.steps
    t_step = []
    T, basis = T0, b0
    for s in range(total):
        t = time.perf_counter()
        r = O.primal_solve(T, basis, max_pivots=sample, threads=threads, log_cap=sample)
        dt = time.perf_counter() - t
        T, basis = r["T"], r["basis"]  # continue the same solve: every pivot costs the same
        if s >= args.warmup:
            t_step.append(dt)
    sec = sum(t_step)
    val = sample * len(t_step) / sec
    line = {
        "impl": "reference", "metric": "simplex pivots/s (dense 4096x8192 fp64 primal tableau)",
        "value": val, "unit": "pivots/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * sec / len(t_step), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": "cfg2 dense LP m=4096 n=8192 primal tableau simplex, tableau 4097x12289",
                   "step": f"{sample} consecutive pivots of the same solve (bounded sample)", "seed": SEED},
        "cpu_baseline": {"value": val, "unit": "pivots/s", "cores": threads, "kind": "port",
                         "sample": f"{sample} pivots per step x {len(t_step)} steps; C++ restatement of the "
                                   "reference C# loops (no .NET toolchain in the image; pinned bit for bit to the "
                                   "reference's own sources executed by oracle/csharp), row loop of Pivot "
                                   f"split over {threads} threads (bit-identical)"},
        "e2e": {"value": val, "unit": "pivots/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "tableau_gbs": 16.0 * R * CC * val / 1e9,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--max-pivots", type=int, default=int(os.environ.get("LPR_BENCH_MAX_PIVOTS", "-1")),
                    help="cap pivots per step (debug); default: solve to optimality")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    rank, world, local, dist = dist_setup(args.gpus)
    import numpy as np
    import lpr_381_group_v22_b200 as L
    from lpr_381_group_v22_b200 import _native as N

    if L.device_count() < 1:
        raise SystemExit("bench.py needs a CUDA device: lpr_381_group_v22_b200 has no CPU fallback")
    dev = local
    peaks, peak_kind = measured_peaks()
    lib = N.lib()

    def solve_resident(max_pivots, blocked=True):
        """one step with the inputs resident in HBM: returns (pivots, device ms, status)"""
        t = L.DeviceTableau.dense_lp(SEED, M, NV, device=dev)
        r = t.solve(L.RULE_PRIMAL, max_pivots=max_pivots, log_cap=0, blocked=blocked)
        ms = t.last_solve_ms
        z = t.objective()
        t.close()
        return r["n_pivots"], ms, r["status"], z

    # -------- value: inputs resident in HBM --------------------------------------------------
    for _ in range(args.warmup):
        solve_resident(args.max_pivots)
    barrier(dist, local)
    sampler = ClockSampler(dev)
    sampler.start()
    time.sleep(0.25)
    l0 = L.launch_count()
    t0 = time.perf_counter()
    piv_total, ms_total, status, zval = 0, 0.0, None, None
    for _ in range(args.steps):
        p, ms, status, zval = solve_resident(args.max_pivots)
        piv_total += p
        ms_total += ms
    t1 = time.perf_counter()
    launches = L.launch_count() - l0
    clocks = sampler.stop(t0, t1)
    barrier(dist, local)
    ms_max = barrier_max(dist, local, ms_total)      # device time (CUDA events), max over ranks
    wall_max = barrier_max(dist, local, (t1 - t0) * 1e3)
    pivots_per_step = piv_total / args.steps
    value = world * piv_total / (ms_max / 1e3)

    # -------- dominant kernel durations: event pairs around each sweep launch (short passes) ----
    KBLK = int(os.environ.get("LPR_TAB_BLOCK", "16"))
    sweep_us = blk_sweep_us = None
    unblocked = None
    try:
        # (a) one sweep per pivot (flags bit4): the north star's per-pivot rank-1 sweep
        t = L.DeviceTableau.dense_lp(SEED, M, NV, device=dev)
        t.solve(L.RULE_PRIMAL, max_pivots=96, log_cap=0, blocked=False, time_sweeps=True)
        us = C.c_float()
        N.check(lib.lpr_tab_last_sweep_us(t._h, C.byref(us)))
        sweep_us = us.value if us.value > 0 else None
        t.close()
        # (b) delayed-update path: one sweep per KBLK pivots
        t = L.DeviceTableau.dense_lp(SEED, M, NV, device=dev)
        t.solve(L.RULE_PRIMAL, max_pivots=16 * KBLK, log_cap=0, blocked=True, time_sweeps=True)
        N.check(lib.lpr_tab_last_sweep_us(t._h, C.byref(us)))
        blk_sweep_us = us.value if us.value > 0 else None
        t.close()
        # (c) throughput of the per-pivot path over a 1024-pivot window (same tableau, same pivots)
        for _ in range(2):
            pu, msu, _, _ = solve_resident(1024, blocked=False)
        unblocked = {"pivots": pu, "ms": msu, "pivots_per_s": pu / (msu / 1e3), "us_per_pivot": msu * 1e3 / pu}
    except Exception as ex:  # measurement aid only
        sys.stderr.write(f"[bench] per-kernel timing pass failed: {ex!r}\n")

    # -------- e2e: host buffers through the C ABI (rank-local replica) --------------------------
    from lpr_381_group_v22_b200 import bench_workloads as W
    Ah, bh, ch = pinned((M, NV)), pinned((M,)), pinned((NV,))
    _, b, c = W.gen_dense_lp(SEED, M, NV, out=Ah)  # the package's own generator, bit identical to the device one
    bh[:] = b; ch[:] = c
    outT, outx = pinned((R, CC)), pinned((NV,))
    outb = np.zeros(M, dtype=np.int32)
    h2d = 8 * (M * NV + M + NV)
    d2h = 8 * (R * CC + NV + 1) + 4 * M

    e2e_phase = {}

    def solve_e2e(max_pivots):
        tp = [time.perf_counter()]
        h = N.vp()
        N.check(lib.lpr_tab_create_primal(dev, NV, M, N.pd(ch), N.pd(Ah), NV, None, None, N.pd(bh), 1, C.byref(h)))
        tp.append(time.perf_counter())
        st, npv, z = C.c_int(), C.c_int64(), C.c_double()
        N.check(lib.lpr_tab_solve(h, L.RULE_PRIMAL, max_pivots, 0, C.byref(st), C.byref(npv), None, 0))
        tp.append(time.perf_counter())
        N.check(lib.lpr_tab_objective(h, C.byref(z)))
        N.check(lib.lpr_tab_extract_solution(h, NV, N.pd(outx)))
        N.check(lib.lpr_tab_get_basis(h, N.pi(outb)))
        tp.append(time.perf_counter())
        N.check(lib.lpr_tab_read(h, N.pd(outT)))
        tp.append(time.perf_counter())
        lib.lpr_tab_destroy(h)
        tp.append(time.perf_counter())
        for name, a, b2 in (("create_h2d_build", 0, 1), ("solve", 1, 2), ("z_x_basis", 2, 3), ("read_tableau_d2h", 3, 4),
                            ("destroy", 4, 5)):
            e2e_phase[name] = round((tp[b2] - tp[a]) * 1e3, 3)
        return npv.value, z.value

    solve_e2e(args.max_pivots)  # warm-up
    barrier(dist, local)
    te0 = time.perf_counter()
    e2e_piv = 0
    e2e_steps = max(1, min(args.steps, 2))
    for _ in range(e2e_steps):
        p, ze = solve_e2e(args.max_pivots)
        e2e_piv += p
    te = time.perf_counter() - te0
    te_max = barrier_max(dist, local, te)
    e2e_value = world * e2e_piv / te_max
    assert ze == zval, "e2e and resident solves disagree"

    # -------- B&B simplex pool through the second driver (one process per GPU over torch.distributed / NCCL): every
    # rank takes part.  Kept beside the in-library driver because a pool is host-API heavy (~40 CUDA calls per 0.7 ms
    # batch) and eight pools driven by eight THREADS of one process contend inside the CUDA driver (4.4x at 8 GPUs),
    # while eight processes do not (7.9x)
    bb_tr = None
    if world > 1 and not env_flag("LPR_BENCH_SKIP_BB"):
        os.environ.setdefault("LPR_BB_PREALLOC_MB", "126976")
        os.environ.setdefault("LPR_BB_MAX_DEPTH", "192")
        from lpr_381_group_v22_b200.bench_workloads import run_bb_cfg5
        try:
            slices = int(os.environ.get("LPR_BENCH_BB_SLICES", "12"))
            bb_tr = run_bb_cfg5(512, 1024, 385, dev, dist, slices * 1024, 1024,
                                float(os.environ.get("LPR_BENCH_BB_SLICE_MS", "10")), with_cuts=False)
        except Exception as ex:
            bb_tr = {"error": repr(ex)}

    # -------- from here on rank 0 works alone; the other ranks wait on the rendezvous store (a CPU wait: a NCCL barrier
    # would park a spinning kernel on their GPUs, which rank 0's in-library multi-GPU legs are about to use) ------------
    barrier(dist, local)
    if rank != 0:
        wait_for_rank0(dist, rank)
        dist.destroy_process_group()
        return

    # -------- cpu baseline: oracle, single thread like the reference, bounded sample ------------------------------------
    import oracle_lib as O  # the checker: cpu_baseline legs only
    cpu = None
    if world == 1:
        sample = int(os.environ.get("LPR_CPU_SAMPLE_PIVOTS", "96"))
        T0, b0 = O.primal_build(list(ch), [(Ah[i], "<=", bh[i]) for i in range(M)])
        tc = time.perf_counter()
        r = O.primal_solve(T0, b0, max_pivots=sample, threads=1, log_cap=0)
        dtc = time.perf_counter() - tc
        del T0, r
        cpu = {"value": sample / dtc, "unit": "pivots/s", "cores": 1, "kind": "port",
               "sample": f"first {sample} pivots of the same cfg2 solve ({dtc:.1f} s); C++ restatement of "
                         "PrimalSimplexSolver.cs:152-211 (no .NET toolchain in the image; pinned bit for bit to "
                         "the reference's own sources executed by oracle/csharp), single thread like "
                         f"the reference; host has {os.cpu_count()} cores"}
    del Ah, outT

    # -------- branch & bound node pools (BASELINE configs[4] / configs[3]) partitioned over the box's GPUs INSIDE the
    # library: rank 0 calls lpr_bb_solve_mgpu / lpr_knap_solve_mgpu with n_gpus = world (host threads + NCCL) ---------
    bb = knap = None
    if not env_flag("LPR_BENCH_SKIP_BB"):
        os.environ.setdefault("LPR_BB_PREALLOC_MB", "126976")  # node slabs carved before the timed region
        os.environ.setdefault("LPR_BB_MAX_DEPTH", "192")        # deep enough for the node budget below
        try:
            bb = run_bb_legs(W, O, world, dev)
        except Exception as ex:  # the headline line must still be printed
            bb = {"error": repr(ex)}
        if bb_tr is not None:
            keep = ("n_gpus", "nodes", "seconds", "nodes_per_s", "pivots_per_node", "open_left", "steals", "rounds",
                    "run_seconds_per_rank", "depth_overflow", "error")
            bb["process_per_gpu_driver"] = {k: bb_tr[k] for k in keep if k in bb_tr}
        try:
            knap = run_knap_legs(W, O, world, dev)
        except Exception as ex:
            knap = {"error": repr(ex)}

    # -------- revised simplex (BASELINE configs[2]): replicas only, measured on rank 0 ------------------------------------
    rev = None
    if not env_flag("LPR_BENCH_SKIP_REV"):
        try:
            rev = W.run_rev_cfg3(8192, 16384, 384, dev)
            rev.update(run_rev_host_legs(W, O, lib, N, C, dev, world))
        except Exception as ex:
            rev = dict(rev or {}, error=repr(ex))
    wait_for_rank0(dist, rank)

    per_pivot_us = ms_max * 1e3 / piv_total
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "sweep_dram_traffic.json")) as f:
            traffic = json.load(f)
    except Exception:
        pass
    peak = float(peaks["hbm_gbs"])
    # dominant kernel of the timed region = the delayed-update sweep: it reads+writes the tableau once (columns
    # 0..C-2; the RHS column lives in the select kernels' mirror) and reads the KBLK pending pivot rows / factor
    # columns; ONE launch applies KBLK pivots.  It runs on the sweep partition (132 of 148 SMs) while the select
    # cluster of the next group runs on the other 16 SMs, so kernel_us is the in-situ (overlapped) duration.
    pipelined = os.environ.get("LPR_TAB_PIPE", "1") != "0"
    bytes_blk = 16.0 * R * (CC - 1 if pipelined else CC) + 8.0 * KBLK * (R + CC)
    kern_us = blk_sweep_us if blk_sweep_us else None
    achieved = bytes_blk / (kern_us * 1e-6) / 1e9 if kern_us else None
    pivot_equiv = BYTES_PER_PIVOT / (per_pivot_us * 1e-6) / 1e9
    per_pivot_kernel = None
    if sweep_us and unblocked:
        ach1 = BYTES_SWEEP / (sweep_us * 1e-6) / 1e9
        achp = BYTES_PER_PIVOT / (unblocked["us_per_pivot"] * 1e-6) / 1e9
        per_pivot_kernel = {
            "what": "one rank-1 sweep per pivot (lpr_tab_solve flags bit4): the north star's per-pivot kernel",
            "kernel": "lpr::k_sweep<0,false,true,4>", "kernel_us": sweep_us, "bytes_per_launch": BYTES_SWEEP,
            "achieved": ach1, "frac_of_measured": ach1 / peak, "frac_of_nominal_8tbs": ach1 / NOMINAL_HBM_GBS,
            "traffic": (traffic or {}).get("dram_bytes_per_launch"),
            "whole_pivot": {"us": unblocked["us_per_pivot"], "pivots_per_s": unblocked["pivots_per_s"],
                            "bytes": BYTES_PER_PIVOT, "achieved": achp, "frac_of_measured": achp / peak,
                            "frac_of_nominal_8tbs": achp / NOMINAL_HBM_GBS}}
    line = {
        "metric": "simplex pivots/s (dense 4096x8192 fp64 primal tableau)",
        "value": value, "unit": "pivots/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": "cfg2 dense LP m=4096 n=8192 primal tableau simplex, tableau 4097x12289 "
                               "(BASELINE.json configs[1])",
                   "step": "one full Solve() from the slack basis to optimality",
                   "pivots_per_step": pivots_per_step, "delayed_update_block": KBLK, "status": L.STATUS_NAMES[status], "seed": SEED,
                   "l2": "tableau 402.8 MB > 126 MB L2 (inputs larger than L2; no flush needed)",
                   "parallelism": "replicas" if world > 1 else "single GPU"},
        "tableau_gbs": 16.0 * R * CC * value / world / 1e9,
        "wall_ms_per_step": wall_max / args.steps,
        "e2e": {"value": e2e_value, "unit": "pivots/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": e2e_steps, "phase_ms_last_step": e2e_phase, "what": "lpr_tab_create_primal(host A,b,c pinned) + lpr_tab_solve + "
                                            "objective + extract_solution + basis + lpr_tab_read(final tableau)"},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": {"bound": "hbm",
                     "kernel": (f"lpr::k_pipe_sweep_ca<16,2,16,32> (applies {KBLK} delayed rank-1 updates per launch, "
                                "cp.async shared-memory ring, out of place, overlapped with lpr::k_pipe_select3<768,1>)")
                     if pipelined else
                     f"lpr::k_blk_sweep<4,{8 if KBLK <= 8 else 16},64> (applies {KBLK} delayed rank-1 updates per launch)",
                     "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": (achieved / peak) if achieved else None,
                     "peak_kind": f"{peak_kind} copy bandwidth (MEASURED_PEAKS.json)",
                     "bytes_per_launch": bytes_blk, "pivots_per_launch": KBLK, "kernel_us": kern_us,
                     "kernel_us_how": "CUDA event pairs around each sweep launch on the sweep stream, 16-group pass, "
                                      "select cluster of the next group running concurrently",
                     "group_period_us": per_pivot_us * KBLK,
                     "traffic": (traffic or {}).get("blocked_dram_bytes_per_launch"),
                     "traffic_how": "from profile: dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture "
                                    "of this kernel (profiles/sweep_dram_traffic.json), not measured in this run",
                     "note": "bit-identical delayed-update path: every element still goes through the same "
                             "multiply/subtract roundings in the same order, but the tableau is swept once per "
                             f"{KBLK} pivots; pivot_equivalent = 16RC-per-pivot bytes x pivots/s",
                     "pivot_equivalent": {"bytes_per_pivot": BYTES_PER_PIVOT, "us_per_pivot": per_pivot_us,
                                          "gbs": pivot_equiv, "x_of_nominal_8tbs": pivot_equiv / NOMINAL_HBM_GBS},
                     "per_pivot_kernel": per_pivot_kernel},
        "cpu_baseline": cpu,
        "bb": bb,
        "knapsack": knap,
        "revised": rev,
    }
    if rev and "dense_rows" in rev:
        rev["roofline"] = {"bound": "hbm", "achieved": rev["dense_rows"]["gbs"], "peak": peak, "unit": "GB/s",
                           "frac": rev["dense_rows"]["gbs"] / peak,
                           "note": "24 m^2 + 8 m n bytes per iteration (A once, B^-1 read once for u and x_B, read+written "
                                   "once by the update) with the zero-multiplier row skip disabled"}
    if bb and "nodes_per_s" in bb:
        # per node 16*R*C for each of the two child copies (AddConstraint) + 16*R*C per pivot executed in
        # the node; R, C taken at the shallowest children (root + one row / one column), i.e. a lower bound
        rc_bytes = 16.0 * (512 + 2) * (1024 + 512 + 2)
        per_node = rc_bytes * (2.0 + bb["pivots_per_node"])
        ach = per_node * bb["nodes_per_s"] / 1e9
        bb["roofline"] = {"bound": "hbm", "bytes_per_node": per_node, "achieved": ach, "peak": peak * world, "unit": "GB/s",
                          "frac": ach / (peak * world),
                          "note": "algorithmic bytes (two child copies + pivots/node sweeps) x nodes/s over the measured "
                                  "copy bandwidth of the GPUs in use; node tableaux of a batch are partly L2 resident"}
    # compact multi-GPU record as the LAST key, so that it survives a truncated tail of the line
    mg = {"n_gpus": world, "how": "node pool partitioned inside the library (lpr_*_solve_mgpu: host threads + NCCL)"}
    if bb and "nodes_per_s" in bb:
        mg.update(bb_nodes_per_s=bb["nodes_per_s"], bb_gbs=bb["roofline"]["achieved"], bb_frac=bb["roofline"]["frac"],
                  bb_bytes_per_node=bb["roofline"]["bytes_per_node"], bb_pivots_per_node=bb["pivots_per_node"])
        ppg = bb.get("process_per_gpu_driver") or {}
        if "nodes_per_s" in ppg:
            rcb = 16.0 * (512 + 2) * (1024 + 512 + 2)
            mg.update(bb_process_per_gpu_nodes_per_s=ppg["nodes_per_s"],
                      bb_process_per_gpu_gbs=rcb * (2.0 + ppg["pivots_per_node"]) * ppg["nodes_per_s"] / 1e9)
        if isinstance(bb.get("closed_instance"), dict):
            mg.update(bb_closed_nodes=bb["closed_instance"].get("nodes"), bb_closed_z=bb["closed_instance"].get("incumbent_z"),
                      bb_incumbent_sha=bb["closed_instance"].get("incumbent_sha"),
                      bb_closed_matches_oracle=bb["closed_instance"].get("matches_oracle"))
    if knap and "nodes_per_s" in knap:
        mg.update(knap_nodes_per_s=knap["nodes_per_s"], knap_best=knap["best_value"], knap_selection_sha=knap["selection_sha"],
                  knap_equals_dp=knap.get("equals_dp"))
        if isinstance(knap.get("hard"), dict) and "nodes_per_s" in knap["hard"]:
            mg.update(knap_hard_nodes_per_s=knap["hard"]["nodes_per_s"], knap_hard_nodes=knap["hard"]["nodes"],
                      knap_hard_best=knap["hard"]["best_value"], knap_hard_selection_sha=knap["hard"]["selection_sha"])
    line["multi_gpu"] = mg
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def wait_for_rank0(dist, rank):
    """ranks != 0 block (on the CPU) until rank 0 has finished the legs it runs alone"""
    if dist is None:
        return
    from datetime import timedelta
    store = dist.distributed_c10d._get_default_store()
    if rank == 0:
        store.set("lpr_rank0_done", "1")
    else:
        store.wait(["lpr_rank0_done"], timedelta(minutes=60))


def run_bb_legs(W, O, world, dev):
    """cfg5 (m=512, n=1024 dense IP): LP relaxation on the tableau path, Gomory cuts at the root, then branch & bound
    simplex with reference semantics, pruning on, 12 time slices of 10 ms per GPU whatever the GPU count (weak scaling);
    plus a cfg5-family instance whose tree CLOSES, for a GPU-count independent incumbent hash checked against the
    sequential oracle; plus the oracle's nodes/s on the cfg5 root as cpu_baseline."""
    import numpy as np
    import lpr_381_group_v22_b200 as L
    m, n, seed = 512, 1024, 385
    A, b, c = W.gen_dense_ip(seed, m, n)
    final, lp, lp_ms = W.lp_relaxation(A, b, c, dev)
    cuts = None
    try:
        for attempt in range(2):  # the first call of a process also allocates the second tableau buffer and scratch
            with L.DeviceTableau.from_host(final, device=dev, row_cap=final.shape[0] + 40) as tc:
                tq = time.perf_counter()
                rc = tc.cutting_plane(max_cuts=32)
                dq = time.perf_counter() - tq
                dev_ms = tc.last_solve_ms
        piv = int(rc["log"][:, 2].sum() + rc["log"][:, 3].sum()) + int(rc["n_cuts"])
        cuts = dict(n_cuts=rc["n_cuts"], status=L.STATUS_NAMES[rc["status"]], seconds=dq, device_ms=dev_ms,
                    dual_pivots=int(rc["log"][:, 2].sum()), primal_pivots=int(rc["log"][:, 3].sum()), pivots=piv,
                    us_per_pivot=dq * 1e6 / max(1, piv), device_us_per_pivot=dev_ms * 1e3 / max(1, piv),
                    what="second of two lpr_tab_cutting_plane(max_cuts=32) calls on the 513x1537 relaxation tableau: host "
                         "clock around the call and CUDA events around its one cooperative launch (k_persist)")
    except Exception as ex:
        cuts = {"error": repr(ex)}
    slices = int(os.environ.get("LPR_BENCH_BB_SLICES", "12"))
    slice_ms = float(os.environ.get("LPR_BENCH_BB_SLICE_MS", "10"))
    # the pools are driven by host threads: a neighbour's burst on the box's CPUs shows up as a slow run, so the
    # measurement is repeated and the median run is the one reported (all of them are listed)
    reps = [W.bb_mgpu(final, n, world, max_rounds=slices, slice_seconds=slice_ms * 1e-3)
            for _ in range(int(os.environ.get("LPR_BENCH_BB_REPEATS", "3")))]
    out = sorted(reps, key=lambda r: r["nodes_per_s"])[len(reps) // 2]
    out["repeats_nodes_per_s"] = [r["nodes_per_s"] for r in reps]
    out.update(workload=f"cfg5 dense IP m={m} n={n} B&B simplex (root {final.shape[0]}x{final.shape[1]}), "
                        f"{slices} slices of {slice_ms:g} ms per GPU",
               lp_relaxation_pivots=lp["n_pivots"], lp_relaxation_ms=lp_ms, root_cuts=cuts,
               note="reference semantics (4-d.p. rounding, first-row-with-a-1 extraction, SURVEY Q8/Q10): this tree never "
                    "closes (the rounding keeps re-branching), so throughput is measured on time slices and the "
                    "GPU-count independent incumbent is demonstrated on closed_instance")
    # oracle on the same root: sequential ExecuteBranchAndBound, bounded sample
    k = int(os.environ.get("LPR_CPU_SAMPLE_BB_NODES", "32"))
    tq = time.perf_counter()
    ref = O.bb_solve(final, n, prune=True, max_nodes=k)
    dq = time.perf_counter() - tq
    out["cpu_baseline"] = {"value": ref["nodes"] / dq, "unit": "nodes/s", "cores": 1, "kind": "port",
                           "sample": f"first {ref['nodes']} nodes of ExecuteBranchAndBound on the same cfg5 root ({dq:.1f} s); "
                                     "C++ restatement of BranchBoundSimplexSolver.cs:1006-1233, single thread"}
    # a cfg5-family instance that terminates: incumbent hash at this GPU count == the sequential oracle's
    try:
        cm, cn, cseed, cdiv = 6, 32, 42, 2.0
        A2, b2, c2 = W.gen_binary_ip(cseed, cm, cn, cdiv)
        f2, _, _ = W.lp_relaxation(A2, b2, c2, dev)
        os.environ["LPR_BB_PREALLOC_MB"] = "256"
        closed = W.bb_mgpu(f2, cn, world, slice_seconds=1e-3)
        refc = O.bb_solve(f2, cn, prune=True, max_nodes=-1, log_cap=1 << 16)
        closed["oracle_nodes"] = int(refc["nodes"])
        closed["matches_oracle"] = bool(refc["has_solution"] and closed["incumbent_z"] == refc["z"]
                                        and closed["incumbent_sha"] == W.sha16(refc["x"]))
        closed["workload"] = f"cfg5 family m={cm} n={cn} + {cn} bound rows, b = rowsum/{cdiv:g}, seed {cseed}: tree closes"
        out["closed_instance"] = closed
    except Exception as ex:
        out["closed_instance"] = {"error": repr(ex)}
    finally:
        os.environ["LPR_BB_PREALLOC_MB"] = "126976"
    return out


def run_knap_legs(W, O, world, dev):
    """cfg4 (n = 10^4 weakly correlated, seed 384) to proven optimality with the reference's own check (B&B == DP,
    Program.cs:467-470) asserted in the run; 'hard' = a strongly correlated instance with seconds of work, which
    is what the scaling figure is quoted on (it closes, so its selection hash must be the same at every GPU count); cpu_baseline = the oracle's depth-first B&B on cfg4."""
    w, v, cap = W.gen_knapsack(384, 10000)
    out, ch = W.knap_mgpu(w, v, cap, world)
    out["workload"] = f"cfg4 knapsack n=10000 weakly correlated, capacity {cap:.0f}, to proven optimality"
    dp, dp_s = W.knap_dp_check(w, v, cap, dev)
    out["dp_value"], out["dp_seconds"] = dp, dp_s
    out["equals_dp"] = bool(abs(out["best_value"] - dp) < 1e-6)
    assert out["equals_dp"], f"knapsack B&B {out['best_value']} != DP {dp}"
    k = int(os.environ.get("LPR_CPU_SAMPLE_KNAP_NODES", "150000"))
    tq = time.perf_counter()
    ref = O.knap_bb(cap, w, v, max_nodes=k)
    dq = time.perf_counter() - tq
    out["cpu_baseline"] = {"value": ref["nodes"] / dq, "unit": "nodes/s", "cores": 1, "kind": "port",
                           "sample": f"first {ref['nodes']} nodes of the oracle's depth-first B&B on cfg4 ({dq:.1f} s), single thread"}
    try:
        hn, hseed = int(os.environ.get("LPR_BENCH_KNAP_HARD_N", "120")), 395
        w2, v2, cap2 = W.gen_knapsack_hard(hseed, hn)
        hard, _ = W.knap_mgpu(w2, v2, cap2, world, max_nodes=int(os.environ.get("LPR_BENCH_KNAP_HARD_NODES", "-1")))
        hard["workload"] = (f"cfg4-hard: strongly correlated (v = w + 100) n={hn}, seed {hseed}, capacity {cap2:.0f}, "
                            "to proven optimality")
        out["hard"] = hard
    except Exception as ex:
        out["hard"] = {"error": repr(ex)}
    return out


def run_rev_host_legs(W, O, lib, N, C, dev, world):
    """cfg3 through the C ABI with HOST arrays (A, b, c pinned -> lpr_rev_create -> 256 iterations -> x, z, y, basis
    back), and the oracle's iterations/s on the same model as cpu_baseline."""
    import numpy as np
    m, n, seed, iters = 8192, 16384, 384, 256
    A = pinned((m, n))
    _, b, c = W.gen_dense_lp(seed, m, n, out=A)
    out = {}
    bb_, cc_ = N.f64(b), N.f64(c)
    h = N.vp()  # warm-up (like the cfg2 e2e leg): the first gigabyte-sized cudaMalloc after the pools were freed is slow
    N.check(lib.lpr_rev_create(dev, m, n, N.pd(A), N.pd(bb_), N.pd(cc_), 0, C.byref(h)))
    lib.lpr_rev_destroy(h)
    t0 = time.perf_counter()
    h = N.vp()
    N.check(lib.lpr_rev_create(dev, m, n, N.pd(A), N.pd(bb_), N.pd(cc_), 0, C.byref(h)))
    t1 = time.perf_counter()
    st, nit, z = C.c_int(), C.c_int64(), C.c_double()
    N.check(lib.lpr_rev_solve(h, iters, 0, C.byref(st), C.byref(nit), None, 0))
    t2 = time.perf_counter()
    x, y, basis = np.zeros(n), np.zeros(m), np.zeros(m, dtype=np.int32)
    N.check(lib.lpr_rev_read_x(h, N.pd(x)))
    N.check(lib.lpr_rev_read_z(h, C.byref(z)))
    N.check(lib.lpr_rev_read_y(h, N.pd(y)))
    N.check(lib.lpr_rev_read_basis(h, N.pi(basis)))
    t3 = time.perf_counter()
    lib.lpr_rev_destroy(h)
    out["e2e"] = {"value": nit.value / (t3 - t0), "unit": "iterations/s", "iterations": nit.value,
                  "h2d_bytes_per_step": 8 * (m * n + m + n), "d2h_bytes_per_step": 8 * (n + m + 1) + 4 * m,
                  "phase_ms": {"create_h2d": (t1 - t0) * 1e3, "solve": (t2 - t1) * 1e3, "read_back": (t3 - t2) * 1e3},
                  "what": "lpr_rev_create(host A, b, c pinned) + lpr_rev_solve(256 iterations) + x, z, y, basis back"}
    if world == 1:
        k1, k2 = 1, int(os.environ.get("LPR_CPU_SAMPLE_REV_ITERS", "5"))
        tq = time.perf_counter(); O.rev_solve(A, b, c, False, max_iter=k1, log_cap=8); d1 = time.perf_counter() - tq
        tq = time.perf_counter(); O.rev_solve(A, b, c, False, max_iter=k2, log_cap=8); d2 = time.perf_counter() - tq
        per = max(1e-9, (d2 - d1) / (k2 - k1))
        out["cpu_baseline"] = {"value": 1.0 / per, "unit": "iterations/s", "cores": 1, "kind": "port",
                               "sample": f"iterations {k1 + 1}..{k2} of the same cfg3 solve ({d2:.1f} s incl. set-up); C++ "
                                         "restatement of RevisedPrimalSimplexSolver.cs:82-251 without the O(m^2 n) snapshot "
                                         "product, single thread"}
    return out


if __name__ == "__main__":
    main()
