"""Synthetic workloads of BASELINE.json configs[3] (knapsack B&B) and configs[4] (B&B simplex) driven
through the distributed node-pool driver; shared by bench.py and tools/bb_bench.py.  The instances are
generated on the device / with the package's own generator (no oracle on this path)."""
import ctypes as C
import time

import numpy as np

from . import _native as N
from .distributed import BBPool, KnapPool, _Comm, run_distributed, warmup_comm
from .tableau import DeviceTableau


def _splitmix64(x):
    x = (x + np.uint64(0x9E3779B97F4A7C15))
    z = x
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def u01(seed, stream, idx):
    """counter based generator of SURVEY.md 8(d): bit identical to the CUDA and oracle generators"""
    with np.errstate(over="ignore"):
        k = np.uint64(seed) + (np.uint64(stream) << np.uint64(40)) + np.asarray(idx, dtype=np.uint64)
        return (_splitmix64(k) >> np.uint64(11)).astype(np.float64) * 2.0 ** -53


def gen_dense_lp(seed, m, n, out=None):
    """SURVEY 8(d) cfg2 / cfg3 model: A = 0.1 + u, b = (n/4)(1 + u), c = 1 + u (bit identical to the CUDA generator)"""
    A = out if out is not None else np.empty((m, n))
    step = max(1, (1 << 24) // n)
    for r0 in range(0, m, step):  # row blocks: the index array of the whole matrix would be another gigabyte
        r1 = min(m, r0 + step)
        A[r0:r1] = 0.1 + u01(seed, 0, np.arange(r0 * n, r1 * n, dtype=np.uint64)).reshape(r1 - r0, n)
    b = (n / 4.0) * (1.0 + u01(seed, 1, np.arange(m, dtype=np.uint64)))
    c = 1.0 + u01(seed, 2, np.arange(n, dtype=np.uint64))
    return A, b, c


def gen_dense_ip(seed, m, n):
    A = 1.0 + np.floor(20.0 * u01(seed, 0, np.arange(m * n, dtype=np.uint64))).reshape(m, n)
    b = np.floor(A.sum(axis=1) / 4.0)
    c = 1.0 + np.floor(30.0 * u01(seed, 2, np.arange(n, dtype=np.uint64)))
    return A, b, c


def gen_knapsack(seed, n):
    w = 1.0 + np.floor(1000.0 * u01(seed, 0, np.arange(n, dtype=np.uint64)))
    v = np.maximum(1.0, w + np.floor(200.0 * u01(seed, 1, np.arange(n, dtype=np.uint64))) - 100.0)
    return w, v, float(np.floor(w.sum() / 2.0))


def gen_binary_ip(seed, m, n, div):
    """cfg5-family instance whose tree closes under the reference's semantics: the cfg5 generator's A and c,
    b = floor(row sum / div), plus one `x_j <= 1` row per variable (what menu option 3 appends, Program.cs:372-382)"""
    A, _, c = gen_dense_ip(seed, m, n)
    b = np.floor(A.sum(axis=1) / div)
    return np.vstack([A, np.eye(n)]), np.concatenate([b, np.ones(n)]), c


def sha16(arr):
    import hashlib
    return hashlib.sha256(np.ascontiguousarray(arr).tobytes()).hexdigest()[:16]


def lp_relaxation(A, b, c, device):
    """final tableau of the LP relaxation, solved by the tableau path on `device`"""
    m, n = A.shape
    coef = N.f64(A); rhs = N.f64(b); obj = N.f64(c)
    h = N.vp()
    N.check(N.lib().lpr_tab_create_primal(device, n, m, N.pd(obj), N.pd(coef), n, None, None, N.pd(rhs), 1, C.byref(h)))
    tab = DeviceTableau(h)
    lp = tab.solve(log_cap=0)
    ms = tab.last_solve_ms
    final = tab.read()
    tab.close()
    return final, lp, ms


def bb_mgpu(final, n_vars, n_gpus, max_rounds=-1, slice_seconds=0.0, max_nodes=-1):
    """lpr_bb_solve_mgpu: the node pool partitioned over n_gpus devices INSIDE the library (host threads + NCCL)"""
    from .integer_programming import solve_bb_mgpu
    r = solve_bb_mgpu(final, n_vars, True, n_gpus=n_gpus, max_nodes=max_nodes, max_rounds=max_rounds,
                      slice_seconds=slice_seconds)
    st = r["stats"]
    dt = max(st["seconds"], 1e-9)
    return dict(n_gpus=n_gpus, nodes=r["nodes"], seconds=dt, nodes_per_s=r["nodes"] / dt, pivots_in_nodes=r["pivots"],
                pivots_per_node=r["pivots"] / max(1, r["nodes"]), status=STATUS(r["status"]),
                incumbent_z=(r["z"] if r["has_solution"] else None),
                incumbent_sha=(sha16(r["x"]) if r["has_solution"] else None), open_left=st["open_left"],
                depth_overflow=st["depth_overflow"], rounds=st["rounds"], steals=st["steals"],
                nodes_moved=st["nodes_moved"], nodes_per_gpu=st["nodes_per_gpu"],
                run_seconds_per_gpu=[round(v, 4) for v in st["run_seconds_per_gpu"]],
                phase_seconds_rank0=dict(seed=st["seed_seconds"], exchange=st["exchange_seconds"], steal=st["steal_seconds"]),
                setup_seconds=st["setup_seconds"], nccl_version=st["nccl_version"])


def knap_mgpu(w, v, cap, n_gpus, max_nodes=-1, slice_seconds=2e-3):
    """lpr_knap_solve_mgpu: same, for the knapsack pool"""
    n = len(w)
    best, nodes, st = C.c_double(), C.c_int64(), C.c_int()
    ch = np.zeros(n, dtype=np.uint8)
    stats = N.MgpuStats()
    N.check(N.lib().lpr_knap_solve_mgpu(n_gpus, None, float(cap), n, N.pd(N.f64(w)), N.pd(N.f64(v)), max_nodes, -1,
                                        float(slice_seconds), C.byref(best), ch.ctypes.data_as(N.bp), C.byref(nodes),
                                        C.byref(st), C.byref(stats)))
    d = stats.as_dict()
    dt = max(d["seconds"], 1e-9)
    rec_bytes = 8 * (3 * ((n + 63) // 64) + 4)
    return dict(n_gpus=n_gpus, nodes=nodes.value, seconds=dt, nodes_per_s=nodes.value / dt, status=STATUS(st.value),
                best_value=best.value, selection_sha=sha16(ch), items_chosen=int(ch.sum()),
                weight_used=float(np.dot(ch, w)), node_record_bytes=rec_bytes,
                record_gbs=4.0 * rec_bytes * nodes.value / dt / 1e9, open_left=d["open_left"], rounds=d["rounds"],
                steals=d["steals"], nodes_moved=d["nodes_moved"], nodes_per_gpu=d["nodes_per_gpu"],
                run_seconds_per_gpu=[round(x, 4) for x in d["run_seconds_per_gpu"]],
                phase_seconds_rank0=dict(seed=d["seed_seconds"], exchange=d["exchange_seconds"], steal=d["steal_seconds"]),
                setup_seconds=d["setup_seconds"], nccl_version=d["nccl_version"]), ch


def knap_dp_check(w, v, cap, device=0):
    """KnapsackBranchBoundSolver.Solve(int, int[], int[]) on the device: the arbiter of Program.cs:467-470"""
    wi = N.i32(w.astype(np.int64)); vi = N.i32(v.astype(np.int64))
    best = C.c_double()
    ch = np.zeros(len(wi), dtype=np.uint8)
    t = time.perf_counter()
    N.check(N.lib().lpr_knap_dp(device, int(cap), len(wi), N.pi(wi), N.pi(vi), C.byref(best), ch.ctypes.data_as(N.bp)))
    return best.value, time.perf_counter() - t


def gen_knapsack_hard(seed, n, R=1000):
    """'cfg4-hard': Pisinger's strongly correlated family, v = w + R/10 -- every item has nearly the same value/weight
    ratio, so the LP bound prunes late: n = 120 (seed 395) is a 1.58 G-node tree, 2.5 s of work for one B200, and it
    CLOSES, so the selection can be compared between GPU counts (tools/knap_probe.py has the survey of families)"""
    idx = np.arange(n, dtype=np.uint64)
    w = 1.0 + np.floor(R * u01(seed, 0, idx))
    v = w + R / 10
    return w, v, float(np.floor(w.sum() / 2.0))

def _sync_time(comm, t):
    return comm.allreduce_max(t)


def run_bb_cfg5(m, n, seed, device, dist, max_nodes, chunk, slice_ms=0.0, with_cuts=True):
    """LP relaxation with the tableau solver, then branch & bound simplex with reference semantics
    (4-d.p. rounding, dual-then-primal node solves), node cap lifted to `max_nodes` per rank-round budget,
    pruning on; the pool is partitioned across the ranks.  `slice_ms` > 0: rounds are time slices of that
    length (at most `chunk` nodes each), the number of rounds stays max_nodes / chunk."""
    import os
    os.environ.setdefault("LPR_BB_PREALLOC_MB", "126976")  # node slabs carved before the timed region
    os.environ.setdefault("LPR_BB_MAX_DEPTH", "192")       # deep enough for the node budgets used here
    comm = _Comm(dist, f"cuda:{device}")
    A, b, c = gen_dense_ip(seed, m, n)
    coef = N.f64(A); rhs = N.f64(b); obj = N.f64(c)
    h = N.vp()
    N.check(N.lib().lpr_tab_create_primal(device, n, m, N.pd(obj), N.pd(coef), n, None, None, N.pd(rhs), 1, C.byref(h)))
    tab = DeviceTableau(h)
    t0 = time.perf_counter()
    lp = tab.solve(log_cap=0)
    lp_ms = tab.last_solve_ms
    final = tab.read()
    tab.close()
    # Gomory cuts at the root (the "cutting plane" half of BASELINE configs[4]): CuttingPlaneSolver.CuttingPlaneSolution
    # on the relaxation's final tableau, cut rows generated on the device; reported beside the tree search, which like
    # the reference's menu path starts from the relaxation itself.  Rank 0 only, outside the timed region.
    cuts = None
    if comm.rank == 0 and with_cuts:
        try:
            with DeviceTableau.from_host(final, device=device, row_cap=final.shape[0] + 40) as tc:
                tq = time.perf_counter()
                rc = tc.cutting_plane(max_cuts=32)
                dq = time.perf_counter() - tq
                cuts = dict(n_cuts=rc["n_cuts"], status=N.STATUS_NAMES[rc["status"]], seconds=dq,
                            dual_pivots=int(rc["log"][:, 2].sum()), primal_pivots=int(rc["log"][:, 3].sum()),
                            cuts_per_s=rc["n_cuts"] / dq if dq > 0 else None,
                            what="host clock around lpr_tab_cutting_plane(max_cuts=32) on the 513x1537 relaxation tableau")
        except Exception as ex:
            cuts = {"error": repr(ex)}
    # every rank solved the same relaxation to the same bits: all start from the root and split it without a
    # transfer (run_distributed, replicated_root)
    pool = BBPool(final, n, prune=True, device=device)
    warmup_comm(dist, f"cuda:{device}", n)
    t1 = time.perf_counter()
    res = run_distributed(pool, dist, f"cuda:{device}", chunk_nodes=chunk, payload_len=n,
                          max_rounds=max(1, max_nodes // max(1, chunk)), chunk_seconds=slice_ms * 1e-3,
                          replicated_root=True)
    dt = _sync_time(comm, time.perf_counter() - t1)
    piv = sum(v[0] for v in comm.allgather_ints([pool.pivots]))
    left = sum(v[0] for v in comm.allgather_ints([pool.open_count()]))
    st = pool.stats()
    ovf = sum(v[0] for v in comm.allgather_ints([st["depth_overflow"]]))
    pool.close()
    inc = res["incumbent"]
    return dict(depth_overflow=ovf, max_depth=st["max_depth"],
                note="reference semantics (4-d.p. rounding, first-row-with-a-1 extraction, SURVEY Q8/Q10): "
                     "many children need no pivot; depth_overflow > 0 means chains hit the slab depth headroom",
                workload=f"cfg5 dense IP m={m} n={n} B&B simplex (root {final.shape[0]}x{final.shape[1]})",
                n_gpus=comm.world, nodes=res["nodes_total"], seconds=dt, nodes_per_s=res["nodes_total"] / dt,
                pivots_in_nodes=piv, pivots_per_node=piv / max(1, res["nodes_total"]),
                node_pivots_per_s=piv / dt, lp_relaxation_pivots=lp["n_pivots"], lp_relaxation_ms=lp_ms,
                incumbent_z=(inc[0] if inc else None), incumbent_nonzeros=(int(np.count_nonzero(inc[2])) if inc else None),
                open_left=left, steals=res["steals"], nodes_moved=res["nodes_moved"], rounds=res["rounds"],
                finished=(left == 0), phase_seconds_rank0=res["seconds_rank0"],
                run_seconds_per_rank=res["run_seconds_per_rank"], root_cuts=cuts)


def run_knap_cfg4(n_items, seed, device, dist, max_nodes, chunk):
    comm = _Comm(dist, f"cuda:{device}")
    w, v, cap = gen_knapsack(seed, n_items)
    pool = KnapPool(cap, w, v, device=device, with_root=(comm.rank == 0))
    warmup_comm(dist, f"cuda:{device}", n_items)
    t1 = time.perf_counter()
    res = run_distributed(pool, dist, f"cuda:{device}", chunk_nodes=chunk, payload_len=n_items,
                          max_rounds=max(1, max_nodes // max(1, chunk)), seed_nodes_per_rank=64, low_water=64)
    dt = _sync_time(comm, time.perf_counter() - t1)
    left = sum(x[0] for x in comm.allgather_ints([pool.open_count()]))
    pool.close()
    inc = res["incumbent"]
    rec_bytes = 8 * (3 * ((n_items + 63) // 64) + 4)
    return dict(workload=f"cfg4 knapsack n={n_items} weakly correlated, capacity {cap:.0f}", n_gpus=comm.world,
                nodes=res["nodes_total"], seconds=dt, nodes_per_s=res["nodes_total"] / dt,
                node_record_bytes=rec_bytes, record_gbs=3.0 * rec_bytes * res["nodes_total"] / dt / 1e9,
                best_value=(inc[0] if inc else None), items_chosen=(int(np.count_nonzero(inc[2])) if inc else None),
                weight_used=(float(np.dot(inc[2], w)) if inc else None), open_left=left, steals=res["steals"],
                nodes_moved=res["nodes_moved"], rounds=res["rounds"], finished=(left == 0),
                phase_seconds_rank0=res["seconds_rank0"], run_seconds_per_rank=res["run_seconds_per_rank"])


def run_rev_cfg3(m, n, seed, device, iters=256):
    """BASELINE configs[2]: revised primal simplex on a synthetic dense LP generated in HBM, `iters` iterations from
    the slack basis (device-timed), once as shipped and once with the zero-multiplier row skip disabled (every
    iteration then moves the full 24 m^2 + 8 m n bytes), plus one on-device refactorisation of B^-1."""
    import os
    lib = N.lib()
    byts = 24.0 * m * m + 8.0 * m * n
    out = dict(workload=f"cfg3 dense LP m={m} n={n} revised primal simplex, B^-1 {m}x{m} on device",
               bytes_per_iteration=byts, iterations=iters)
    for tag, dense in (("as_shipped", False), ("dense_rows", True)):
        if dense:
            os.environ["LPR_REV_DENSE"] = "1"
        try:
            h = N.vp()
            N.check(lib.lpr_rev_create_dense_lp(device, seed, m, n, C.byref(h)))
            st, nit, ms = C.c_int(), C.c_int64(), C.c_float()
            N.check(lib.lpr_rev_solve(h, iters, 0, C.byref(st), C.byref(nit), None, 0))
            N.check(lib.lpr_rev_last_solve_ms(h, C.byref(ms)))
            out[tag] = dict(us_per_iteration=ms.value * 1e3 / max(1, nit.value), iterations_per_s=nit.value / ms.value * 1e3,
                            gbs=byts * nit.value / ms.value / 1e6, status=STATUS(st.value))
            if dense:
                rms, res, fl = C.c_float(), C.c_double(), C.c_double()
                N.check(lib.lpr_rev_refactor(h))
                N.check(lib.lpr_rev_last_refactor_ms(h, C.byref(rms)))
                N.check(lib.lpr_rev_last_refactor_info(h, C.byref(res), C.byref(fl)))
                out["refactorisation"] = dict(ms=rms.value, fp64_tflops=fl.value / rms.value / 1e9,
                                              residual_before=res.value,
                                              what="Newton-Schulz refresh of B^-1, two m^3 FP64 DMMA GEMMs (mma.sync m8n8k4)")
            lib.lpr_rev_destroy(h)
        finally:
            if dense:
                os.environ.pop("LPR_REV_DENSE", None)
    return out


def STATUS(code):
    return N.STATUS_NAMES[code] if 0 <= code < len(N.STATUS_NAMES) else str(code)
