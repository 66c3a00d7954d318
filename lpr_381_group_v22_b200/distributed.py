"""Multi-GPU branch & bound: one process per GPU, the open-node pool partitioned across ranks.

The tableau and revised simplex paths do not shard (DESIGN.md "multi-GPU": replicas only).  The two
branch & bound solvers do: nodes are independent units, so every rank owns a slice of the open-node
pool in its GPU's HBM and expands it in batches; between batches the ranks

  1. all-reduce the incumbent value (NCCL all-reduce MAX over NVLink -- the north star's "allreduce-min"
     on the negated objective) and break ties on the DFS path key, so the incumbent -- and therefore
     the final answer -- is the same for every rank count,
  2. learn every rank's open-node count (termination = all zero) -- the counts ride in the same all-reduce
     as the incumbent value (round_status), so a round without a new incumbent costs one collective,
  3. steal work: ranks that ran dry receive the shallowest nodes of the fullest ranks (peer
     send/recv of node records).

The orchestration below only needs a `pool` object with the small interface of BBPool/KnapPool, which
is what the world_size-2 gloo tests exercise with a CPU stand-in pool.
"""
import ctypes as C

import numpy as np

from . import _native as N


# ------------------------------------------------------------------------------------------------
# pools (thin wrappers over the C ABI)
# ------------------------------------------------------------------------------------------------
def _alloc_bytes(cap, torch_device):
    """byte buffer for node records: a CUDA tensor when the records travel over NCCL (the library copies
    device-to-device into it, so stolen nodes never touch host memory), a numpy array otherwise."""
    if torch_device is not None and str(torch_device) != "cpu":
        import torch
        t = torch.empty(cap, dtype=torch.uint8, device=torch_device)
        return t, C.c_void_p(t.data_ptr())
    a = np.empty(cap, dtype=np.uint8)
    return a, a.ctypes.data_as(N.vp)


def _bytes_ptr(data):
    """(owner, pointer, byte count): the caller holds `owner` until the native call that reads the pointer returns"""
    if hasattr(data, "data_ptr"):  # torch tensor (host or device)
        data = data.contiguous()
        return data, C.c_void_p(data.data_ptr()), int(data.numel())
    a = np.ascontiguousarray(data, dtype=np.uint8)
    return a, a.ctypes.data_as(N.vp), int(a.size)


class BBPool:
    """Open-node pool of the branch & bound simplex solver (lpr_bb_*)."""

    def __init__(self, root_tableau, n_vars, prune=True, device=0, rows=None, cols=None):
        self.n_vars = n_vars
        h = N.vp()
        if root_tableau is not None:
            T = N.f64(root_tableau)
            rows, cols = T.shape
            N.check(N.lib().lpr_bb_create(device, rows, cols, N.pd(T), n_vars, int(prune), C.byref(h)))
        else:
            N.check(N.lib().lpr_bb_create(device, rows, cols, None, n_vars, int(prune), C.byref(h)))
        self._h = h
        self.rows, self.cols = rows, cols
        self.pivots = 0

    def close(self):
        if self._h is not None:
            N.lib().lpr_bb_destroy(self._h)
            self._h = None

    def open_count(self):
        n = C.c_int64()
        N.check(N.lib().lpr_bb_open_count(self._h, C.byref(n)))
        return n.value

    def run(self, max_nodes, max_seconds=0.0):
        done, piv = C.c_int64(), C.c_int64()
        N.check(N.lib().lpr_bb_run_timed(self._h, max_nodes, float(max_seconds), C.byref(done), C.byref(piv)))
        self.pivots += piv.value
        return done.value

    def keep_stride(self, offset, stride):
        N.check(N.lib().lpr_bb_keep_stride(self._h, int(offset), int(stride)))

    def stats(self):
        done, piv, ovf, md = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int()
        N.check(N.lib().lpr_bb_stats(self._h, C.byref(done), C.byref(piv), C.byref(ovf), C.byref(md)))
        return dict(processed=done.value, pivots=piv.value, depth_overflow=ovf.value, max_depth=md.value)

    def get_incumbent(self):
        has, z, klen = C.c_int(), C.c_double(), C.c_int(0)
        x = np.zeros(self.n_vars)
        N.check(N.lib().lpr_bb_get_incumbent(self._h, C.byref(has), C.byref(z), N.pd(x), None, C.byref(klen)))
        if not has.value:
            return None
        key = np.zeros(max(1, klen.value), dtype=np.int32)
        kl = C.c_int(klen.value)
        N.check(N.lib().lpr_bb_get_incumbent(self._h, C.byref(has), C.byref(z), N.pd(x), N.pi(key), C.byref(kl)))
        return z.value, tuple(int(k) for k in key[:klen.value]), x

    def set_incumbent(self, value, key, payload):
        k = N.i32(list(key)) if len(key) else None
        x = N.f64(payload)
        N.check(N.lib().lpr_bb_set_incumbent(self._h, float(value), N.pd(x), N.pi(k) if k is not None else None, len(key)))

    def export_nodes(self, max_nodes, torch_device=None):
        md = self.stats()["max_depth"]
        per = 16 + md + 16 + 8 * (self.rows + md) * (self.cols + md)  # header + key + deepest possible tableau
        # the exporter stops when the buffer is full, so a steal is capped in bytes (256 MB) instead of sizing the
        # buffer for `max_nodes` records of the deepest possible tableau (which asked for gigabytes per steal)
        cap = max(int(per), min(int(per) * max(1, max_nodes), 256 << 20))
        buf, ptr = _alloc_bytes(cap, torch_device)
        nbytes, n = C.c_int64(), C.c_int()
        N.check(N.lib().lpr_bb_export_nodes(self._h, max_nodes, ptr, cap, C.byref(nbytes), C.byref(n)))
        return buf[:nbytes.value], n.value

    def import_nodes(self, data):
        owner, ptr, size = _bytes_ptr(data)
        if size:
            N.check(N.lib().lpr_bb_import_nodes(self._h, ptr, size))
        del owner


class KnapPool:
    """Open-node pool of the knapsack branch & bound (lpr_knap_*)."""

    def __init__(self, capacity, weights, values, device=0, with_root=True):
        self.w = N.f64(weights)
        self.v = N.f64(values)
        self.n = len(self.w)
        h = N.vp()
        N.check(N.lib().lpr_knap_create(device, float(capacity), self.n, N.pd(self.w), N.pd(self.v), C.byref(h)))
        self._h = h
        self.rec_bytes = 8 * (3 * ((self.n + 63) // 64) + 4)
        if not with_root:
            self.export_nodes(1)

    def close(self):
        if self._h is not None:
            N.lib().lpr_knap_destroy(self._h)
            self._h = None

    def open_count(self):
        n = C.c_int64()
        N.check(N.lib().lpr_knap_open_count(self._h, C.byref(n)))
        return n.value

    def run(self, max_nodes, max_seconds=0.0):
        done, st = C.c_int64(), C.c_int()
        N.check(N.lib().lpr_knap_run_timed(self._h, max_nodes, float(max_seconds), C.byref(done), C.byref(st)))
        return done.value

    def get_incumbent(self):
        best, kb = C.c_double(), C.c_int()
        ch = np.zeros(self.n, dtype=np.uint8)
        key = np.zeros((self.n + 63) // 64, dtype=np.uint64)
        N.check(N.lib().lpr_knap_get_incumbent(self._h, C.byref(best), ch.ctypes.data_as(N.bp),
                                               key.ctypes.data_as(N.u64p), C.byref(kb)))
        if kb.value < 0:
            return None
        bits = tuple(int((int(key[i >> 6]) >> (i & 63)) & 1) for i in range(kb.value))
        return best.value, bits, ch.astype(np.float64)

    def set_incumbent(self, value, key, payload):
        words = np.zeros((self.n + 63) // 64, dtype=np.uint64)
        for i, b in enumerate(key):
            if b:
                words[i >> 6] |= np.uint64(1) << np.uint64(i & 63)
        ch = np.ascontiguousarray(np.asarray(payload) != 0, dtype=np.uint8)
        N.check(N.lib().lpr_knap_set_incumbent(self._h, float(value), ch.ctypes.data_as(N.bp),
                                               words.ctypes.data_as(N.u64p), len(key)))

    def export_nodes(self, max_nodes, torch_device=None):
        cap = self.rec_bytes * max(1, max_nodes)
        buf, ptr = _alloc_bytes(cap, torch_device)
        nbytes, n = C.c_int64(), C.c_int()
        N.check(N.lib().lpr_knap_export_nodes(self._h, max_nodes, ptr, cap, C.byref(nbytes), C.byref(n)))
        return buf[:nbytes.value], n.value

    def import_nodes(self, data):
        owner, ptr, size = _bytes_ptr(data)
        if size:
            N.check(N.lib().lpr_knap_import_nodes(self._h, ptr, size))
        del owner


# ------------------------------------------------------------------------------------------------
# orchestration
# ------------------------------------------------------------------------------------------------
def better(a, b):
    """incumbent order: larger value first, ties -> DFS-first key (a prefix precedes its extensions)."""
    if b is None:
        return a is not None
    if a is None:
        return False
    if a[0] != b[0]:
        return a[0] > b[0]
    return tuple(a[1]) < tuple(b[1])


class _Comm:
    """torch.distributed plumbing (NCCL on GPUs, gloo in the CPU tests)."""

    def __init__(self, dist, device):
        import torch
        self.torch = torch
        self.dist = dist
        self.rank = dist.get_rank() if dist is not None else 0
        self.world = dist.get_world_size() if dist is not None else 1
        self.dev = device if (dist is not None and dist.get_backend() == "nccl") else "cpu"

    def allreduce_max(self, value):
        if self.dist is None:
            return value
        t = self.torch.tensor([value], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def allreduce_max_vec(self, values):
        """one all-reduce(MAX) over a short float64 vector (the per-round status word of run_distributed)"""
        if self.dist is None:
            return list(values)
        t = self.torch.tensor(list(values), dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return t.tolist()

    def allgather_ints(self, values):
        if self.dist is None:
            return [list(values)]
        t = self.torch.tensor(list(values), dtype=self.torch.int64, device=self.dev)
        out = [self.torch.zeros_like(t) for _ in range(self.world)]
        self.dist.all_gather(out, t)
        return [o.tolist() for o in out]

    def bcast_array(self, arr, src, dtype):
        if self.dist is None:
            return arr
        t = self.torch.as_tensor(np.ascontiguousarray(arr, dtype=dtype)).to(self.dev)
        self.dist.broadcast(t, src=src)
        return t.cpu().numpy()

    def send_bytes(self, data, dst):
        if hasattr(data, "data_ptr"):
            t = data.contiguous()
        else:
            t = self.torch.from_numpy(np.ascontiguousarray(data, dtype=np.uint8)).to(self.dev)
        self.dist.send(t, dst=dst)
        if self.dev != "cpu":
            self.torch.cuda.current_stream(t.device).synchronize()  # keep `t` alive until NCCL has read it

    def recv_bytes(self, nbytes, src, keep_on_device=False):
        t = self.torch.empty(nbytes, dtype=self.torch.uint8, device=self.dev)
        self.dist.recv(t, src=src)
        if keep_on_device and self.dev != "cpu":
            # NCCL recv only enqueues: the library reads this buffer from its own stream, so the bytes
            # must have landed before the pointer is handed over
            self.torch.cuda.current_stream(t.device).synchronize()
            return t
        return t.cpu().numpy()


def exchange_incumbent(pool, comm, payload_len, zmax=None):
    """all-reduce MAX of the value (skipped when the caller already has it), then the DFS-first key among the
    ranks that hold that value; the winner's payload (x or the chosen-item vector) is broadcast so every rank
    ends with the same incumbent."""
    inc = pool.get_incumbent()
    if zmax is None:
        zmax = comm.allreduce_max(inc[0] if inc is not None else float("-inf"))
    if zmax == float("-inf"):
        return None
    mine = inc is not None and inc[0] == zmax
    lens = [v[0] for v in comm.allgather_ints([len(inc[1]) if mine else -1])]
    kmax = max(max(lens), 1)
    pad = [-1] * kmax
    if mine:
        pad[:len(inc[1])] = list(inc[1])
    allk = comm.allgather_ints(pad)
    key, owner = min((tuple(k[:lens[r]]), r) for r, k in enumerate(allk) if lens[r] >= 0)
    payload = inc[2] if comm.rank == owner else np.zeros(payload_len)
    payload = comm.bcast_array(payload, owner, np.float64)
    best = (zmax, key, payload)
    if better(best, inc):
        pool.set_incumbent(*best)
    return best


def round_status(pool, comm, agreed):
    """The per-round exchange, ONE all-reduce(MAX) over [z, changed, count_0 .. count_{world-1}]: the best
    incumbent value of any rank, whether any rank's incumbent differs from the one agreed on last time (only
    then are keys and the payload exchanged), and every rank's open-node count (each rank fills its own slot;
    counts are >= 0, so MAX against the other ranks' zeros returns them unchanged)."""
    inc = pool.get_incumbent()
    mine = None if inc is None else (inc[0], tuple(inc[1]))
    same = agreed is not None and mine == (agreed[0], tuple(agreed[1])) if mine is not None else agreed is None
    vec = [float("-inf") if inc is None else float(inc[0]), 0.0 if same else 1.0] + [0.0] * comm.world
    vec[2 + comm.rank] = float(pool.open_count())
    out = comm.allreduce_max_vec(vec)
    return out[0], out[1] != 0.0, [int(c) for c in out[2:]]


def warmup_comm(dist, device, payload_len=1):
    """NCCL sets up the channels of each collective kind and of each point-to-point pair lazily (tens to
    hundreds of ms): run every one once, outside any timed region, so that the first incumbent exchange and the
    first steal do not pay for it."""
    if dist is None:
        return
    comm = _Comm(dist, device)
    for _ in range(2):  # every collective run_distributed uses, at the sizes it uses them
        comm.allreduce_max(0.0)
        comm.allreduce_max_vec([0.0] * (2 + comm.world))
        comm.allgather_ints([0])
        comm.allgather_ints([0] * 64)
        comm.allgather_ints([0, 0, 0, 0])
        for src in range(comm.world):
            comm.bcast_array(np.zeros(payload_len), src, np.float64)
    token = np.zeros(8, dtype=np.uint8)
    for a in range(comm.world):          # every unordered pair once, lower rank sends first
        for b in range(a + 1, comm.world):
            if comm.rank == a:
                comm.send_bytes(token, b)
                comm.recv_bytes(8, b)
            elif comm.rank == b:
                comm.recv_bytes(8, a)
                comm.send_bytes(token, a)
    dist.barrier()


def steal_plan(counts, min_keep=2, low_water=0):
    """deterministic (donor, receiver, n) plan computed identically on every rank: ranks whose pool is
    empty (or at most `low_water`) receive half of the pool of the currently fullest rank."""
    counts = list(counts)
    plan = []
    receivers = [r for r, c in enumerate(counts) if c <= low_water]
    for r in receivers:
        donor = max(range(len(counts)), key=lambda q: (counts[q], -q))
        give = counts[donor] // 2
        if donor == r or counts[donor] < min_keep or give < 1 or counts[donor] <= 2 * max(1, low_water):
            continue
        plan.append((donor, r, give))
        counts[donor] -= give
        counts[r] += give
    return plan


def run_distributed(pool, dist=None, device="cpu", chunk_nodes=4096, payload_len=None, max_rounds=1 << 30,
                    seed_nodes_per_rank=8, low_water=0, chunk_seconds=0.0, replicated_root=False):
    """Drive `pool` (rank-local) to completion together with the other ranks.  Returns a dict with the
    incumbent (identical on every rank), node counts and exchange statistics.  A round ends after
    `chunk_nodes` nodes or, when `chunk_seconds` > 0 and the pool supports it, after that time slice --
    whichever comes first: node costs differ between subtrees (pivots per node), so time-sliced rounds keep
    every rank busy until the exchange instead of waiting for the rank with the most expensive nodes.
    `replicated_root`: every rank was given the same root; all of them expand it identically (the kernels are
    bit-reproducible) and keep every world-th open node, so the start-up partition needs no transfer."""
    import time
    comm = _Comm(dist, device)
    rank, world = comm.rank, comm.world
    processed, rounds, steals, moved = 0, 0, 0, 0
    t_seed = t_steal = t_run = t_inc = 0.0
    t_a = time.perf_counter()
    if payload_len is None:
        payload_len = getattr(pool, "n_vars", None) or getattr(pool, "n")
    # seeding: rank 0 owns the root; expand it a little so that the first steal round has work to share
    if world > 1 and (rank == 0 or replicated_root):
        guard = 0
        while 0 < pool.open_count() < seed_nodes_per_rank * world and guard < 64:
            n = pool.run(max(1, seed_nodes_per_rank))
            processed += n if rank == 0 else 0  # the replicas' copies of the seed nodes are not counted
            guard += 1
        if replicated_root:
            pool.keep_stride(rank, world)
    t_seed = time.perf_counter() - t_a
    agreed = None
    while rounds < max_rounds:
        rounds += 1
        t_a = time.perf_counter()
        zmax, changed, counts = round_status(pool, comm, agreed)
        if changed:
            agreed = exchange_incumbent(pool, comm, payload_len, zmax) if world > 1 else pool.get_incumbent()
        t_b = time.perf_counter()
        if sum(counts) == 0:
            break
        for donor, recv, give in steal_plan(counts, low_water=low_water):
            if rank == donor:
                try:
                    data, n = pool.export_nodes(give, comm.dev)  # device staging when the pool supports it
                except TypeError:
                    data, n = pool.export_nodes(give)
                size = int(data.numel()) if hasattr(data, "numel") else int(data.size)
                hdr = np.array([size], dtype=np.int64).view(np.uint8)
                comm.send_bytes(hdr, recv)
                if size:
                    comm.send_bytes(data, recv)
                steals += 1
                moved += n
            elif rank == recv:
                nbytes = int(comm.recv_bytes(8, donor).view(np.int64)[0])
                if nbytes:
                    pool.import_nodes(comm.recv_bytes(nbytes, donor, keep_on_device=True))
        t_c = time.perf_counter()
        if pool.open_count() > 0:
            processed += pool.run(chunk_nodes, chunk_seconds) if chunk_seconds > 0 else pool.run(chunk_nodes)
        t_d = time.perf_counter()
        t_inc += t_b - t_a
        t_steal += t_c - t_b
        t_run += t_d - t_c
    best = exchange_incumbent(pool, comm, payload_len) if world > 1 else pool.get_incumbent()
    # subtrees cut at the slab depth headroom make the incumbent unproven: callers must see it
    ovf = pool.stats().get("depth_overflow", 0) if hasattr(pool, "stats") else 0
    totals = comm.allgather_ints([processed, steals, moved, int(t_run * 1e6), int(ovf)])
    return dict(incumbent=best, nodes_local=processed, nodes_total=sum(t[0] for t in totals),
                depth_overflow=sum(t[4] for t in totals),
                steals=sum(t[1] for t in totals), nodes_moved=sum(t[2] for t in totals), rounds=rounds,
                world=world, rank=rank, run_seconds_per_rank=[t[3] * 1e-6 for t in totals],
                seconds_rank0=dict(seed=t_seed, status_and_incumbent=t_inc, steal=t_steal, run=t_run))
