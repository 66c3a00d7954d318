"""World-size-2 (and 3) gloo tests on CPU of the multi-GPU branch & bound orchestration
(lpr_381_group_v22_b200/distributed.py): incumbent all-reduce + DFS-key tie break, termination,
work stealing.  The GPU pools need a device, so a CPU stand-in pool with the same interface is used:
a knapsack B&B with the semantics of the oracle (orc_knap_bb), which is also the checker."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)


class CpuKnapPool:
    """stand-in for KnapPool: same node semantics as knapsack.cu / orc_knap_bb, nodes kept in Python"""

    def __init__(self, capacity, w, v, with_root=True):
        self.cap, self.n = float(capacity), len(w)
        order = sorted(range(self.n), key=lambda i: (-(v[i] / w[i]), i))
        self.rank = order
        self.w = [float(w[i]) for i in order]
        self.v = [float(v[i]) for i in order]
        self.open = [((-1,) * self.n, ())] if with_root else []  # (fix tuple, key)
        self.inc = None

    def open_count(self):
        return len(self.open)

    def _eval(self, fix):
        cap = self.cap - sum(self.w[p] for p in range(self.n) if fix[p] == 1)
        val = sum(self.v[p] for p in range(self.n) if fix[p] == 1)
        if cap < 0:
            return ("inf", None, None, None)
        sel = [1 if fix[p] == 1 else 0 for p in range(self.n)]
        crit = -1
        for p in range(self.n):
            if fix[p] != -1:
                continue
            if self.w[p] <= cap:
                cap -= self.w[p]
                val += self.v[p]
                sel[p] = 1
            else:
                crit = p
                break
        if crit < 0 or cap == 0:
            return ("cand", val, sel, crit)
        return ("branch", val + self.v[crit] * (cap / self.w[crit]), sel, crit)

    def run(self, max_nodes, max_seconds=0.0):
        import time
        done = 0
        t0 = time.perf_counter()
        while self.open and done < max_nodes:
            if max_seconds > 0 and done > 0 and time.perf_counter() - t0 >= max_seconds:
                break  # time slice over (lpr_bb_run_timed semantics: at least one node per call)
            fix, key = self.open.pop()
            done += 1
            kind, val, sel, crit = self._eval(fix)
            if kind == "inf":
                continue
            if kind == "cand":
                if self.inc is None or val > self.inc[0] or (val == self.inc[0] and key < self.inc[1]):
                    ch = np.zeros(self.n)
                    for p in range(self.n):
                        if sel[p]:
                            ch[self.rank[p]] = 1
                    self.inc = (val, key, ch)
                continue
            if self.inc is not None and (val < self.inc[0] or (val == self.inc[0] and key > self.inc[1])):
                continue
            one = list(fix); one[crit] = 1
            zero = list(fix); zero[crit] = 0
            self.open.append((tuple(one), key + (1,)))
            self.open.append((tuple(zero), key + (0,)))
        return done

    def keep_stride(self, offset, stride):
        self.open = [nd for i, nd in enumerate(self.open) if i % stride == offset]

    def get_incumbent(self):
        return self.inc

    def set_incumbent(self, value, key, payload):
        cand = (value, tuple(key), np.asarray(payload, dtype=float))
        if self.inc is None or cand[0] > self.inc[0] or (cand[0] == self.inc[0] and cand[1] < self.inc[1]):
            self.inc = cand

    def export_nodes(self, k):
        out = self.open[:k]
        self.open = self.open[k:]
        rows = []
        for fix, key in out:
            rows.append(np.concatenate([[len(key)], np.array(fix) + 1, np.array(key + (0,) * (self.n - len(key)))]))
        data = np.array(rows, dtype=np.int16).tobytes() if rows else b""
        return np.frombuffer(data, dtype=np.uint8).copy(), len(out)

    def import_nodes(self, data):
        a = np.frombuffer(np.ascontiguousarray(data).tobytes(), dtype=np.int16).reshape(-1, 1 + 2 * self.n)
        for row in a:
            kl = int(row[0])
            self.open.append((tuple(int(x) - 1 for x in row[1:1 + self.n]), tuple(int(x) for x in row[1 + self.n:1 + self.n + kl])))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, seed, n, chunk, q, replicated=False, slice_s=0.0):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle_lib as O
    from lpr_381_group_v22_b200.distributed import run_distributed
    w, v, cap = O.gen_knapsack(seed, n)
    pool = CpuKnapPool(cap, w, v, with_root=(rank == 0 or replicated))
    res = run_distributed(pool, dist, "cpu", chunk_nodes=chunk, payload_len=n, seed_nodes_per_rank=2,
                          replicated_root=replicated, chunk_seconds=slice_s)
    inc = res["incumbent"]
    q.put((rank, inc[0], tuple(inc[1]), inc[2].tolist(), res["nodes_total"], res["steals"], res["nodes_moved"]))
    dist.destroy_process_group()


@pytest.mark.parametrize("world,seed,n,chunk", [(2, 3, 24, 8), (2, 5, 30, 4), (3, 7, 28, 5)])
def test_distributed_bb_same_answer_as_sequential(world, seed, n, chunk):
    import oracle_lib as O
    w, v, cap = O.gen_knapsack(seed, n)
    ref = O.knap_bb(cap, w, v)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, seed, n, chunk, q)) for r in range(world)]
    for p in procs:
        p.start()
    outs = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    outs.sort()
    for rank, best, key, chosen, total, steals, moved in outs:
        assert best == ref["best"]
        assert [int(c) for c in chosen] == ref["chosen"].tolist()  # DFS-first optimum on every rank
    assert len({o[2] for o in outs}) == 1  # same incumbent key everywhere
    assert outs[0][5] >= 1 and outs[0][6] >= 1  # work was actually stolen
    assert outs[0][4] >= ref["nodes"] // 4


@pytest.mark.parametrize("world,seed,n,chunk,slice_s", [(2, 4, 26, 6, 0.0), (3, 8, 27, 5, 0.0), (2, 6, 28, 10 ** 6, 2e-4)])
def test_distributed_bb_replicated_root(world, seed, n, chunk, slice_s):
    """every rank expands the same root and keeps every world-th node (no start-up transfer): same answer, and
    the seed nodes are counted once; the last case cuts the rounds by time slices instead of node counts"""
    import oracle_lib as O
    w, v, cap = O.gen_knapsack(seed, n)
    ref = O.knap_bb(cap, w, v)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, seed, n, chunk, q, True, slice_s)) for r in range(world)]
    for p in procs:
        p.start()
    outs = sorted(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, best, key, chosen, total, steals, moved in outs:
        assert best == ref["best"]
        assert [int(c) for c in chosen] == ref["chosen"].tolist()
    assert len({o[2] for o in outs}) == 1
    assert len({o[4] for o in outs}) == 1 and outs[0][4] >= ref["nodes"] // 4


def test_round_status_single_process():
    """the fused status word: value, 'somebody's incumbent changed' flag, per-rank open counts"""
    import oracle_lib as O
    from lpr_381_group_v22_b200.distributed import _Comm, round_status
    w, v, cap = O.gen_knapsack(9, 20)
    pool = CpuKnapPool(cap, w, v)
    comm = _Comm(None, "cpu")
    assert round_status(pool, comm, None) == (float("-inf"), False, [1])
    pool.run(10 ** 6)
    z, changed, counts = round_status(pool, comm, None)
    assert z == pool.get_incumbent()[0] and changed and counts == [0]
    assert round_status(pool, comm, pool.get_incumbent())[1] is False


def test_steal_plan_and_order():
    from lpr_381_group_v22_b200.distributed import better, steal_plan
    assert steal_plan([10, 0]) == [(0, 1, 5)]
    assert steal_plan([0, 0, 9, 3]) == [(2, 0, 4), (2, 1, 2)]
    assert steal_plan([1, 0]) == []
    assert steal_plan([4, 4]) == []
    a = (15.0, (0, 1), None)
    assert better(a, None) and not better(None, a)
    assert better((16.0, (1,), None), a)
    assert better((15.0, (0,), None), a)          # an ancestor path precedes its extensions
    assert better((15.0, (0, 0, 1), None), a) and not better((15.0, (1,), None), a)


def test_single_process_driver_equals_oracle():
    import oracle_lib as O
    from lpr_381_group_v22_b200.distributed import run_distributed
    w, v, cap = O.gen_knapsack(9, 26)
    ref = O.knap_bb(cap, w, v)
    res = run_distributed(CpuKnapPool(cap, w, v), None, "cpu", chunk_nodes=7, payload_len=26)
    assert res["incumbent"][0] == ref["best"]
    assert [int(c) for c in res["incumbent"][2]] == ref["chosen"].tolist()
    assert res["nodes_total"] == ref["nodes"]  # same pruning decisions as the sequential DFS
