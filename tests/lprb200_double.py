"""TEST DOUBLE of the COMPUTE entry points of the C ABI (include/lprb200.h), backed by the CPU oracle.

Purpose, and the only use: tests/test_csharp_shims.py executes the C# P/Invoke shims under csharp/ with the interpreter
of oracle/csharp/.  Host-only entry points (model parser, formatters, result files) go to the REAL liblprb200.so, which
needs no GPU; the compute entry points need a B200, so in the CPU suite they are answered by this double, which follows
the semantics include/lprb200.h documents for each call.  On a GPU box the same shim tests run against the real
library for everything (tests/test_csharp_shims_gpu.py).  This module is not a fallback of the product: nothing under
lpr_381_group_v22_b200/ can reach it, it has no C symbols, and it is never loaded outside that one test file.
"""
import ctypes as C

import numpy as np

import net_reference as R
import oracle_lib as O

OK, E_INVALID = 0, 1
RUNNING, OPTIMAL, UNBOUNDED, INFEASIBLE, ITER_LIMIT, NODE_LIMIT, PIVOT_TOO_SMALL = range(7)
RULE_PRIMAL, RULE_PRIMAL2, RULE_DUAL, RULE_SENS = range(4)


class _Tab:
    def __init__(self, T, row_cap=0, col_cap=0, basis=None):
        self.T = np.array(T, dtype=np.float64)
        self.row_cap = max(row_cap, self.T.shape[0])
        self.col_cap = max(col_cap, self.T.shape[1])
        R_, C_ = self.T.shape
        self.basis = np.array(basis if basis is not None else np.arange(C_ - R_, C_ - 1), dtype=np.int32)


class _Rev:
    def __init__(self, A, b, c, is_min):
        self.A, self.b, self.c, self.is_min = A, b, c, bool(is_min)
        self.k = 0
        self.state = None
        self.snaps = None


class OracleBackedDouble:
    """call(name, params, values, return_type) like oracle/csharp/pinvoke.NativeLibrary"""

    def __init__(self, real):
        self.real = real            # NativeLibrary over the real liblprb200.so: host-only entry points
        self.handles = {}
        self.next = 0x10000
        self.keep = []
        self.calls = []
        self.error = C.create_string_buffer(b"", 512)
        self.failed_last = False

    def call(self, name, params, values, rettype):
        if name == "lpr_last_error" and not self.failed_last:
            return self.real.call(name, params, values, rettype)      # the message of the real library's last failure
        fn = getattr(self, name, None)
        if fn is None:
            self.failed_last = False
            return self.real.call(name, params, values, rettype)
        self.calls.append(name)
        if name != "lpr_last_error":
            self.failed_last = False
        try:
            r = fn(*values)
        except KeyError:
            return self._fail("invalid handle")
        return OK if r is None else r

    # ---- helpers
    def _fail(self, msg):
        self.error.value = msg.encode("utf-8")
        self.failed_last = True
        return E_INVALID

    def _new(self, obj):
        self.next += 16
        self.handles[self.next] = obj
        return self.next

    def _text(self, s, text_ref, len_ref):
        raw = s.encode("utf-8")
        buf = C.create_string_buffer(raw, len(raw) + 1)
        self.keep.append(buf)
        text_ref.set(C.addressof(buf))
        len_ref.set(len(raw))

    @staticmethod
    def _mat(arr, rows, cols):
        return np.array(arr.data, dtype=np.float64).reshape(rows, cols)

    def lpr_last_error(self):
        return C.addressof(self.error)

    # ---- dense tableau
    def lpr_tab_create(self, device, rows, cols, row_cap, col_cap, host, out):
        out.set(self._new(_Tab(self._mat(host, rows, cols), row_cap, col_cap)))

    def lpr_tab_create_primal(self, device, n, m, objective, coef, stride, count, relation, rhs, is_max, out):
        rel = {0: "<=", 1: ">=", 2: "="}
        cons = [(coef.data[i * stride:i * stride + count.data[i]], rel[relation.data[i]], rhs.data[i]) for i in range(m)]
        T, basis = O.primal_build(list(objective.data), cons, bool(is_max))
        out.set(self._new(_Tab(T, basis=basis)))

    def lpr_tab_destroy(self, h):
        if h:
            del self.handles[h]

    def lpr_tab_dims(self, h, rows, cols, ld):
        t = self.handles[h]
        rows.set(t.T.shape[0]); cols.set(t.T.shape[1]); ld.set((t.T.shape[1] + 15) // 16 * 16)

    def lpr_tab_read(self, h, host):
        t = self.handles[h]
        if len(host.data) != t.T.size:
            return self._fail("lpr_tab_read: host array has the wrong size")
        host.data[:] = t.T.ravel().tolist()

    def lpr_tab_get_basis(self, h, basis):
        t = self.handles[h]
        basis.data[:len(t.basis)] = [int(v) for v in t.basis]

    def lpr_tab_objective(self, h, z):
        z.set(float(self.handles[h].T[0, -1]))

    def lpr_tab_extract_solution(self, h, n, x):
        x.data[:n] = O.primal_extract(self.handles[h].T, n).tolist()

    def _solve(self, t, rule, max_pivots, flags):
        if rule == RULE_PRIMAL:
            r = O.primal_solve(t.T, t.basis, max_pivots=max_pivots)
            t.basis = r["basis"]
        elif rule == RULE_PRIMAL2:
            r = O.primal2_solve(t.T, int(max_pivots), bool(flags & 1))
        elif rule == RULE_DUAL:
            r = O.dual_solve(t.T, int(max_pivots), bool(flags & 1))
        else:
            r = O.sens_resolve(t.T, t.basis, int(max_pivots))
            t.basis = r["basis"]
        t.T = r["T"]
        return r

    def lpr_tab_solve(self, h, rule, max_pivots, flags, status, n_pivots, log, log_cap):
        r = self._solve(self.handles[h], rule, max_pivots, flags)
        status.set(int(r["status"])); n_pivots.set(int(r["n_pivots"]))
        if log is not None:
            flat = r["log"][:log_cap].ravel().tolist()
            log.data[:len(flat)] = flat

    def lpr_tab_step(self, h, rule, enter, leave, status):
        t = self.handles[h]
        if rule == RULE_PRIMAL:
            r = self._solve(t, rule, 1, 0)
        else:
            r = self._solve(t, rule, 1, 1)       # with printSteps the iteration counter advances: one pivot, then the cap
        if r["n_pivots"] == 1:
            enter.set(int(r["log"][0, 1])); leave.set(int(r["log"][0, 0])); status.set(RUNNING)
        else:
            enter.set(-1); leave.set(-1); status.set(int(r["status"]))

    def lpr_tab_cutting_plane(self, h, max_cuts, status, n_cuts, log, cap):
        t = self.handles[h]
        r = O.cutting_plane(t.T, max_cuts=max_cuts, extra_rows=t.row_cap - t.T.shape[0])
        t.T = r["T"]
        status.set(int(r["status"])); n_cuts.set(int(r["n_cuts"]))
        flat = r["log"][:cap].ravel().tolist()
        log.data[:len(flat)] = flat

    # ---- sensitivity
    def lpr_tab_sens_rebuild_basis(self, h):
        t = self.handles[h]
        t.basis = O.sens_rebuild_basis(t.T)

    def lpr_tab_sens_solution(self, h, x):
        v = O.sens_solution(self.handles[h].T)
        x.data[:len(v)] = v.tolist()

    def lpr_tab_sens_add_constraint(self, h, tech, rhs_minus_ax):
        t = self.handles[h]
        if t.T.shape[0] + 1 > t.row_cap or t.T.shape[1] + 1 > t.col_cap:
            return self._fail("lpr_tab_sens_add_constraint: no headroom")
        t.T, t.basis = O.sens_add_constraint(t.T, t.basis, tech.data[:t.T.shape[1] - 1], rhs_minus_ax)

    # ---- B&B building blocks
    def lpr_tab_create_bb(self, device, n, m, objective, cons, stride, length, row_cap, col_cap, out):
        rows = [cons.data[i * stride:i * stride + length.data[i]] for i in range(m)]
        out.set(self._new(_Tab(O.bb_formulate(list(objective.data), rows), row_cap, col_cap)))

    def lpr_tab_bb_node_solve_ex(self, h, is_min, max_pivots, status, n_pivots, log, cap):
        t = self.handles[h]
        r = O.bb_node_solve_ex(t.T, bool(is_min), max_pivots=max_pivots)
        t.T = r["T"]
        status.set(int(r["status"])); n_pivots.set(int(r["n_pivots"]))
        if log is not None:
            flat = r["log"][:cap].ravel().tolist()
            log.data[:len(flat)] = flat

    def lpr_tab_round4(self, h):
        t = self.handles[h]
        t.T = O.bb_round(t.T)

    def lpr_tab_bb_add_constraint(self, parent, n_vars, var, bound, typ, child):
        child.set(self._new(_Tab(O.bb_add_constraint(self.handles[parent].T, n_vars, var, bound, typ))))

    def lpr_bb_solve(self, device, rows, cols, T, n_vars, prune, max_nodes, x, z, has, nodes, pivots, node_log, node_z,
                     cap, status):
        r = O.bb_solve(self._mat(T, rows, cols), n_vars, prune=bool(prune), max_nodes=max_nodes)
        x.data[:n_vars] = r["x"].tolist()
        z.set(float(r["z"])); has.set(int(r["has_solution"])); nodes.set(int(r["nodes"])); pivots.set(int(r["pivots"]))
        status.set(int(r["status"]))

    # ---- revised simplex
    def lpr_rev_create(self, device, m, n, A, b, c, is_min, out):
        out.set(self._new(_Rev(self._mat(A, m, n), list(b.data), list(c.data), is_min)))

    def lpr_rev_destroy(self, h):
        if h:
            del self.handles[h]

    def _rev_run(self, r, max_iter):
        r.state = O.rev_solve(r.A, r.b, r.c, r.is_min, max_iter=max_iter)
        return r.state

    def lpr_rev_solve(self, h, max_iter, refactor_every, status, n_iter, log, cap):
        st = self._rev_run(self.handles[h], max_iter)
        status.set(int(st["status"])); n_iter.set(int(st["n_iter"]))

    def lpr_rev_begin(self, h):
        r = self.handles[h]
        r.k = 0
        r.snaps = None

    def lpr_rev_step(self, h, status, enter, leave_row, leave_var):
        r = self.handles[h]
        st = self._rev_run(r, r.k + 1)
        if st["n_iter"] == r.k + 1:
            lr, e, lv = st["log"][r.k].tolist()
            r.k += 1
            status.set(RUNNING); enter.set(e); leave_row.set(lr); leave_var.set(lv)
        else:
            status.set(int(st["status"])); enter.set(-1); leave_row.set(-1); leave_var.set(-1)
            r.k = -1 if st["status"] == OPTIMAL else r.k

    def lpr_rev_format_snapshot(self, h, text, length):
        r = self.handles[h]
        if r.snaps is None:
            r.snaps = R.revised_solve_with_snapshots(r.c, r.A.tolist(), r.b, r.is_min)[0]
        self._text(r.snaps[-1] if r.k < 0 else r.snaps[r.k - 1], text, length)

    def lpr_rev_read_basis(self, h, basis):
        st = self.handles[h].state
        basis.data[:len(st["basis"])] = [int(v) for v in st["basis"]]

    def lpr_rev_read_x(self, h, x):
        x.data[:] = self.handles[h].state["x"].tolist()

    def lpr_rev_read_z(self, h, z):
        z.set(float(self.handles[h].state["z"]))

    def lpr_rev_read_y(self, h, y):
        y.data[:] = self.handles[h].state["y"].tolist()

    # ---- knapsack
    def lpr_knap_solve(self, device, capacity, n, weights, values, max_nodes, best, chosen, nodes, status):
        r = O.knap_bb(capacity, weights.data, values.data, max_nodes)
        best.set(float(r["best"])); nodes.set(int(r["nodes"])); status.set(int(r["status"]))
        chosen.data[:] = [int(v) for v in r["chosen"]]

    def lpr_knap_dp(self, device, capacity, n, weights, values, best, chosen):
        b, ch = O.knap_dp(capacity, weights.data, values.data)
        best.set(float(b))
        chosen.data[:] = [int(v) for v in ch]
