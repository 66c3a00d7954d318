"""DeviceTableau: a dense simplex tableau resident in B200 HBM (thin wrapper over lpr_tab_*).

Replaces the `double[,] tableau` field of the reference solvers
(Simplex/PrimalSimplexSolver.cs:12, Simplex/PrimalSimplexSolver2.cs:12) and the
`(double[] obj, List<double[]> rows)` pair of DualSimplex.cs / CuttingPlaneSolver.cs.
"""
import ctypes as C

import numpy as np

from . import _native as N


class DeviceTableau:
    def __init__(self, handle):
        self._h = handle

    # ---- construction ---------------------------------------------------------------------
    @classmethod
    def from_host(cls, T, device=0, row_cap=0, col_cap=0):
        T = N.f64(T)
        if T.ndim != 2:
            raise ValueError("tableau must be 2-D")
        h = N.vp()
        N.check(N.lib().lpr_tab_create(device, T.shape[0], T.shape[1], row_cap, col_cap, N.pd(T), C.byref(h)))
        return cls(h)

    @classmethod
    def from_bb_model(cls, objective, constraints, device=0, row_cap=0, col_cap=0):
        """DualSimplexSolverBB.FormulateTableau (BranchBoundSimplexSolver.cs:28-113) executed on the device;
        constraints are the reference's ragged rows [coefficients..., rhs, type flag]."""
        obj = N.f64(objective)
        m = len(constraints)
        stride = max([len(c) for c in constraints] + [2])
        cons = np.zeros((max(1, m), stride))
        ln = np.zeros(max(1, m), dtype=np.int32)
        for i, c in enumerate(constraints):
            cons[i, :len(c)] = c
            ln[i] = len(c)
        h = N.vp()
        N.check(N.lib().lpr_tab_create_bb(device, len(obj), m, N.pd(obj), N.pd(cons), stride, N.pi(ln), row_cap, col_cap,
                                          C.byref(h)))
        return cls(h)

    @classmethod
    def from_model(cls, objective, constraints, is_maximization=True, device=0):
        """PrimalSimplexSolver..ctor (PrimalSimplexSolver.cs:27-87) executed on the device."""
        n = len(objective)
        m = len(constraints)
        stride = max([len(c.Coefficients) for c in constraints] + [n, 1])
        coef = np.zeros((m, stride))
        cnt = np.zeros(m, dtype=np.int32)
        rel = np.zeros(m, dtype=np.int32)
        rhs = np.zeros(m)
        for i, c in enumerate(constraints):
            k = len(c.Coefficients)
            coef[i, :k] = c.Coefficients
            cnt[i] = k
            rel[i] = N.REL.get(c.Relation, 0)  # anything but ">=" is treated as "<=" (:42-50)
            rhs[i] = c.RHS
        obj = N.f64(objective)
        h = N.vp()
        N.check(N.lib().lpr_tab_create_primal(device, n, m, N.pd(obj), N.pd(coef), stride, N.pi(cnt), N.pi(rel),
                                              N.pd(rhs), int(bool(is_maximization)), C.byref(h)))
        return cls(h)

    @classmethod
    def dense_lp(cls, seed, m, n, device=0):
        """Synthetic dense LP of SURVEY.md 8(d), generated directly in HBM."""
        h = N.vp()
        N.check(N.lib().lpr_tab_create_dense_lp(device, seed, m, n, C.byref(h)))
        return cls(h)

    def close(self):
        if self._h is not None:
            N.lib().lpr_tab_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # ---- data movement ----------------------------------------------------------------------
    @property
    def shape(self):
        r, c, ld = C.c_int(), C.c_int(), C.c_int()
        N.check(N.lib().lpr_tab_dims(self._h, C.byref(r), C.byref(c), C.byref(ld)))
        return r.value, c.value

    @property
    def ld(self):
        r, c, ld = C.c_int(), C.c_int(), C.c_int()
        N.check(N.lib().lpr_tab_dims(self._h, C.byref(r), C.byref(c), C.byref(ld)))
        return ld.value

    def read(self, out=None):
        r, c = self.shape
        if out is None:
            out = np.empty((r, c))
        N.check(N.lib().lpr_tab_read(self._h, N.pd(out)))
        return out

    def upload(self, T):
        T = N.f64(T)
        if T.shape != self.shape:
            raise ValueError("shape mismatch")
        N.check(N.lib().lpr_tab_upload(self._h, N.pd(T)))

    def read_row(self, i):
        out = np.empty(self.shape[1])
        N.check(N.lib().lpr_tab_read_row(self._h, i, N.pd(out)))
        return out

    def read_col(self, j):
        out = np.empty(self.shape[0])
        N.check(N.lib().lpr_tab_read_col(self._h, j, N.pd(out)))
        return out

    @property
    def basis(self):
        b = np.zeros(max(1, self.shape[0] - 1), dtype=np.int32)
        N.check(N.lib().lpr_tab_get_basis(self._h, N.pi(b)))
        return b[:self.shape[0] - 1]

    @basis.setter
    def basis(self, b):
        b = N.i32(b)
        if len(b) != self.shape[0] - 1:
            raise ValueError("basis needs rows-1 entries")
        N.check(N.lib().lpr_tab_set_basis(self._h, N.pi(b)))

    # ---- solving ----------------------------------------------------------------------------
    def solve(self, rule=N.RULE_PRIMAL, max_pivots=-1, print_steps=False, log_cap=1 << 16, fused=True, blocked=True,
              time_sweeps=False, pipelined=True):
        st = C.c_int()
        npv = C.c_int64()
        log = np.zeros((max(1, log_cap), 2), dtype=np.int32)
        flags = (1 if print_steps else 0) | (0 if fused else 4) | (0 if blocked else 16) | (8 if time_sweeps else 0) | \
            (0 if pipelined else 32)
        N.check(N.lib().lpr_tab_solve(self._h, rule, max_pivots, flags, C.byref(st), C.byref(npv),
                                      N.pi(log) if log_cap > 0 else None, log_cap))
        return dict(status=st.value, n_pivots=npv.value, log=log[:min(npv.value, log_cap)].copy())

    def step(self, rule=N.RULE_PRIMAL):
        e, l, st = C.c_int(), C.c_int(), C.c_int()
        N.check(N.lib().lpr_tab_step(self._h, rule, C.byref(e), C.byref(l), C.byref(st)))
        return e.value, l.value, st.value

    def pivot_at(self, row, col, skip_eps=0.0, skip_mode=0):
        N.check(N.lib().lpr_tab_pivot_at(self._h, row, col, skip_eps, skip_mode))

    def extract_solution(self, n):
        x = np.zeros(n)
        N.check(N.lib().lpr_tab_extract_solution(self._h, n, N.pd(x)))
        return x

    def objective(self):
        z = C.c_double()
        N.check(N.lib().lpr_tab_objective(self._h, C.byref(z)))
        return z.value

    @property
    def last_solve_ms(self):
        ms = C.c_float()
        N.check(N.lib().lpr_tab_last_solve_ms(self._h, C.byref(ms)))
        return ms.value

    @property
    def last_sweep_us(self):
        us = C.c_float()
        N.check(N.lib().lpr_tab_last_sweep_us(self._h, C.byref(us)))
        return us.value

    # ---- SensitivityAnalyzer building blocks (SensitivityAnalyzer.cs:609-723) ----------------------------
    def sens_rebuild_basis(self):
        N.check(N.lib().lpr_tab_sens_rebuild_basis(self._h))

    def sens_solution(self):
        x = np.zeros(max(1, self.shape[1] - 1))
        N.check(N.lib().lpr_tab_sens_solution(self._h, N.pd(x)))
        return x[:self.shape[1] - 1]

    def sens_add_constraint(self, tech, rhs_minus_ax):
        tech = N.f64(tech)
        if tech.shape != (self.shape[1] - 1,):
            raise ValueError("tech must have cols-1 entries")
        N.check(N.lib().lpr_tab_sens_add_constraint(self._h, N.pd(tech), float(rhs_minus_ax)))

    # ---- cutting plane / B&B building blocks ------------------------------------------------------
    def append_row(self, row):
        row = N.f64(row)
        N.check(N.lib().lpr_tab_append_row(self._h, N.pd(row)))

    def gomory_cut(self, append=False):
        ch = C.c_int()
        cut = np.zeros(self.shape[1])
        N.check(N.lib().lpr_tab_gomory_cut(self._h, C.byref(ch), N.pd(cut), int(append)))
        return ch.value, cut

    def cutting_plane(self, max_cuts=-1, log_cap=256):
        st, nc = C.c_int(), C.c_int()
        log = np.zeros((log_cap, 4), dtype=np.int32)
        N.check(N.lib().lpr_tab_cutting_plane(self._h, max_cuts, C.byref(st), C.byref(nc), N.pi(log), log_cap))
        return dict(status=st.value, n_cuts=nc.value, log=log[:min(nc.value, log_cap)].copy())

    def round4(self):
        N.check(N.lib().lpr_tab_round4(self._h))

    def bb_node_solve(self, max_pivots=-1, log_cap=4096, is_min=False):
        st = C.c_int()
        npv = C.c_int64()
        log = np.zeros((log_cap, 2), dtype=np.int32)
        N.check(N.lib().lpr_tab_bb_node_solve_ex(self._h, int(bool(is_min)), max_pivots, C.byref(st), C.byref(npv),
                                                 N.pi(log), log_cap))
        return dict(status=st.value, n_pivots=npv.value, log=log[:min(npv.value, log_cap)].copy())

    def bb_add_constraint(self, n_vars, var, bound, typ):
        h = N.vp()
        N.check(N.lib().lpr_tab_bb_add_constraint(self._h, n_vars, var, float(bound), typ, C.byref(h)))
        return DeviceTableau(h)

    def bb_branch_var(self, n_vars):
        var = C.c_int()
        val = C.c_double()
        x = np.zeros(n_vars)
        N.check(N.lib().lpr_tab_bb_branch_var(self._h, n_vars, C.byref(var), C.byref(val), N.pd(x)))
        return var.value, val.value, x
