"""ctypes binding of the CPU oracle (oracle/liblpr_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_DIR = os.path.join(os.path.dirname(_HERE), "oracle")
_LIB = os.path.join(ORACLE_DIR, "liblpr_oracle.so")

RUNNING, OPTIMAL, UNBOUNDED, INFEASIBLE, ITER_LIMIT, NODE_LIMIT, PIVOT_TOO_SMALL, NO_CUT_NEEDED, \
    NO_PIVOT_COL, CUT_STEP_DONE = range(10)

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_lp = C.POINTER(C.c_int64)
_bp = C.POINTER(C.c_uint8)


def build():
    src = os.path.join(ORACLE_DIR, "lpr_oracle.cpp")
    if (not os.path.exists(_LIB)) or os.path.getmtime(_LIB) < max(
            os.path.getmtime(src), os.path.getmtime(os.path.join(ORACLE_DIR, "lpr_oracle.h"))):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "-s"])
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.orc_u01.restype = C.c_double
        _lib.orc_u01.argtypes = [C.c_uint64, C.c_uint64]
        _lib.orc_splitmix64.restype = C.c_uint64
        _lib.orc_splitmix64.argtypes = [C.c_uint64]
        for f in ("orc_net_round", "orc_net_round4", "orc_frac"):
            getattr(_lib, f).restype = C.c_double
            getattr(_lib, f).argtypes = [C.c_double]
        _lib.orc_knap_dp.restype = C.c_double
        _lib.orc_knap_dp_value.restype = C.c_double
        _lib.orc_knap_bb.restype = C.c_double
        _lib.orc_knap_bb.argtypes = [C.c_double, C.c_int, _dp, _dp, C.c_int64, _bp, _lp, _ip]
        _lib.orc_gen_dense_lp.argtypes = [C.c_uint64, C.c_int, C.c_int, _dp, _dp, _dp]
        _lib.orc_gen_dense_ip.argtypes = [C.c_uint64, C.c_int, C.c_int, _dp, _dp, _dp]
        _lib.orc_gen_knapsack.argtypes = [C.c_uint64, C.c_int, _dp, _dp, _dp]
    return _lib


def _d(a):
    return a.ctypes.data_as(_dp)


def _i(a):
    return a.ctypes.data_as(_ip)


def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


# ---------------------------------------------------------------- generators
def gen_dense_lp(seed, m, n):
    A = np.empty((m, n)); b = np.empty(m); c = np.empty(n)
    lib().orc_gen_dense_lp(seed, m, n, _d(A), _d(b), _d(c))
    return A, b, c


def gen_dense_ip(seed, m, n):
    A = np.empty((m, n)); b = np.empty(m); c = np.empty(n)
    lib().orc_gen_dense_ip(seed, m, n, _d(A), _d(b), _d(c))
    return A, b, c


def gen_knapsack(seed, n):
    w = np.empty(n); v = np.empty(n); cap = C.c_double()
    lib().orc_gen_knapsack(seed, n, _d(w), _d(v), C.byref(cap))
    return w, v, cap.value


# ---------------------------------------------------------------- primal tableau
REL = {"<=": 0, ">=": 1, "=": 2}


def primal_build(objective, constraints, is_max=True):
    """constraints: list of (coefficients, relation, rhs) like InputFileParser.Constraint."""
    n = len(objective); m = len(constraints)
    stride = max([len(c[0]) for c in constraints] + [n, 1])
    coef = np.zeros((m, stride)); cnt = np.zeros(m, dtype=np.int32)
    rel = np.zeros(m, dtype=np.int32); rhs = np.zeros(m)
    for i, (co, r, b) in enumerate(constraints):
        coef[i, :len(co)] = co; cnt[i] = len(co); rel[i] = REL[r]; rhs[i] = b
    T = np.empty((m + 1, n + m + 1)); basis = np.empty(m, dtype=np.int32)
    obj = f64(objective)
    lib().orc_primal_build(n, m, _d(obj), _d(coef), stride, _i(cnt), _i(rel), _d(rhs), int(is_max), _d(T), _i(basis))
    return T, basis


def primal_solve(T, basis=None, max_pivots=-1, threads=1, log_cap=1 << 16):
    T = f64(T).copy(); R, Cc = T.shape
    if basis is None:
        basis = np.arange(Cc - R, Cc - 1, dtype=np.int32)
    basis = np.ascontiguousarray(basis, dtype=np.int32).copy()
    st = C.c_int(); npv = C.c_int64(); log = np.zeros((log_cap, 2), dtype=np.int32)
    lib().orc_primal_solve(R, Cc, _d(T), _i(basis), C.c_int64(max_pivots), C.byref(st), C.byref(npv), _i(log),
                           C.c_int64(log_cap), threads)
    return dict(T=T, basis=basis, status=st.value, n_pivots=npv.value, log=log[:min(npv.value, log_cap)].copy())


def primal_extract(T, n):
    T = f64(T); R, Cc = T.shape; x = np.empty(n)
    lib().orc_primal_extract(R, Cc, n, _d(T), _d(x))
    return x


def _rule_solve(fn, T, max_iters, print_steps, log_cap=1 << 14):
    T = f64(T).copy(); R, Cc = T.shape
    st = C.c_int(); npv = C.c_int64(); log = np.zeros((log_cap, 2), dtype=np.int32)
    fn(R, Cc, _d(T), max_iters, int(print_steps), C.byref(st), C.byref(npv), _i(log), C.c_int64(log_cap))
    return dict(T=T, status=st.value, n_pivots=npv.value, log=log[:min(npv.value, log_cap)].copy())


def primal2_solve(T, max_iters=10000, print_steps=False):
    return _rule_solve(lib().orc_primal2_solve, T, max_iters, print_steps)


def dual_solve(T, max_iters=10000, print_steps=True):
    return _rule_solve(lib().orc_dual_solve, T, max_iters, print_steps)


def sens_resolve(T, basis, max_iter=10000, log_cap=1 << 14):
    T = f64(T).copy(); R, Cc = T.shape
    basis = np.ascontiguousarray(basis, dtype=np.int32).copy()
    st = C.c_int(); npv = C.c_int64(); log = np.zeros((log_cap, 2), dtype=np.int32)
    lib().orc_sens_resolve(R, Cc, _d(T), _i(basis), max_iter, C.byref(st), C.byref(npv), _i(log), C.c_int64(log_cap))
    return dict(T=T, basis=basis, status=st.value, n_pivots=npv.value, log=log[:npv.value].copy())


def sens_rebuild_basis(T):
    T = f64(T); R, Cc = T.shape
    basis = np.zeros(max(1, R - 1), dtype=np.int32)
    lib().orc_sens_rebuild_basis(R, Cc, _d(T), _i(basis))
    return basis[:R - 1]


def sens_solution(T):
    T = f64(T); R, Cc = T.shape
    x = np.zeros(max(1, Cc - 1))
    lib().orc_sens_solution(R, Cc, _d(T), _d(x))
    return x[:Cc - 1]


def sens_add_constraint(T, basis, tech, rhs_minus_ax):
    T = f64(T); R, Cc = T.shape
    b = np.zeros(R, dtype=np.int32); b[:R - 1] = basis
    tech = f64(tech)
    out = np.zeros((R + 1, Cc + 1))
    lib().orc_sens_add_constraint(R, Cc, _d(T), _i(b), _d(tech), C.c_double(rhs_minus_ax), _d(out))
    return out, b


# ---------------------------------------------------------------- cutting plane
def gomory_cut(T, literal_sort=False):
    """literal_sort: see cutting_plane"""
    T = f64(T); R, Cc = T.shape; cut = np.zeros(Cc)
    lib().orc_set_gomory_first_min(0 if literal_sort else 1)
    row = lib().orc_gomory_cut(R, Cc, _d(T), _d(cut))
    lib().orc_set_gomory_first_min(0)
    return row, cut


def cutting_plane(T, max_cuts=-1, extra_rows=64, literal_sort=False):
    """literal_sort=True: the cut row is element 0 of the reference's List<T>.Sort (the Framework's unstable
    introspective sort), which is what the executed reference does (tests/test_reference_run.py).  The default is the
    plain first minimum, which is what the CUDA kernels implement and what the GPU parity tests compare with: the two
    differ only when more than 16 fractional rows tie exactly for the best key."""
    T = f64(T); R, Cc = T.shape
    cap = R + (extra_rows if max_cuts < 0 else max_cuts + 1)
    buf = np.zeros((cap, Cc)); buf[:R] = T
    Rio = C.c_int(R); st = C.c_int(); nc = C.c_int(); log = np.zeros((cap, 4), dtype=np.int32)
    lib().orc_set_gomory_first_min(0 if literal_sort else 1)
    lib().orc_cutting_plane(C.byref(Rio), Cc, _d(buf), cap, max_cuts, C.byref(st), C.byref(nc), _i(log), cap)
    lib().orc_set_gomory_first_min(0)
    return dict(T=buf[:Rio.value].copy(), status=st.value, n_cuts=nc.value, log=log[:nc.value].copy())


def gomory_tie_corners():
    """how many cut-row choices so far had the literal sort and the first minimum disagree"""
    lib().orc_gomory_tie_corners.restype = C.c_int64
    return lib().orc_gomory_tie_corners()


# ---------------------------------------------------------------- branch & bound simplex
def bb_round(T):
    T = f64(T).copy()
    lib().orc_bb_round_tableau(C.c_int64(T.size), _d(T))
    return T


def bb_dual_pivot(T):
    T = f64(T); out = np.zeros_like(T); r = C.c_int(-1); c = C.c_int(-1)
    ok = lib().orc_bb_dual_pivot(T.shape[0], T.shape[1], _d(T), _d(out), C.byref(r), C.byref(c))
    return ok, out, r.value, c.value


def bb_primal_pivot(T):
    T = f64(T); out = np.zeros_like(T); r = C.c_int(-1); c = C.c_int(-1)
    ok = lib().orc_bb_primal_pivot(T.shape[0], T.shape[1], _d(T), _d(out), C.byref(r), C.byref(c))
    return ok, out, r.value, c.value


def bb_node_solve(T, max_pivots=-1, log_cap=4096):
    T = f64(T).copy(); npv = C.c_int64(); log = np.zeros((log_cap, 2), dtype=np.int32)
    res = lib().orc_bb_node_solve(T.shape[0], T.shape[1], _d(T), C.c_int64(max_pivots), C.byref(npv), _i(log),
                                  C.c_int64(log_cap))
    return dict(status=res, T=T, n_pivots=npv.value, log=log[:min(npv.value, log_cap)].copy())


def bb_node_solve_ex(T, is_min, max_pivots=-1, log_cap=4096):
    T = f64(T).copy(); R, Cc = T.shape
    npv = C.c_int64(); log = np.zeros((log_cap, 2), dtype=np.int32)
    st = lib().orc_bb_node_solve_ex(R, Cc, _d(T), int(bool(is_min)), C.c_int64(max_pivots), C.byref(npv), _i(log),
                                    C.c_int64(log_cap))
    return dict(T=T, status=st, n_pivots=npv.value, log=log[:min(npv.value, log_cap)].copy())


def bb_formulate(objective, constraints):
    """FormulateTableau :28-113 on the reference's ragged rows [coefficients..., rhs, type flag]"""
    obj = f64(objective); n = len(obj); m = len(constraints)
    stride = max([len(c) for c in constraints] + [2])
    cons = np.zeros((max(1, m), stride)); ln = np.zeros(max(1, m), dtype=np.int32)
    for i, c in enumerate(constraints):
        cons[i, :len(c)] = c; ln[i] = len(c)
    T = np.zeros((m + 1, n + m + 1))
    lib().orc_bb_formulate(n, m, _d(obj), _d(cons), stride, _i(ln), _d(T))
    return T


def bb_configure_problem(objective, constraints):
    """ConfigureProblem :1233-1251: one row x_i <= 1 per variable, each ONE ENTRY LONGER than a normal row"""
    n = len(objective)
    cons = [list(c) for c in constraints]
    for i in range(n):
        row = [0.0] * (n + 3)
        row[i] = 1.0
        row[n + 1] = 1.0
        cons.append(row)
    return list(objective), cons


def bb_identify_basic(T):
    T = f64(T); basic = np.zeros(T.shape[1], dtype=np.int32)
    k = lib().orc_bb_identify_basic(T.shape[0], T.shape[1], _d(T), _i(basic))
    return basic[:k].copy()


def bb_add_constraint(T, n_vars, var, bound, typ):
    T = f64(T); R, Cc = T.shape; out = np.zeros((R + 1, Cc + 1))
    lib().orc_bb_add_constraint(R, Cc, _d(T), n_vars, var, C.c_double(bound), typ, _d(out))
    return out


def bb_branch_var(T, n_vars):
    T = f64(T); val = C.c_double(0.0)
    v = lib().orc_bb_branch_var(T.shape[0], T.shape[1], _d(T), n_vars, C.byref(val))
    return v, val.value


def bb_extract(T, n_vars):
    T = f64(T); x = np.zeros(n_vars)
    lib().orc_bb_extract(T.shape[0], T.shape[1], _d(T), n_vars, _d(x))
    return x


def bb_solve(T, n_vars, prune=False, max_nodes=20, log_cap=1 << 16):
    T = f64(T); R, Cc = T.shape
    x = np.zeros(n_vars); z = C.c_double(); has = C.c_int(); nodes = C.c_int64(); piv = C.c_int64()
    nlog = np.zeros((log_cap, 4), dtype=np.int32); nz = np.zeros(log_cap)
    st = lib().orc_bb_solve(R, Cc, _d(T), n_vars, int(prune), C.c_int64(max_nodes), _d(x), C.byref(z),
                            C.byref(has), C.byref(nodes), C.byref(piv), _i(nlog), _d(nz), C.c_int64(log_cap))
    k = min(nodes.value, log_cap)
    return dict(status=st, x=x, z=z.value, has_solution=bool(has.value), nodes=nodes.value, pivots=piv.value,
                node_log=nlog[:k].copy(), node_z=nz[:k].copy())


# ---------------------------------------------------------------- revised simplex
def rev_solve(A, b, c, is_min=False, max_iter=-1, want_binv=False, log_cap=1 << 16):
    A = f64(A); b = f64(b); c = f64(c); m, n = A.shape
    basis = np.zeros(m, dtype=np.int32); x = np.zeros(n); z = C.c_double(); y = np.zeros(m); xB = np.zeros(m)
    Binv = np.zeros((m, m)) if want_binv else None
    nit = C.c_int64(); log = np.zeros((log_cap, 3), dtype=np.int32)
    st = lib().orc_rev_solve(m, n, _d(A), _d(b), _d(c), int(is_min), C.c_int64(max_iter), C.byref(nit), _i(basis),
                             _d(x), C.byref(z), _d(y), _d(xB), _d(Binv) if want_binv else None, _i(log),
                             C.c_int64(log_cap))
    return dict(status=st, n_iter=nit.value, basis=basis, x=x, z=z.value, y=y, xB=xB, Binv=Binv,
                log=log[:min(nit.value, log_cap)].copy())


# ---------------------------------------------------------------- knapsack
def knap_dp(capacity, weights, values):
    w = np.ascontiguousarray(weights, dtype=np.int32); v = np.ascontiguousarray(values, dtype=np.int32)
    ch = np.zeros(len(w), dtype=np.uint8)
    best = lib().orc_knap_dp(int(capacity), len(w), _i(w), _i(v), ch.ctypes.data_as(_bp))
    return best, ch


def knap_dp_value(capacity, weights, values):
    """value-only DP with O(capacity) memory: the arbiter at BASELINE cfg4's size"""
    w = np.ascontiguousarray(weights, dtype=np.int32); v = np.ascontiguousarray(values, dtype=np.int32)
    return lib().orc_knap_dp_value(int(capacity), len(w), _i(w), _i(v))


def knap_bb(capacity, weights, values, max_nodes=-1):
    w = f64(weights); v = f64(values); ch = np.zeros(len(w), dtype=np.uint8)
    nodes = C.c_int64(); st = C.c_int()
    best = lib().orc_knap_bb(C.c_double(capacity), len(w), _d(w), _d(v), C.c_int64(max_nodes),
                             ch.ctypes.data_as(_bp), C.byref(nodes), C.byref(st))
    return dict(best=best, chosen=ch, nodes=nodes.value, status=st.value)
