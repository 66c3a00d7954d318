"""ctypes binding of liblprb200.so (the C ABI declared in include/lprb200.h).

This is the only place the package touches native code.  There is NO CPU fallback: if the
shared library is missing the import of any solver raises, and without a CUDA device every
compute entry point returns LPR_E_CUDA which is raised here as LprError.
"""
import ctypes as C
import os

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "liblprb200.so")

OK = 0
RUNNING, OPTIMAL, UNBOUNDED, INFEASIBLE, ITER_LIMIT, NODE_LIMIT, PIVOT_TOO_SMALL, NO_CUT_NEEDED, \
    NO_PIVOT_COL, CUT_STEP_DONE, DEPTH_LIMIT = range(11)
STATUS_NAMES = ["running", "optimal", "unbounded", "infeasible", "iter_limit", "node_limit",
                "pivot_too_small", "no_cut_needed", "no_pivot_col", "cut_step_done", "depth_limit"]
RULE_PRIMAL, RULE_PRIMAL2, RULE_DUAL, RULE_SENS = range(4)
REL = {"<=": 0, ">=": 1, "=": 2}

dp = C.POINTER(C.c_double)
ip = C.POINTER(C.c_int)
lp = C.POINTER(C.c_int64)
bp = C.POINTER(C.c_uint8)
u64p = C.POINTER(C.c_uint64)
vp = C.c_void_p


class MgpuStats(C.Structure):
    """lpr_mgpu_stats of include/lprb200.h"""
    _fields_ = [("n_gpus", C.c_int), ("nccl_version", C.c_int), ("ranks_per_gpu", C.c_int), ("reserved", C.c_int),
                ("rounds", C.c_int64), ("steals", C.c_int64),
                ("nodes_moved", C.c_int64), ("open_left", C.c_int64), ("depth_overflow", C.c_int64),
                ("seconds", C.c_double), ("setup_seconds", C.c_double), ("seed_seconds", C.c_double),
                ("exchange_seconds", C.c_double), ("steal_seconds", C.c_double), ("nodes_per_gpu", C.c_int64 * 16),
                ("run_seconds_per_gpu", C.c_double * 16)]

    def as_dict(self):
        n = self.n_gpus
        return dict(n_gpus=n, nccl_version=self.nccl_version, ranks_per_gpu=self.ranks_per_gpu, rounds=self.rounds, steals=self.steals,
                    nodes_moved=self.nodes_moved, open_left=self.open_left, depth_overflow=self.depth_overflow,
                    seconds=self.seconds, setup_seconds=self.setup_seconds, seed_seconds=self.seed_seconds,
                    exchange_seconds=self.exchange_seconds, steal_seconds=self.steal_seconds,
                    nodes_per_gpu=list(self.nodes_per_gpu[:n]), run_seconds_per_gpu=list(self.run_seconds_per_gpu[:n]))


class LprError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"liblprb200 error {code}: {msg}")
        self.code = code


_lib = None

# name -> (restype, argtypes); every symbol include/lprb200.h declares
SIGNATURES = {
    "lpr_version": (C.c_int, []),
    "lpr_last_error": (C.c_char_p, []),
    "lpr_device_count": (C.c_int, [ip]),
    "lpr_launch_count": (C.c_int64, []),
    "lpr_tab_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, dp, C.POINTER(vp)]),
    "lpr_tab_create_primal": (C.c_int, [C.c_int, C.c_int, C.c_int, dp, dp, C.c_int, ip, ip, dp, C.c_int,
                                        C.POINTER(vp)]),
    "lpr_tab_create_dense_lp": (C.c_int, [C.c_int, C.c_uint64, C.c_int, C.c_int, C.POINTER(vp)]),
    "lpr_tab_destroy": (C.c_int, [vp]),
    "lpr_tab_dims": (C.c_int, [vp, ip, ip, ip]),
    "lpr_tab_upload": (C.c_int, [vp, dp]),
    "lpr_tab_read": (C.c_int, [vp, dp]),
    "lpr_tab_read_row": (C.c_int, [vp, C.c_int, dp]),
    "lpr_tab_read_col": (C.c_int, [vp, C.c_int, dp]),
    "lpr_tab_get_basis": (C.c_int, [vp, ip]),
    "lpr_tab_set_basis": (C.c_int, [vp, ip]),
    "lpr_tab_solve": (C.c_int, [vp, C.c_int, C.c_int64, C.c_int, ip, lp, ip, C.c_int64]),
    "lpr_tab_step": (C.c_int, [vp, C.c_int, ip, ip, ip]),
    "lpr_tab_pivot_at": (C.c_int, [vp, C.c_int, C.c_int, C.c_double, C.c_int]),
    "lpr_tab_extract_solution": (C.c_int, [vp, C.c_int, dp]),
    "lpr_tab_objective": (C.c_int, [vp, dp]),
    "lpr_tab_last_solve_ms": (C.c_int, [vp, C.POINTER(C.c_float)]),
    "lpr_tab_last_sweep_us": (C.c_int, [vp, C.POINTER(C.c_float)]),
    "lpr_tab_sens_rebuild_basis": (C.c_int, [vp]),
    "lpr_tab_sens_solution": (C.c_int, [vp, dp]),
    "lpr_tab_sens_add_constraint": (C.c_int, [vp, dp, C.c_double]),
    "lpr_tab_append_row": (C.c_int, [vp, dp]),
    "lpr_tab_gomory_cut": (C.c_int, [vp, ip, dp, C.c_int]),
    "lpr_tab_cutting_plane": (C.c_int, [vp, C.c_int, ip, ip, ip, C.c_int]),
    "lpr_rev_create": (C.c_int, [C.c_int, C.c_int, C.c_int, dp, dp, dp, C.c_int, C.POINTER(vp)]),
    "lpr_rev_create_dense_lp": (C.c_int, [C.c_int, C.c_uint64, C.c_int, C.c_int, C.POINTER(vp)]),
    "lpr_rev_destroy": (C.c_int, [vp]),
    "lpr_rev_solve": (C.c_int, [vp, C.c_int64, C.c_int, ip, lp, ip, C.c_int64]),
    "lpr_rev_refactor": (C.c_int, [vp]),
    "lpr_rev_refactor_ex": (C.c_int, [vp, C.c_int]),
    "lpr_rev_last_refactor_path": (C.c_int, [vp, ip, dp]),
    "lpr_rev_write_binv": (C.c_int, [vp, dp]),
    "lpr_rev_begin": (C.c_int, [vp]),
    "lpr_rev_step": (C.c_int, [vp, ip, ip, ip, ip]),
    "lpr_rev_format_snapshot": (C.c_int, [vp, C.POINTER(vp), lp]),
    "lpr_fmt_fixed": (C.c_int, [C.c_double, C.c_int, C.c_char_p, C.c_int]),
    "lpr_rev_read_basis": (C.c_int, [vp, ip]),
    "lpr_rev_read_x": (C.c_int, [vp, dp]),
    "lpr_rev_read_z": (C.c_int, [vp, dp]),
    "lpr_rev_read_y": (C.c_int, [vp, dp]),
    "lpr_rev_read_xb": (C.c_int, [vp, dp]),
    "lpr_rev_read_binv": (C.c_int, [vp, dp]),
    "lpr_rev_last_solve_ms": (C.c_int, [vp, C.POINTER(C.c_float)]),
    "lpr_rev_last_refactor_ms": (C.c_int, [vp, C.POINTER(C.c_float)]),
    "lpr_rev_last_refactor_info": (C.c_int, [vp, dp, dp]),
    "lpr_tab_round4": (C.c_int, [vp]),
    "lpr_tab_bb_node_solve": (C.c_int, [vp, C.c_int64, ip, lp, ip, C.c_int64]),
    "lpr_tab_bb_node_solve_ex": (C.c_int, [vp, C.c_int, C.c_int64, ip, lp, ip, C.c_int64]),
    "lpr_tab_create_bb": (C.c_int, [C.c_int, C.c_int, C.c_int, dp, dp, C.c_int, ip, C.c_int, C.c_int, C.POINTER(vp)]),
    "lpr_tab_bb_add_constraint": (C.c_int, [vp, C.c_int, C.c_int, C.c_double, C.c_int, C.POINTER(vp)]),
    "lpr_tab_bb_branch_var": (C.c_int, [vp, C.c_int, ip, dp, dp]),
    "lpr_bb_solve": (C.c_int, [C.c_int, C.c_int, C.c_int, dp, C.c_int, C.c_int, C.c_int64, dp, dp, ip, lp, lp,
                               ip, dp, C.c_int64, ip]),
    "lpr_bb_create": (C.c_int, [C.c_int, C.c_int, C.c_int, dp, C.c_int, C.c_int, C.POINTER(vp)]),
    "lpr_bb_destroy": (C.c_int, [vp]),
    "lpr_bb_open_count": (C.c_int, [vp, lp]),
    "lpr_bb_run": (C.c_int, [vp, C.c_int64, lp, lp]),
    "lpr_bb_run_timed": (C.c_int, [vp, C.c_int64, C.c_double, lp, lp]),
    "lpr_bb_keep_stride": (C.c_int, [vp, C.c_int, C.c_int]),
    "lpr_bb_stats": (C.c_int, [vp, lp, lp, lp, ip]),
    "lpr_bb_get_incumbent": (C.c_int, [vp, ip, dp, dp, ip, ip]),
    "lpr_bb_set_incumbent": (C.c_int, [vp, C.c_double, dp, ip, C.c_int]),
    "lpr_bb_export_nodes": (C.c_int, [vp, C.c_int, vp, C.c_int64, lp, ip]),
    "lpr_bb_import_nodes": (C.c_int, [vp, vp, C.c_int64]),
    "lpr_knap_dp": (C.c_int, [C.c_int, C.c_int, C.c_int, ip, ip, dp, bp]),
    "lpr_knap_create": (C.c_int, [C.c_int, C.c_double, C.c_int, dp, dp, C.POINTER(vp)]),
    "lpr_knap_destroy": (C.c_int, [vp]),
    "lpr_knap_run": (C.c_int, [vp, C.c_int64, lp, ip]),
    "lpr_knap_run_timed": (C.c_int, [vp, C.c_int64, C.c_double, lp, ip]),
    "lpr_bb_solve_mgpu": (C.c_int, [C.c_int, ip, C.c_int, C.c_int, dp, C.c_int, C.c_int, C.c_int64, C.c_int64, C.c_double,
                                    dp, dp, ip, lp, lp, ip, C.POINTER(MgpuStats)]),
    "lpr_knap_solve_mgpu": (C.c_int, [C.c_int, ip, C.c_double, C.c_int, dp, dp, C.c_int64, C.c_int64, C.c_double, dp, bp,
                                      lp, ip, C.POINTER(MgpuStats)]),
    "lpr_nccl_version": (C.c_int, [ip]),
    "lpr_knap_open_count": (C.c_int, [vp, lp]),
    "lpr_knap_keep_stride": (C.c_int, [vp, C.c_int, C.c_int]),
    "lpr_knap_get_incumbent": (C.c_int, [vp, dp, bp, u64p, ip]),
    "lpr_knap_set_incumbent": (C.c_int, [vp, C.c_double, bp, u64p, C.c_int]),
    "lpr_knap_export_nodes": (C.c_int, [vp, C.c_int, vp, C.c_int64, lp, ip]),
    "lpr_knap_import_nodes": (C.c_int, [vp, vp, C.c_int64]),
    "lpr_knap_solve": (C.c_int, [C.c_int, C.c_double, C.c_int, dp, dp, C.c_int64, dp, bp, lp, ip]),
    "lpr_model_parse_file": (C.c_int, [C.c_char_p, C.POINTER(vp)]),
    "lpr_model_parse_text": (C.c_int, [C.c_char_p, C.c_int64, C.POINTER(vp)]),
    "lpr_model_from_dense": (C.c_int, [C.c_int, C.c_int, dp, dp, ip, dp, C.c_int, C.POINTER(vp)]),
    "lpr_model_destroy": (C.c_int, [vp]),
    "lpr_model_info": (C.c_int, [vp, ip, ip, ip, ip]),
    "lpr_model_problem_type": (C.c_int, [vp, C.c_char_p, C.c_int]),
    "lpr_model_message": (C.c_int, [vp, C.c_char_p, C.c_int]),
    "lpr_model_objective": (C.c_int, [vp, dp]),
    "lpr_model_constraint": (C.c_int, [vp, C.c_int, dp, C.c_int, ip, C.c_char_p, C.c_int, dp]),
    "lpr_model_sign": (C.c_int, [vp, C.c_int, C.c_char_p, C.c_int]),
    "lpr_model_add_cli_bound_rows": (C.c_int, [vp]),
    "lpr_model_add_upper_bound_rows": (C.c_int, [vp]),
    "lpr_tab_create_from_model": (C.c_int, [C.c_int, vp, C.c_int, C.POINTER(vp)]),
    "lpr_model_save_binary": (C.c_int, [vp, C.c_char_p]),
    "lpr_model_load_binary": (C.c_int, [C.c_char_p, C.POINTER(vp)]),
    "lpr_fmt_f3": (C.c_int, [C.c_double, C.c_char_p, C.c_int]),
    "lpr_fmt_n3": (C.c_int, [C.c_double, C.c_char_p, C.c_int]),
    "lpr_fmt_table": (C.c_int, [dp, C.c_int, C.c_int, C.c_int64, C.c_int, C.c_char_p, C.POINTER(C.c_char_p),
                                C.c_int, C.POINTER(vp), lp]),
    "lpr_tab_format": (C.c_int, [vp, C.c_int, C.c_char_p, C.POINTER(C.c_char_p), C.c_int, C.POINTER(vp), lp]),
    "lpr_fmt_general": (C.c_int, [C.c_double, C.c_char_p, C.c_int]),
    "lpr_model_canonical_form": (C.c_int, [vp, C.POINTER(vp), lp]),
    "lpr_out_write_full_results": (C.c_int, [C.c_char_p, C.c_char_p, vp, C.POINTER(C.c_char_p), C.c_int, C.c_double, dp,
                                             C.c_int, C.c_int, C.c_char_p]),
    "lpr_out_write_snapshots_only": (C.c_int, [C.c_char_p, C.c_char_p, C.POINTER(C.c_char_p), C.c_int, C.c_double, dp,
                                               C.c_int, C.c_int, C.c_char_p]),
}


def lib():
    """Load liblprb200.so (built in-tree by __graft_entry__.build()).  Fails loudly."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
                "lpr_381_group_v22_b200 has no CPU fallback.")
        l = C.CDLL(LIB_PATH)
        missing = [name for name in SIGNATURES if not hasattr(l, name)]
        if missing:
            raise ImportError(f"{LIB_PATH} does not export {missing}: stale build? run __graft_entry__.build()")
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(rc):
    if rc != OK:
        raise LprError(rc, lib().lpr_last_error().decode("utf-8", "replace"))


def device_count():
    n = C.c_int(0)
    rc = lib().lpr_device_count(C.byref(n))
    return n.value if rc == OK else 0


def launch_count():
    return int(lib().lpr_launch_count())


def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def pd(a):
    return a.ctypes.data_as(dp) if a is not None else None


def pi(a):
    return a.ctypes.data_as(ip) if a is not None else None
