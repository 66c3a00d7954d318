"""lpr_381_group_v22_b200 -- B200-native dense simplex pivot path behind the solver classes of
Storm-Tarran/LPR_381_Group_V22.  All arithmetic runs in liblprb200.so (hand-written sm_100a CUDA);
there is no CPU fallback."""
from . import _native
from ._native import (CUT_STEP_DONE, INFEASIBLE, ITER_LIMIT, NO_CUT_NEEDED, NO_PIVOT_COL, NODE_LIMIT, OPTIMAL,
                      PIVOT_TOO_SMALL, RULE_DUAL, RULE_PRIMAL, RULE_PRIMAL2, RULE_SENS, RUNNING, STATUS_NAMES,
                      UNBOUNDED, DEPTH_LIMIT, LprError, device_count, launch_count)
from .io import (Constraint, InputFileParser, Model, OutputFileWrite, add_cli_bound_rows,
                 add_upper_bound_constraints)
from .simplex import (DualSimplexSolver, InvalidOperationException, PrimalSimplexSolver, PrimalSimplexSolver2,
                      RevisedPrimalSimplexSolver)
from .tableau import DeviceTableau
from .sensitivity_analysis import SensitivityAnalyzer
from .integer_programming import (BranchAndBoundAdapter, BranchBoundSimplexSolver, CuttingPlaneSolver,
                                  KnapsackBranchBoundSimplex, KnapsackBranchBoundSolver, solve_bb_mgpu)

__all__ = [
    "Constraint", "InputFileParser", "Model", "OutputFileWrite", "add_cli_bound_rows", "add_upper_bound_constraints", "DeviceTableau",
    "PrimalSimplexSolver", "PrimalSimplexSolver2", "DualSimplexSolver", "RevisedPrimalSimplexSolver",
    "BranchAndBoundAdapter", "BranchBoundSimplexSolver", "CuttingPlaneSolver", "KnapsackBranchBoundSimplex",
    "KnapsackBranchBoundSolver", "solve_bb_mgpu", "SensitivityAnalyzer", "InvalidOperationException", "LprError", "device_count", "launch_count",
]
