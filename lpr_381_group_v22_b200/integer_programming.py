"""Host-side mirrors of the reference's IntegerProgramming/ classes; all tableau work runs on the GPU.

  BranchAndBoundAdapter        IntegerProgramming/BranchAndBoundAdapter.cs:7-52
  BranchBoundSimplexSolver     IntegerProgramming/BranchBoundSimplexSolver.cs (BranchAndBound :490-1312,
                               DualSimplexSolverBB :12-469)
  CuttingPlaneSolver           IntegerProgramming/CuttingPlaneSolver.cs:8-231
  KnapsackBranchBoundSimplex / KnapsackBranchBoundSolver   Program.cs:430-471 (bodies missing upstream;
                               specification in DESIGN.md "Knapsack B&B")
"""
import ctypes as C

import numpy as np

from . import _native as N
from .simplex import InvalidOperationException
from .tableau import DeviceTableau


def _solve_bb(tableau, n_vars, enable_pruning, max_nodes, device=0, log_cap=4096):
    T = N.f64(tableau)
    R, Cc = T.shape
    x = np.zeros(n_vars)
    z = C.c_double()
    has = C.c_int()
    nodes = C.c_int64()
    piv = C.c_int64()
    st = C.c_int()
    nlog = np.zeros((max(1, log_cap), 4), dtype=np.int32)
    nz = np.zeros(max(1, log_cap))
    N.check(N.lib().lpr_bb_solve(device, R, Cc, N.pd(T), n_vars, int(bool(enable_pruning)), max_nodes, N.pd(x),
                                 C.byref(z), C.byref(has), C.byref(nodes), C.byref(piv), N.pi(nlog), N.pd(nz), log_cap,
                                 C.byref(st)))
    k = min(nodes.value, log_cap)
    return dict(x=x, z=z.value, has_solution=bool(has.value), nodes=nodes.value, pivots=piv.value, status=st.value,
                node_log=nlog[:k].copy(), node_z=nz[:k].copy())


def solve_bb_mgpu(tableau, n_vars, enable_pruning=True, n_gpus=1, max_nodes=-1, max_rounds=-1, slice_seconds=0.0,
                  devices=None):
    """lpr_bb_solve_mgpu: the open-node pool partitioned over `n_gpus` devices INSIDE the library (host threads +
    NCCL, csrc/multi_gpu.cu); this is what a single-process caller like the reference's Program.cs can reach."""
    T = N.f64(tableau)
    R, Cc = T.shape
    x = np.zeros(n_vars)
    z, has, nodes, piv, st = C.c_double(), C.c_int(), C.c_int64(), C.c_int64(), C.c_int()
    stats = N.MgpuStats()
    dev = N.i32(devices) if devices is not None else None
    N.check(N.lib().lpr_bb_solve_mgpu(n_gpus, N.pi(dev), R, Cc, N.pd(T), n_vars, int(bool(enable_pruning)), max_nodes,
                                      max_rounds, float(slice_seconds), N.pd(x), C.byref(z), C.byref(has),
                                      C.byref(nodes), C.byref(piv), C.byref(st), C.byref(stats)))
    return dict(x=x, z=z.value, has_solution=bool(has.value), nodes=nodes.value, pivots=piv.value, status=st.value,
                stats=stats.as_dict())


class BranchAndBoundAdapter:
    """BranchAndBoundAdapter.SolveFromPrimal (BranchAndBoundAdapter.cs:9-24)."""

    @staticmethod
    def SolveFromPrimal(primal, enablePruning=False, isMin=False, max_nodes=20, device=0, n_gpus=1):
        if primal.FinalTableau is None:
            raise InvalidOperationException("Primal simplex has not been solved yet.")
        final = primal.FinalTableau
        n = len(primal.SolutionVector) if primal.SolutionVector is not None else max(1, final.shape[1] - 1)
        if n_gpus > 1:  # throughput mode: the pool partitioned over the box's GPUs (not the reference's 20-node order)
            r = solve_bb_mgpu(final, n, enablePruning, n_gpus, max_nodes)
        else:
            r = _solve_bb(final, n, enablePruning, max_nodes, device)  # isMin is ignored by the reference (Q9)
        x = list(r["x"]) if r["has_solution"] else []
        z = r["z"] if r["has_solution"] else float("-inf")
        BranchAndBoundAdapter.LastRun = r
        return x, z


class BranchBoundSimplexSolver:
    class DualSimplexSolverBB:
        """DualSimplexSolverBB (BranchBoundSimplexSolver.cs:12-469) on device tableaux."""

        def __init__(self, device=0):
            self._device = device

        @staticmethod
        def _strip_flags(constraints):
            """what FormulateTableau does to the CALLER's rows (:42-56): negate the '>=' rows, drop the flags"""
            for row in constraints:
                if row[-1] == 1:
                    for j in range(len(row)):
                        row[j] = -1 * row[j]
            for row in constraints:
                del row[-1]

        def FormulateTableau(self, objectiveFunction, constraints):  # :28-113
            """Builds the tableau on the device and, like the reference, mutates `constraints` in place."""
            with DeviceTableau.from_bb_model(objectiveFunction, constraints, device=self._device) as t:
                T = t.read()
            self._strip_flags(constraints)
            return T

        def PrepareInput(self, objectiveFunction, constraints, isMinimization):  # :281-287
            surplusCount = sum(1 for c in constraints if c[-1] == 1 or c[-1] == 2)
            slackCount = sum(1 for c in constraints if c[-1] != 1 and c[-1] != 2)
            tableau = self.FormulateTableau(objectiveFunction, constraints)
            return tableau, isMinimization, surplusCount, slackCount, len(objectiveFunction)

        def DoDualSimplex(self, objectiveFunction, constraints, isMinimization, tableauOverride=None):
            """:289-468.  Returns (final_tableau, optimalValue or None, pivot_rows, pivot_cols).  Without
            tableauOverride the tableau comes from FormulateTableau (built on the device, never read back) and
            isMinimization selects the optimality test / entering rule; with it (the form B&B uses, :1107/:1174)
            the given tableau is solved."""
            if tableauOverride is None:
                t = DeviceTableau.from_bb_model(objectiveFunction, constraints, device=self._device)
                self._strip_flags(constraints)
            else:
                t = DeviceTableau.from_host(tableauOverride, device=self._device)
            with t:
                r = t.bb_node_solve(is_min=bool(isMinimization))
                T = t.read()
            if r["status"] != N.OPTIMAL:
                return T, None, None, None
            return T, float(T[0, -1]), [int(p[0]) for p in r["log"]], [int(p[1]) for p in r["log"]]

    class BranchAndBound:
        """BranchAndBound (BranchBoundSimplexSolver.cs:490-1312)."""

        def __init__(self, device=0):
            self._device = device
            self.objectiveCoefficients = [0.0, 0.0]
            self.LastRun = None

        def SetNumVars(self, n):  # :497-500
            self.objectiveCoefficients = [0.0] * n

        def RoundTableau(self, tableau):  # :552-567
            with DeviceTableau.from_host(tableau, device=self._device) as t:
                t.round4()
                return t.read()

        def AddConstraint(self, newConstraints, baseTableau=None):  # :694-803 (one constraint per call)
            if baseTableau is None:
                print("Input tableau required")
                return None, None
            if len(newConstraints) != 1:
                raise NotImplementedError("the reference only ever adds one bound row per call (:1104,:1171)")
            row = list(newConstraints[0])
            coeffs, bound, typ = row[:-2], row[-2], int(row[-1])
            var = coeffs.index(1) if 1 in coeffs else 0
            with DeviceTableau.from_host(baseTableau, device=self._device) as t:
                with t.bb_add_constraint(len(self.objectiveCoefficients), var, bound, typ) as ch:
                    return ch.read(), None

        @staticmethod
        def ConfigureProblem(objective, constraints):  # :1233-1251 (appends to the caller's list)
            n = len(objective)
            for i in range(n):
                row = [0.0] * (n + 3)   # one entry longer than a normal row: the reference's own quirk
                row[i] = 1.0
                row[n + 1] = 1.0
                constraints.append(row)
            return objective, constraints

        def RunBranchAndBound(self, objectivePassed, constraintsPassed, isMin, max_nodes=20):  # :1253-1298
            """LP relaxation from the model rows (x_i <= 1 rows appended), 4-d.p. rounding, then the DFS.
            Returns (bestSolution or None, bestValue) like ExecuteBranchAndBound."""
            self.objectiveCoefficients = list(objectivePassed)
            constraintMatrix = [list(r) for r in constraintsPassed]
            self.objectiveCoefficients, constraintMatrix = self.ConfigureProblem(self.objectiveCoefficients, constraintMatrix)
            solver = BranchBoundSimplexSolver.DualSimplexSolverBB(self._device)
            T, opt, _, _ = solver.DoDualSimplex(list(self.objectiveCoefficients), [list(r) for r in constraintMatrix], isMin)
            if opt is None:
                raise InvalidOperationException("initial LP relaxation failed")  # RoundAllTableaux(null) throws :1273
            self.simplexTableaux = [self.RoundTableau(T)]
            return self.ExecuteBranchAndBound(self.simplexTableaux, False, max_nodes)

        def ExecuteBranchAndBound(self, initialTableaux, enablePruning=False, max_nodes=20):  # :1006-1233
            r = _solve_bb(initialTableaux[-1], len(self.objectiveCoefficients), enablePruning, max_nodes, self._device)
            self.LastRun = r
            return (list(r["x"]) if r["has_solution"] else None), (r["z"] if r["has_solution"] else float("-inf"))


class CuttingPlaneSolver:
    """CuttingPlaneSolver.CuttingPlaneSolution (CuttingPlaneSolver.cs:64-229): mutates objectiveRow and
    constraintRows in place (appending the cut rows), recursion unrolled on the device side."""

    def __init__(self, device=0, max_cuts=-1):
        self._device = device
        self._max_cuts = max_cuts
        self.Status = None
        self.CutLog = None

    def CuttingPlaneSolution(self, objectiveRow, constraintRows):
        if objectiveRow is None:
            raise ValueError("objectiveRow")
        if constraintRows is None or len(constraintRows) == 0:
            raise ValueError("No constraint rows.")
        width = len(objectiveRow)
        if any(len(r) != width for r in constraintRows):
            raise ValueError("All rows (objective & constraints) must have the same length.")
        T = np.vstack([N.f64(objectiveRow)[None, :], N.f64(constraintRows)])
        headroom = 64 if self._max_cuts < 0 else self._max_cuts + 1
        with DeviceTableau.from_host(T, device=self._device, row_cap=T.shape[0] + headroom) as t:
            r = t.cutting_plane(self._max_cuts)
            out = t.read()
        self.Status = r["status"]
        self.CutLog = r["log"]
        objectiveRow[:] = out[0]
        n_old = len(constraintRows)
        for i in range(n_old):
            constraintRows[i][:] = out[i + 1]
        for i in range(n_old, out.shape[0] - 1):
            constraintRows.append(out[i + 1].copy())


class KnapsackItem:
    def __init__(self, Id, Value, Weight):
        self.Id, self.Value, self.Weight = Id, Value, Weight


class KnapsackBranchBoundSimplex:
    """Contract from Program.cs:444-463: ctor(capacity, double[] weights, double[] values), Solve() -> best
    value, PrintIterations(), GetSelectedItemsOriginal() -> items with Id (0-based), Value, Weight."""

    def __init__(self, capacity, weights, values, device=0, max_nodes=-1, n_gpus=1):
        self.capacity = float(capacity)
        self.weights = N.f64(weights)
        self.values = N.f64(values)
        self._device = device
        self._max_nodes = max_nodes
        self._n_gpus = n_gpus
        self.mgpu_stats = None
        self.chosen = None
        self.best = None
        self.nodes = 0
        self.status = None

    def Solve(self):
        n = len(self.weights)
        best = C.c_double()
        chosen = np.zeros(n, dtype=np.uint8)
        nodes = C.c_int64()
        st = C.c_int()
        if self._n_gpus > 1:  # node pool partitioned over the box's GPUs inside the library (csrc/multi_gpu.cu)
            stats = N.MgpuStats()
            N.check(N.lib().lpr_knap_solve_mgpu(self._n_gpus, None, self.capacity, n, N.pd(self.weights),
                                                N.pd(self.values), self._max_nodes, -1, 0.0, C.byref(best),
                                                chosen.ctypes.data_as(N.bp), C.byref(nodes), C.byref(st),
                                                C.byref(stats)))
            self.mgpu_stats = stats.as_dict()
        else:
            N.check(N.lib().lpr_knap_solve(self._device, self.capacity, n, N.pd(self.weights), N.pd(self.values),
                                           self._max_nodes, C.byref(best), chosen.ctypes.data_as(N.bp),
                                           C.byref(nodes), C.byref(st)))
        self.best, self.chosen, self.nodes, self.status = best.value, chosen, nodes.value, st.value
        return self.best

    def PrintIterations(self):
        print(f"Knapsack B&B: {self.nodes} nodes processed, best value {self.best}")

    def GetSelectedItemsOriginal(self):
        return [KnapsackItem(i, float(self.values[i]), float(self.weights[i]))
                for i in range(len(self.weights)) if self.chosen is not None and self.chosen[i]]


class KnapsackBranchBoundSolver:
    """KnapsackBranchBoundSolver.Solve(int, int[], int[]) -- the DP arbiter of Program.cs:467-470."""

    @staticmethod
    def Solve(capacity, weights, values, device=0, return_chosen=False):
        w = N.i32(weights)
        v = N.i32(values)
        best = C.c_double()
        chosen = np.zeros(len(w), dtype=np.uint8)
        N.check(N.lib().lpr_knap_dp(device, int(capacity), len(w), N.pi(w), N.pi(v), C.byref(best),
                                    chosen.ctypes.data_as(N.bp)))
        return (best.value, chosen) if return_chosen else best.value
