"""Host-side mirrors of the reference's Simplex/ solver classes.

Same member names, argument meaning and error behaviour as the C# classes, with every pivot
executed by liblprb200 on the GPU (no CPU arithmetic on the path):

  PrimalSimplexSolver         Simplex/PrimalSimplexSolver.cs:10-280
  PrimalSimplexSolver2        Simplex/PrimalSimplexSolver2.cs:8-229
  DualSimplexSolver           Simplex/DualSimplex.cs:6-241
  RevisedPrimalSimplexSolver  Simplex/RevisedPrimalSimplexSolver.cs:10-449

The C# P/Invoke shim a maintainer would drop into the reference is in csharp/ (INTEGRATION.md);
this Python layer exists because the image has no .NET toolchain and the parity tests must
drive the same C ABI.
"""
import ctypes as C

import numpy as np

from . import _native as N
from .tableau import DeviceTableau
from .utilities import NumFormat, TableIterationFormater

# IterationSnapshots are text dumps of the whole tableau (3 per pivot in the reference).  They are
# produced only below this many tableau elements unless trace=True is forced (SURVEY 8b "Snapshots").
TRACE_MAX_ELEMENTS = 4096


class InvalidOperationException(RuntimeError):
    pass


class PrimalSimplexSolver:
    """Simplex/PrimalSimplexSolver.cs.  Members: IterationSnapshots, FinalZ, SolutionVector,
    FinalTableau, FinalLabels, FinalTable, Solve(), GetFinalTableau(), BasicVariables."""

    def __init__(self, objective, constraints, isMaximization=True, device=0, trace=None, max_pivots=-1):
        self.numVariables = len(objective)
        self.numConstraints = len(constraints)
        self._tab = DeviceTableau.from_model(objective, constraints, isMaximization, device=device)
        rows, cols = self._tab.shape
        self._trace = (rows * cols <= TRACE_MAX_ELEMENTS) if trace is None else bool(trace)
        self._max_pivots = max_pivots
        self.IterationSnapshots = []
        self.FinalZ = 0.0
        self.SolutionVector = None
        self.FinalTableau = None
        self.Status = N.RUNNING
        self.PivotLog = []
        self.console = []
        if self._trace:
            self._capture("Initial Tableau")  # :86

    # -- helpers ---------------------------------------------------------------------------------
    def _capture(self, title):
        # the tableau stays in HBM: rows stream through pinned buffers into the native formatter (lpr_tab_format)
        self.IterationSnapshots.append(TableIterationFormater.FormatDevice(self._tab, self.numVariables, title))

    def _col_label(self, col):  # :253-254
        return f"x{col + 1}" if col < self.numVariables else f"t{col - self.numVariables + 1}"

    def _solution_summary(self, title="Optimal solution"):  # :256-267
        # {v:F6} with the .NET Framework rules (15 significant digits, half away from zero, no "-0.000000"): native
        lines = [title + ":", "Z = " + NumFormat.Fixed(self.FinalZ, 6)]
        if self.SolutionVector is not None:
            lines += [f"x{i + 1} = " + NumFormat.Fixed(self.SolutionVector[i], 6) for i in range(self.numVariables)]
        return "\r\n".join(lines) + "\r\n"

    # -- Solve :102-150 ---------------------------------------------------------------------------
    def Solve(self):
        if self._trace:
            iteration = 0
            while True:
                if 0 <= self._max_pivots <= iteration:
                    self.Status = N.ITER_LIMIT
                    break
                e, l, st = self._tab.step(N.RULE_PRIMAL)
                if st != N.RUNNING:
                    self.Status = st
                    break
                iteration += 1
                self.PivotLog.append((l, e))
                self.console.append(f"\nIteration {iteration}: pivot @ constraint {l}, column {self._col_label(e)}")
                self._capture(f"Iteration {iteration} - After pivot")  # :148
        else:
            r = self._tab.solve(N.RULE_PRIMAL, max_pivots=self._max_pivots)
            self.Status = r["status"]
            self.PivotLog = [tuple(x) for x in r["log"].tolist()]
        if self.Status == N.OPTIMAL:  # :110-126
            self.FinalZ = self._tab.objective()
            self.SolutionVector = list(self._tab.extract_solution(self.numVariables))
            self.FinalTableau = self._tab.read()
            self.console.append("Optimal Solution Found!")
            if self._trace:
                self.IterationSnapshots.append(
                    TableIterationFormater.Format(self.FinalTableau, self.numVariables, "Final Tableau (Optimal)")
                    + "\r\n" + self._solution_summary() + "\r\n")
        elif self.Status == N.UNBOUNDED:  # :129-135 FinalZ stays 0, SolutionVector stays null
            self.console.append("Unbounded Solution!")
            self.FinalTableau = self._tab.read()
            if self._trace:
                self.IterationSnapshots.append(
                    TableIterationFormater.Format(self.FinalTableau, self.numVariables, "Unbounded Tableau"))

    def GetFinalTableau(self):  # :269-273 returns a clone of the live tableau
        return self._tab.read()

    @property
    def BasicVariables(self):  # :275-278
        return [int(b) for b in self._tab.basis]

    @property
    def FinalLabels(self):  # :23
        return [self._col_label(b) for b in self.BasicVariables]

    @property
    def FinalTable(self):  # :24
        if self.FinalTableau is None:
            return ""
        return TableIterationFormater.Format(self.FinalTableau, self.numVariables, "Final Table", self.FinalLabels)

    @property
    def DeviceTableau(self):
        return self._tab


class PrimalSimplexSolver2:
    """Simplex/PrimalSimplexSolver2.cs: primal simplex from a ready-made tableau (EPS = 1e-10)."""

    def __init__(self, objectiveRow, constraintRows, device=0):
        if objectiveRow is None:
            raise ValueError("objectiveRow")  # ArgumentNullException :26
        if constraintRows is None or len(constraintRows) == 0:
            raise ValueError("No constraint rows.")  # :27
        w = len(objectiveRow)
        if any(len(r) != w for r in constraintRows):
            raise ValueError("All rows (obj & constraints) must have the same length.")  # :29-30
        T = np.vstack([N.f64(objectiveRow)[None, :], N.f64(constraintRows)])
        self._tab = DeviceTableau.from_host(T, device=device)
        self._isOptimal = False
        self.IterationSnapshots = []
        self.FinalZ = 0.0
        self.SolutionVector = []
        self.PivotLog = []

    def Solve(self, maxIters=10_000, printSteps=False):  # :46-97
        r = self._tab.solve(N.RULE_PRIMAL2, max_pivots=maxIters, print_steps=printSteps)
        self.PivotLog += [tuple(x) for x in r["log"].tolist()]
        st = r["status"]
        if st == N.PIVOT_TOO_SMALL:
            raise InvalidOperationException("Pivot too small/zero.")  # :148-149
        if st == N.OPTIMAL:
            self.FinalZ = self._tab.objective()  # :56
            self._isOptimal = True
            return True
        self._isOptimal = False
        return False

    def _ensure_ready(self, solveIfNeeded):  # :220-227
        if not self._isOptimal and solveIfNeeded:
            if not self.Solve():
                raise InvalidOperationException("Could not reach an optimal tableau (unbounded or infeasible).")

    def GetObjectiveRow(self, solveIfNeeded=True):
        self._ensure_ready(solveIfNeeded)
        return self._tab.read_row(0)

    def GetConstraintRows(self, solveIfNeeded=True):
        self._ensure_ready(solveIfNeeded)
        return [row.copy() for row in self._tab.read()[1:]]

    def GetRows(self, solveIfNeeded=True):
        self._ensure_ready(solveIfNeeded)
        return self.GetObjectiveRow(False), self.GetConstraintRows(False)


class DualSimplexSolver:
    """Simplex/DualSimplex.cs: dual simplex that mutates (objectiveRow, constraintRows) in place."""

    EPS = 1e-9

    def __init__(self, device=0):
        self._device = device
        self.PivotLog = []

    def Solve(self, objectiveRow, constraintRows, maxIters=10_000, printSteps=True):  # :14-114
        if objectiveRow is None:
            raise ValueError("objectiveRow")
        if constraintRows is None or len(constraintRows) == 0:
            raise ValueError("No constraint rows.")
        width = len(objectiveRow)
        if any(len(r) != width for r in constraintRows):
            raise ValueError("All rows (obj & constraints) must have the same length.")
        T = np.vstack([N.f64(objectiveRow)[None, :], N.f64(constraintRows)])
        with DeviceTableau.from_host(T, device=self._device) as tab:
            r = tab.solve(N.RULE_DUAL, max_pivots=maxIters, print_steps=printSteps)
            out = tab.read()
        self.PivotLog += [tuple(x) for x in r["log"].tolist()]
        objectiveRow[:] = out[0]  # in-place contract of the reference
        for i, row in enumerate(constraintRows):
            row[:] = out[i + 1]
        if r["status"] == N.PIVOT_TOO_SMALL:
            raise InvalidOperationException("Pivot too small/zero.")  # :155-156
        return r["status"] == N.OPTIMAL

    @staticmethod
    def AnyNegativeRhs(rows):  # :180-185
        if rows is None or len(rows) == 0:
            return False
        return any(r[-1] < -DualSimplexSolver.EPS for r in rows)

    def GetRows(self, objectiveRow, constraintRows, solveIfNeeded=True, maxIters=10_000, printSteps=False):
        if solveIfNeeded and self.AnyNegativeRhs(constraintRows):  # :140-148
            if not self.Solve(objectiveRow, constraintRows, maxIters, printSteps):
                raise InvalidOperationException(
                    "Dual phase could not reach feasibility (infeasible or max iterations).")
        return np.array(objectiveRow, dtype=float), [np.array(r, dtype=float) for r in constraintRows]


class RevisedPrimalSimplexSolver:
    """Simplex/RevisedPrimalSimplexSolver.cs: revised primal simplex with B^-1 resident in HBM.
    Relation is ignored (every row is <=, :55-61).  Infeasible / unbounded / tiny pivot raise the
    reference's Exception messages (:91, :179, :267)."""

    def __init__(self, objective, constraints, isMinimization, device=0, refactor_every=0, max_iter=-1, trace=None):
        if objective is None or len(objective) == 0:
            raise ValueError("Objective cannot be null or empty.")
        if constraints is None or len(constraints) == 0:
            raise ValueError("Constraints cannot be null or empty.")
        self.numVariables = n = len(objective)
        self.numConstraints = m = len(constraints)
        for i, c in enumerate(constraints):
            if len(c.Coefficients) != n:
                raise ValueError(f"Constraint {i + 1} has incorrect number of coefficients.")
        self.isMinimization = bool(isMinimization)
        A = N.f64([c.Coefficients for c in constraints])
        b = N.f64([c.RHS for c in constraints])
        cobj = N.f64(objective)
        h = N.vp()
        N.check(N.lib().lpr_rev_create(device, m, n, N.pd(A), N.pd(b), N.pd(cobj), int(self.isMinimization),
                                       C.byref(h)))
        self._h = h
        self._refactor_every = refactor_every
        self._max_iter = max_iter
        # one CaptureSnapshot text block per iteration like the reference (:226-246), below the size threshold of
        # SURVEY 8b "Snapshots" unless forced; above it Solve() runs entirely on the device and records no text
        self._trace = (m * (n + m + 1) <= TRACE_MAX_ELEMENTS) if trace is None else bool(trace)
        self.IterationSnapshots = []
        self.FinalZ = 0.0
        self.SolutionVector = []
        self.PivotLog = []
        self.Status = N.RUNNING

    def close(self):
        if getattr(self, "_h", None) is not None:
            N.lib().lpr_rev_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _raise_for(self, status):
        if status == N.INFEASIBLE:
            raise Exception("Infeasible basis (negative basic value).")  # :91
        if status == N.UNBOUNDED:
            raise Exception("Unbounded problem (no positive component in direction).")  # :179
        if status == N.PIVOT_TOO_SMALL:
            raise Exception("Pivot too small.")  # :267

    def _snapshot(self):
        text, ln = N.vp(), C.c_int64()
        N.check(N.lib().lpr_rev_format_snapshot(self._h, C.byref(text), C.byref(ln)))
        return C.string_at(text.value, ln.value).decode("utf-8")

    def Solve(self):
        lib = N.lib()
        if self._trace:  # one iteration per call, one CaptureSnapshot text per iteration plus the "Optimal" block
            N.check(lib.lpr_rev_begin(self._h))
            self.PivotLog = []
            st, e, lr, lv = C.c_int(), C.c_int(), C.c_int(), C.c_int()
            while True:
                if 0 <= self._max_iter <= len(self.PivotLog):
                    st.value = N.ITER_LIMIT
                    break
                N.check(lib.lpr_rev_step(self._h, C.byref(st), C.byref(e), C.byref(lr), C.byref(lv)))
                if st.value != N.RUNNING:
                    break
                self.PivotLog.append((lr.value, e.value, lv.value))
                self.IterationSnapshots.append(self._snapshot())
                if self._refactor_every > 0 and len(self.PivotLog) % self._refactor_every == 0:
                    N.check(lib.lpr_rev_refactor(self._h))
            self.Status = st.value
            self.Iterations = len(self.PivotLog)
            self._raise_for(st.value)
            if st.value == N.OPTIMAL:
                self.IterationSnapshots.append(self._snapshot())
        else:
            st = C.c_int()
            nit = C.c_int64()
            cap = 1 << 16
            log = np.zeros((cap, 3), dtype=np.int32)
            N.check(lib.lpr_rev_solve(self._h, self._max_iter, self._refactor_every, C.byref(st), C.byref(nit),
                                      N.pi(log), cap))
            self.Status = st.value
            self.Iterations = nit.value
            self.PivotLog = [tuple(x) for x in log[:min(nit.value, cap)].tolist()]
            self._raise_for(st.value)
        x = np.zeros(self.numVariables)
        z = C.c_double()
        N.check(lib.lpr_rev_read_x(self._h, N.pd(x)))
        N.check(lib.lpr_rev_read_z(self._h, C.byref(z)))
        self.SolutionVector = list(x)
        self.FinalZ = z.value

    @property
    def BasicVariables(self):
        b = np.zeros(self.numConstraints, dtype=np.int32)
        N.check(N.lib().lpr_rev_read_basis(self._h, N.pi(b)))
        return [int(v) for v in b]

    @property
    def DualPrices(self):
        """y = c_B B^-1 (:93): only printed by the reference; exported for the 1e-9 parity check."""
        y = np.zeros(self.numConstraints)
        N.check(N.lib().lpr_rev_read_y(self._h, N.pd(y)))
        return y

    @property
    def BasicValues(self):
        xb = np.zeros(self.numConstraints)
        N.check(N.lib().lpr_rev_read_xb(self._h, N.pd(xb)))
        return xb

    @property
    def BInverse(self):
        binv = np.zeros((self.numConstraints, self.numConstraints))
        N.check(N.lib().lpr_rev_read_binv(self._h, N.pd(binv)))
        return binv
