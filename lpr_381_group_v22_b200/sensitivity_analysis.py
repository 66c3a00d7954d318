"""Host-side mirror of the re-optimisation part of SensitivityAnalysis/SensitivityAnalyzer.cs (SURVEY 8(f)
row 1): the final tableau of a solve stays on the device across "change -> resolve".

  SensitivityAnalyzer(finalTableau, solution, zValue, basicVariables)      :22-41
  AddNewConstraintNonInteractive(tech, rhs)                                :609-659
  ResolveAll = RebuildBasicsFromTableau + DualSimplexIfNeeded + ReOptimize :98-209, :706-723
  ShadowPrices :212-222, CurrentTableau / CurrentZ :727-728, GetBasicRow :64-76

The interactive menus (ranging, add activity, duality prompts) are console code and out of scope.
"""
import numpy as np

from . import _native as N
from .simplex import InvalidOperationException
from .tableau import DeviceTableau


class SensitivityAnalyzer:
    def __init__(self, finalTableau, solution, zValue, basicVariables, device=0, headroom=16):
        T = N.f64(finalTableau).copy()            # :24 Clone()
        self.solutionVector = list(solution)      # :25
        self.finalZ = float(zValue)
        T[0, -1] = self.finalZ                    # :33 "Ensure Z on RHS row 0 is set"
        self._tab = DeviceTableau.from_host(T, device=device, row_cap=T.shape[0] + headroom,
                                            col_cap=T.shape[1] + headroom)
        self._tab.basis = np.asarray(list(basicVariables), dtype=np.int32)
        self._tab.sens_rebuild_basis()            # :36
        self.LastPivotLog = []

    # -- reference members ---------------------------------------------------------------------------
    @property
    def CurrentTableau(self):
        return self._tab.read()

    @property
    def CurrentZ(self):
        return self.finalZ

    @property
    def BasicVariables(self):
        return self._tab.basis.tolist()

    def GetBasicRow(self, col):  # :64-76
        T = self._tab.read_col(col)
        eps = 1e-9
        for i in range(1, len(T)):
            if abs(T[i] - 1.0) < eps and all(not (abs(T[k]) > eps) for k in range(1, len(T)) if k != i):
                return i
        return -1

    def ShadowPrices(self):  # :212-222: Z-C on the slack columns
        rows, cols = self._tab.shape
        m = rows - 1
        n = cols - m - 1
        row0 = self._tab.read_row(0)
        return [float(row0[n + i]) for i in range(m)]

    def ResolveAll(self, maxIter=10000):  # :203-209
        self._tab.sens_rebuild_basis()
        r = self._tab.solve(N.RULE_SENS, max_pivots=maxIter)
        self.LastPivotLog = [tuple(p) for p in r["log"].tolist()]
        if r["status"] == N.INFEASIBLE:
            raise InvalidOperationException("Infeasible after RHS change (dual simplex).")          # :194
        if r["status"] == N.UNBOUNDED:
            raise InvalidOperationException("Unbounded during re-optimization.")                    # :151
        if r["status"] == N.ITER_LIMIT:
            raise InvalidOperationException("Re-optimization exceeded iteration limit.")            # :126 / :184
        self.finalZ = self._tab.objective()                                                         # :156
        self.solutionVector = self._tab.sens_solution().tolist()                                    # :158-164
        return r

    def AddNewConstraintNonInteractive(self, tech, rhs):  # :609-659
        rows, cols = self._tab.shape
        tech = [float(t) for t in tech]
        if len(tech) < cols - 1:
            raise IndexError("tech needs one coefficient per tableau column (the reference indexes tech[j], j < numCols-1)")
        aX = 0.0
        for j in range(min(len(tech), len(self.solutionVector))):  # :643-645, sequential like the C# loop
            aX += tech[j] * self.solutionVector[j]
        self._tab.sens_add_constraint(tech[:cols - 1], rhs - aX)
        return self.ResolveAll()

    def close(self):
        self._tab.close()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()
