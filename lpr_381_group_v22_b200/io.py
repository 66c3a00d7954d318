"""Model types at the boundary: IO/InputFileParser.cs (parser :19-68, Constraint :70-82) and the CLI's extra
constraint rows (Program.cs:114-124, :511-535).  The parsing itself is native (csrc/host_io.cu, lpr_model_*):
this module only mirrors the C# member names over the C ABI."""
import ctypes as C
import os

from . import _native as N


class Constraint:
    """InputFileParser.Constraint (IO/InputFileParser.cs:70-82)."""

    def __init__(self, coefficients, relation, rhs):
        self.Coefficients = list(coefficients)
        self.Relation = relation
        self.RHS = float(rhs)

    def __repr__(self):
        return f"Constraint({self.Coefficients}, {self.Relation!r}, {self.RHS})"


def _check_parse(rc):
    """what throws in the reference comes back as LPR_E_BADARG with the .NET exception's name in front"""
    if rc == N.OK:
        return
    msg = N.lib().lpr_last_error().decode("utf-8", "replace")
    if msg.startswith("FormatException"):
        raise ValueError(msg)
    if msg.startswith("IndexOutOfRangeException"):
        raise IndexError(msg)
    raise N.LprError(rc, msg)


def _string(fn, *args):
    cap = 256
    while True:
        buf = C.create_string_buffer(cap)
        rc = fn(*args, buf, cap)
        if rc == -5 and cap < (1 << 24):  # LPR_E_CAPACITY
            cap *= 16
            continue
        N.check(rc)
        return buf.value.decode("utf-8", "replace")


class Model:
    """Owner of a native lpr_model handle (parsed or dense) -- what lpr_tab_create_from_model consumes."""

    def __init__(self, handle):
        self._h = handle

    @classmethod
    def parse_file(cls, path):
        h = N.vp()
        _check_parse(N.lib().lpr_model_parse_file(os.fsencode(path), C.byref(h)))
        return cls(h)

    @classmethod
    def parse_text(cls, text):
        data = text.encode("utf-8") if isinstance(text, str) else bytes(text)
        h = N.vp()
        _check_parse(N.lib().lpr_model_parse_text(data, len(data), C.byref(h)))
        return cls(h)

    @classmethod
    def from_dense(cls, objective, coef, rhs, relation=None, is_maximization=True):
        c = N.f64(objective)
        A = N.f64(coef).reshape(len(rhs), len(c)) if len(rhs) else N.f64([])
        b = N.f64(rhs)
        rel = N.i32([N.REL[r] if isinstance(r, str) else int(r) for r in relation]) if relation is not None else None
        h = N.vp()
        N.check(N.lib().lpr_model_from_dense(len(c), len(b), N.pd(c), N.pd(A), N.pi(rel) if rel is not None else None,
                                             N.pd(b), int(bool(is_maximization)), C.byref(h)))
        return cls(h)

    @classmethod
    def load_binary(cls, path):
        h = N.vp()
        N.check(N.lib().lpr_model_load_binary(os.fsencode(path), C.byref(h)))
        return cls(h)

    def save_binary(self, path):
        N.check(N.lib().lpr_model_save_binary(self._h, os.fsencode(path)))

    def close(self):
        if self._h is not None:
            N.lib().lpr_model_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def info(self):
        loaded, n, m, ns = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        N.check(N.lib().lpr_model_info(self._h, C.byref(loaded), C.byref(n), C.byref(m), C.byref(ns)))
        return bool(loaded.value), n.value, m.value, ns.value

    @property
    def message(self):
        return _string(N.lib().lpr_model_message, self._h)

    @property
    def problem_type(self):
        return _string(N.lib().lpr_model_problem_type, self._h)

    def objective(self):
        import numpy as np
        c = np.zeros(self.info()[1])
        N.check(N.lib().lpr_model_objective(self._h, N.pd(c)))
        return c.tolist()

    def constraints(self):
        import numpy as np
        out = []
        for i in range(self.info()[2]):
            cnt, rhs = C.c_int(), C.c_double()
            N.check(N.lib().lpr_model_constraint(self._h, i, None, 0, C.byref(cnt), None, 0, C.byref(rhs)))
            co = np.zeros(max(1, cnt.value))
            N.check(N.lib().lpr_model_constraint(self._h, i, N.pd(co), cnt.value, None, None, 0, None))
            rel = _string(lambda buf, cap: N.lib().lpr_model_constraint(self._h, i, None, 0, None, buf, cap, None))
            out.append(Constraint(co[:cnt.value].tolist(), rel, rhs.value))
        return out

    def signs(self):
        return [_string(N.lib().lpr_model_sign, self._h, j) for j in range(self.info()[3])]

    def canonical_form(self):
        """CanonicalFormConverter.CanonicalFormForFile (Utilities/CanonicalFormConverter.cs:57-93)"""
        text, ln = N.vp(), C.c_int64()
        N.check(N.lib().lpr_model_canonical_form(self._h, C.byref(text), C.byref(ln)))
        return C.string_at(text.value, ln.value).decode("utf-8")

    def add_cli_bound_rows(self):
        N.check(N.lib().lpr_model_add_cli_bound_rows(self._h))
        return self

    def add_upper_bound_rows(self):
        N.check(N.lib().lpr_model_add_upper_bound_rows(self._h))
        return self

    def to_device(self, is_maximization=True, device=0):
        """PrimalSimplexSolver..ctor on this model: the tableau is built in HBM from the native arrays"""
        from .tableau import DeviceTableau
        h = N.vp()
        N.check(N.lib().lpr_tab_create_from_model(device, self._h, int(bool(is_maximization)), C.byref(h)))
        return DeviceTableau(h)


class InputFileParser:
    """IO/InputFileParser.cs:10-68: `max|min c...` / `a... rel rhs` / sign line."""

    def __init__(self):
        self.ProblemType = None
        self.ObjectiveCoefficients = []
        self.Constraints = []
        self.SignRestrictions = []
        self.Model = None  # native handle of the last successful read (Model.to_device builds the tableau from it)

    def ReadInputFile(self, file_path):
        model = Model.parse_file(file_path)
        loaded = model.info()[0]
        print(model.message)
        if not loaded:  # the two early returns of :21-34 leave the parser untouched
            model.close()
            return
        self.ProblemType = model.problem_type
        self.ObjectiveCoefficients.extend(model.objective())
        self.Constraints.extend(model.constraints())
        self.SignRestrictions.extend(model.signs())
        self.Model = model


def _strings(items):
    items = list(items or [])
    arr = (C.c_char_p * max(1, len(items)))(*[str(s).encode("utf-8") for s in items])
    return arr, len(items)


class OutputFileWrite:
    """IO/OutputFileWrite.cs:16-137 over lpr_out_write_* (the text is assembled and written natively).  The first
    method takes the parsed model (InputFileParser.Model / io.Model) in place of the reference's four model lists."""

    @staticmethod
    def WriteFullResults(filePath, solverUsed, model, iterationSnapshots, finalZ, solutionVector, append=False,
                         timestamp=None):
        snaps, ns = _strings(iterationSnapshots)
        x = N.f64(solutionVector) if solutionVector is not None and len(solutionVector) else None
        N.check(N.lib().lpr_out_write_full_results(
            os.fsencode(filePath), str(solverUsed).encode("utf-8"), model._h, snaps, ns, float(finalZ),
            N.pd(x) if x is not None else None, len(x) if x is not None else 0, int(bool(append)),
            timestamp.encode("utf-8") if timestamp else None))

    @staticmethod
    def WriteSnapshotsOnly(filePath, solverUsed, snapshots, finalZ, solutionVector, append=True, timestamp=None):
        snaps, ns = _strings(snapshots)
        x = N.f64(solutionVector) if solutionVector is not None and len(solutionVector) else None
        N.check(N.lib().lpr_out_write_snapshots_only(
            os.fsencode(filePath), str(solverUsed).encode("utf-8"), snaps, ns, float(finalZ),
            N.pd(x) if x is not None else None, len(x) if x is not None else 0, int(bool(append)),
            timestamp.encode("utf-8") if timestamp else None))


def add_cli_bound_rows(n, constraints):
    """Program.cs:114-124 / :372-382: menu options 1 and 3 append `x_i <= 1` for every variable;
    the coefficient list has length n+3 with a stray 1 at index n+1.  List form of
    Model.add_cli_bound_rows (lpr_model_add_cli_bound_rows) for callers that hold a List<Constraint>."""
    for i in range(n):
        co = [0.0] * (n + 3)
        co[i] = 1.0
        co[n + 1] = 1.0
        constraints.append(Constraint(co, "<=", 1.0))
    return constraints


def add_upper_bound_constraints(n, sign_restrictions, constraints):
    """Program.cs:511-535 AddUpperBoundConstraints (menu option 2); list form of Model.add_upper_bound_rows."""
    if not sign_restrictions:
        return constraints
    for j in range(n):
        sr = sign_restrictions[min(j, len(sign_restrictions) - 1)] or ""
        s = sr.replace(" ", "")
        if "bin" in s.lower() or "≤1" in s or "<=1" in s:
            co = [0.0] * n
            co[j] = 1.0
            constraints.append(Constraint(co, "<=", 1.0))
    return constraints
