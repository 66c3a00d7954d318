"""Independent pure-Python restatement of the .NET Framework text rules the native formatter implements
(csrc/host_io.cu), used only as a checker by tests/test_host_io.py.

  * double.ToString("F3"): the value is rendered with 15 significant digits first (the Framework's
    DoubleToNumber precision), then rounded half away from zero on that decimal string; a zero result has no sign.
  * NumFormat.N3 (Simplex/RevisedPrimalSimplexSolver.cs:455-465): |x| < 1e-12 -> 0; Math.Round(x, 3,
    MidpointRounding.AwayFromZero) = scale by 1e3, split with modf, bump when |fraction| >= 0.5, divide by 1e3;
    integral results print through double.ToString(), the rest through "0.###".
  * TableIterationFormater.Format (Utilities/TableIterationFormater.cs:22-48).
"""
import math
from decimal import ROUND_HALF_UP, Context, Decimal

_CTX = Context(prec=700)  # 1.8e308 with three decimals needs 312 digits


def net_fixed(x, digits):
    if x != x:
        return "NaN"
    if x in (float("inf"), float("-inf")):
        return "Infinity" if x > 0 else "-Infinity"
    d = Decimal(f"{x:.15g}").quantize(Decimal(1).scaleb(-digits), rounding=ROUND_HALF_UP, context=_CTX)
    if d == 0:
        d = abs(d)
    return f"{d:.{digits}f}"


def F3(x):
    return net_fixed(x, 3)


def round3_away(x):
    if abs(x) < 1e16:
        fr, ip = math.modf(x * 1e3)
        if abs(fr) >= 0.5:
            ip += 1.0 if fr > 0 else -1.0
        x = ip / 1e3
    return x


def N3(x):
    if abs(x) < 1e-12:
        x = 0.0
    r = round3_away(x)
    ri = float(round(r)) if abs(r) < 2 ** 52 else r  # Math.Round(r): half to even
    if abs(r - ri) < 1e-12:
        if ri == 0:
            return "0"
        s = f"{ri:.15g}"
        if "e" in s:  # double.ToString() switches to d.dddE+XX at 1e15
            mant, ex = s.split("e")
            return f"{mant}E{'+' if int(ex) >= 0 else '-'}{abs(int(ex)):02d}"
        return s
    s = net_fixed(r, 3).rstrip("0").rstrip(".")
    return s


def format_table(tab, num_original_vars, title, row_labels=None):
    rows, cols = len(tab), len(tab[0])
    out = [f"\n{title}:", "-" * 80]
    hdr = "Table\t" + "".join(f"x{j + 1}\t" for j in range(num_original_vars))
    hdr += "".join(f"t{j - num_original_vars + 1}\t" for j in range(num_original_vars, cols - 1)) + "RHS"
    out.append(hdr)
    out.append("Z\t" + "".join(F3(tab[0][j]) + "\t" for j in range(cols)))
    for i in range(1, rows):
        label = row_labels[i - 1] if (row_labels is not None and len(row_labels) >= i) else f"{i}"
        out.append(label + "\t" + "".join(F3(tab[i][j]) + "\t" for j in range(cols)))
    return "\r\n".join(out) + "\r\n"


# ---- IO/InputFileParser.cs:19-68, restated with regular expressions (independent of csrc/host_io.cu) ----------------
import re

_NUM = re.compile(r"^[\t\n\v\f\r ]*[+-]?(?:(?:\d[\d,]*)(?:\.\d*)?|\.\d+)(?:[eE][+-]?\d+)?[\t\n\v\f\r ]*$")


def net_parse_double(tok):
    """double.Parse(tok, InvariantCulture): NumberStyles.Float | AllowThousands"""
    t = tok.strip("\t\n\v\f\r ")
    if t in ("NaN", "Infinity", "-Infinity"):
        return float(t.replace("Infinity", "inf").replace("NaN", "nan"))
    if not _NUM.match(tok):
        raise ValueError("FormatException")
    v = float(t.replace(",", ""))
    if v in (float("inf"), float("-inf")):
        raise ValueError("OverflowException")
    return v


def parse_model(text):
    """returns None when ReadInputFile takes its 'not formatted correctly' early return, else
    (problem_type, objective, [(coefficients, relation, rhs)], signs); raises ValueError / IndexError where the
    reference throws FormatException / IndexOutOfRangeException"""
    if text.startswith("﻿"):
        text = text[1:]
    lines = re.split(r"\r\n|\r|\n", text)
    if lines and lines[-1] == "":
        lines.pop()  # File.ReadAllLines: a final line terminator does not start a new line
    if len(lines) < 3:
        return None
    obj = lines[0].strip().split(" ")
    ptype = obj[0].lower()
    c = [net_parse_double(t) for t in obj[1:]]
    rows = []
    for ln in lines[1:-1]:
        parts = [p for p in ln.strip().split(" ") if p != ""]
        if len(parts) < len(c) + 2:
            # coefficients are parsed left to right before the index runs out
            for t in parts[:len(c)]:
                net_parse_double(t)
            raise IndexError("IndexOutOfRangeException")
        co = [net_parse_double(parts[j]) for j in range(len(c))]
        rows.append((co, parts[len(c)], net_parse_double(parts[len(c) + 1])))
    return ptype, c, rows, lines[-1].strip().split(" ")


# ---- double.ToString(), CanonicalFormConverter.CanonicalFormForFile (:57-98), OutputFileWrite (:16-137) -------------
def net_general(x):
    """double.ToString() of the .NET Framework: 15 significant digits, fixed notation when -5 < exponent < 15"""
    if x != x:
        return "NaN"
    if x in (float("inf"), float("-inf")):
        return "Infinity" if x > 0 else "-Infinity"
    if x == 0:
        return "0"
    mant, ex = f"{abs(x):.14e}".split("e")
    digits = mant.replace(".", "").rstrip("0")
    ex = int(ex)
    sign = "-" if x < 0 else ""
    if -5 < ex < 15:
        if ex >= 0:
            ip = digits[:ex + 1].ljust(ex + 1, "0")
            fp = digits[ex + 1:]
            return sign + ip + ("." + fp if fp else "")
        return sign + "0." + "0" * (-ex - 1) + digits
    m = digits[0] + ("." + digits[1:] if len(digits) > 1 else "")
    return f"{sign}{m}E{'+' if ex >= 0 else '-'}{abs(ex):02d}"


def _coeff(c):
    return ("+ " if c >= 0 else "") + net_general(c)


def canonical_form(objective, constraints, signs):
    s = "\n=== Canonical Form ===\r\n" + "Z "
    for i, c in enumerate(objective):
        s += f"{_coeff(c * -1)}x{i + 1} "
    s += "= 0\n"
    for i, (coef, _rel, rhs) in enumerate(constraints):
        for j, a in enumerate(coef):
            s += f"{_coeff(a)}x{j + 1} "
        s += f"+ S{i + 1} " + f"= {net_general(rhs)}\n"
    s += "\nSign Restrictions: " + "".join(f"x{i + 1}: {r} " for i, r in enumerate(signs)) + "\n======================\n\r\n"
    return s


def _final(final_z, x):
    s = "=== Final Results ===\r\n" + f"Z* = {N3(final_z)}\r\n"
    for i, v in enumerate(x or []):
        s += f"x{i + 1} = {N3(v)}\r\n"
    return s


def full_results_text(solver, ptype, objective, constraints, signs, snapshots, final_z, x, ts):
    bar = "=" * 60 + "\r\n"
    s = bar + f"Solver: {solver}\r\nProblem type: {ptype}\r\nTimestamp: {ts}\r\n" + bar + canonical_form(objective, constraints, signs)
    if snapshots:
        s += "=== Iteration Snapshots ===\r\n"
        for i, sn in enumerate(snapshots):
            s += f"--- Iteration {i + 1} ---\r\n{sn}\r\n"
        s += "\r\n"
    return s + _final(final_z, x)


def snapshots_only_text(solver, snapshots, final_z, x, ts):
    bar = "=" * 60 + "\r\n"
    s = bar + f"Solver: {solver}\r\nTimestamp: {ts}\r\n" + bar
    if snapshots:
        s += "=== Solver Log ===\r\n"
        for sn in snapshots:
            s += sn + "\r\n" + ("" if sn.endswith("\n") else "\r\n")
    return s + _final(final_z, x)


# ---- RevisedPrimalSimplexSolver.Solve + CaptureSnapshot (Simplex/RevisedPrimalSimplexSolver.cs:82-387), restated loop
# for loop in plain Python (sequential sums like :398-448) -- the checker of lpr_rev_step / lpr_rev_format_snapshot.
def revised_solve_with_snapshots(objective, A, b, is_min=False, max_iter=10_000):
    """returns (snapshots, pivot_log[(leaveRow, enter, leaveVar)], x, finalZ, basis); raises Exception with the
    reference's messages"""
    EPS = 1e-9
    n, m = len(objective), len(A)
    c_orig = [float(v) for v in objective]
    c = [-v if is_min else v for v in c_orig]                                    # :51
    A = [[float(v) for v in row[:n]] for row in A]
    b = [float(v) for v in b]
    Binv = [[1.0 if i == j else 0.0 for j in range(m)] for i in range(m)]
    basis = [n + i for i in range(m)]
    nonbasic = list(range(n))
    cB = [0.0] * m

    def dot(u, v):
        s = 0.0
        for x, y in zip(u, v):
            s += x * y
        return s

    def matvec(M, v):
        return [dot(row, v) for row in M]

    def vecmat(v, M):
        return [dot(v, [M[i][j] for i in range(len(M))]) for j in range(len(M[0]))]

    def matmul(X, Y):                                                            # :426-441
        R = [[0.0] * len(Y[0]) for _ in X]
        for i in range(len(X)):
            for k in range(len(Y)):
                a = X[i][k]
                if abs(a) < EPS:
                    continue
                for j in range(len(Y[0])):
                    R[i][j] += a * Y[k][j]
        return R

    def label(idx):
        return f"x{idx + 1}" if idx < n else f"S{idx - n + 1}"

    def z_original(xB):
        x = [0.0] * n
        for i in range(m):
            if basis[i] < n:
                x[basis[i]] = max(0.0, xB[i])
        return dot(c_orig, x), x

    def capture(title, xB, y, rcX, rcS, enter, rc_pre, u_pre, ratios, basis_pre, leave_row, leave_var_pre, zw, zo):
        L = [title, "Current Tableau (Revised Simplex)",
             "Problem type: " + ("MIN (solving by MAX of -c)" if is_min else "MAX"), "",
             "Dual prices (y = c_B^T B^{-1}):", "\t".join(N3(v) for v in y), "",
             "Reduced costs:", "  x: " + "\t".join(N3(v) for v in rcX), "  s: " + "\t".join(N3(v) for v in rcS), ""]
        if enter >= 0:
            L += [f"Entering variable (chosen pre-pivot): {label(enter)}  (reduced cost pre = {N3(rc_pre)})",
                  "Direction u = B^{-1} a_enter (pre-pivot):", "\t".join(N3(v) for v in u_pre), "",
                  "Ratio test (xB_i / u_i; ∞ if u_i ≤ 0)  [labels = pre-pivot basis]:"]
            for i in range(m):
                L.append(f"{label(basis_pre[i])}: " + ("∞" if ratios[i] == float("inf") else N3(ratios[i])))
            if leave_row >= 0 and leave_var_pre >= 0:
                L += [f"Pivot (pre→post): {label(leave_var_pre)}  →  {label(enter)}    (pivot = {N3(u_pre[leave_row])})", ""]
        L += [f"Working objective Z_working (maxified): {N3(zw)}",
              f"Original objective Z_original ({'MIN' if is_min else 'MAX'}): {N3(zo)}", ""]
        BA = matmul(Binv, A)
        L.append("Table\t" + "".join(f"x{j + 1}\t" for j in range(n)) + "".join(f"S{j + 1}\t" for j in range(m)) + "RHS")
        L.append("Z~\t" + "".join(N3(v) + "\t" for v in rcX) + "".join(N3(v) + "\t" for v in rcS) + N3(zw))
        for i in range(m):
            L.append(label(basis[i]) + "\t" + "".join(N3(v) + "\t" for v in BA[i]) + "".join(N3(v) + "\t" for v in Binv[i])
                     + N3(xB[i]))
        L.append("Basic Variables: " + ", ".join(label(v) for v in basis))
        return "\r\n".join(L) + "\r\n"

    snaps, log = [], []
    it = 0
    while True:
        xB = matvec(Binv, b)
        if any(v < -EPS for v in xB):
            raise Exception("Infeasible basis (negative basic value).")
        y = vecmat(cB, Binv)
        rcX = [c[j] - dot(y, [A[i][j] for i in range(m)]) for j in range(n)]
        rcS = [-y[k] for k in range(m)]
        enter, best = -1, float("-inf")
        for v in sorted(nonbasic):
            rc = rcX[v] if v < n else rcS[v - n]
            if rc > EPS and (enter == -1 or rc > best + EPS or (abs(rc - best) <= EPS and v < enter)):
                best, enter = rc, v
        if enter == -1:
            zo, x = z_original(xB)
            snaps.append(capture("Optimal", xB, y, rcX, rcS, -1, 0.0, [0.0] * m, [float("inf")] * m, list(basis), -1, -1,
                                 dot(cB, xB), zo))
            return snaps, log, x, zo, list(basis)
        if it >= max_iter:
            raise RuntimeError("iteration cap of the restatement")
        u = matvec(Binv, [A[i][enter] for i in range(m)]) if enter < n else [Binv[i][enter - n] for i in range(m)]
        leave, best_ratio, ratios = -1, 1.7976931348623157e308, [0.0] * m
        for i in range(m):
            if u[i] > EPS:
                ratios[i] = xB[i] / u[i]
                if ratios[i] < best_ratio - EPS or (abs(ratios[i] - best_ratio) <= EPS and
                                                    (leave == -1 or basis[i] < basis[leave])):
                    best_ratio, leave = ratios[i], i
            else:
                ratios[i] = float("inf")
        if leave == -1:
            raise Exception("Unbounded problem (no positive component in direction).")
        leave_var = basis[leave]
        basis_pre = list(basis)
        rc_pre = rcX[enter] if enter < n else rcS[enter - n]
        basis[leave] = enter
        nonbasic.remove(enter)
        if leave_var not in nonbasic:
            nonbasic.append(leave_var)
        cB[leave] = c[enter] if enter < n else 0.0
        pivot = u[leave]
        if abs(pivot) < EPS:
            raise Exception("Pivot too small.")
        E = [[1.0 if i == j else 0.0 for j in range(m)] for i in range(m)]
        for i in range(m):
            E[i][leave] = 1.0 / pivot if i == leave else -u[i] / pivot
        Binv = matmul(E, Binv)
        xB = matvec(Binv, b)
        y2 = vecmat(cB, Binv)
        rcX2 = [c[j] - dot(y2, [A[i][j] for i in range(m)]) for j in range(n)]
        rcS2 = [-v for v in y2]
        log.append((leave, enter, leave_var))
        snaps.append(capture(f"Iteration {it + 1}", xB, y2, rcX2, rcS2, enter, rc_pre, u, ratios, basis_pre, leave,
                             leave_var, dot(cB, xB), z_original(xB)[0]))
        it += 1
