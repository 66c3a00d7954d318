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
