"""Text formatting used by the snapshot path: Utilities/TableIterationFormater.cs:22-48 and
NumFormat.N3 (Simplex/RevisedPrimalSimplexSolver.cs:451-465).  Host-side only (display)."""
from decimal import ROUND_HALF_UP, Decimal


def _net_fixed(x, digits):
    """.NET Framework double.ToString("F<digits>"): the value is first rendered with 15
    significant digits, then rounded half away from zero; a zero result carries no sign."""
    if x != x:
        return "NaN"
    if x in (float("inf"), float("-inf")):
        return "Infinity" if x > 0 else "-Infinity"
    d = Decimal(f"{x:.15g}").quantize(Decimal(1).scaleb(-digits), rounding=ROUND_HALF_UP)
    if d == 0:
        d = abs(d)
    return f"{d:.{digits}f}"


def F3(x):
    return _net_fixed(x, 3)


class TableIterationFormater:
    @staticmethod
    def Format(tab, numOriginalVars, title, rowLabels=None):
        rows, cols = len(tab), len(tab[0])
        out = [f"\n{title}:", "-" * 80]
        hdr = "Table\t" + "".join(f"x{j + 1}\t" for j in range(numOriginalVars))
        hdr += "".join(f"t{j - numOriginalVars + 1}\t" for j in range(numOriginalVars, cols - 1)) + "RHS"
        out.append(hdr)
        out.append("Z\t" + "".join(F3(tab[0][j]) + "\t" for j in range(cols)))
        for i in range(1, rows):
            label = rowLabels[i - 1] if (rowLabels is not None and len(rowLabels) >= i) else f"{i}"
            out.append(label + "\t" + "".join(F3(tab[i][j]) + "\t" for j in range(cols)))
        return "\r\n".join(out) + "\r\n"


class NumFormat:
    EPS = 1e-12

    @staticmethod
    def N3(x):
        if abs(x) < NumFormat.EPS:
            x = 0.0
        r = float(Decimal(f"{x:.15g}").quantize(Decimal("0.001"), rounding=ROUND_HALF_UP))
        if abs(r - round(r)) < NumFormat.EPS:
            return str(int(round(r)))
        s = _net_fixed(r, 3).rstrip("0").rstrip(".")
        if s.startswith("0."):
            s = s  # "0.###" keeps the leading zero
        return s
