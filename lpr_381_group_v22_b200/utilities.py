"""Text formatting used by the snapshot path: Utilities/TableIterationFormater.cs:22-48 and NumFormat.N3
(Simplex/RevisedPrimalSimplexSolver.cs:451-465).  The formatting is native (csrc/host_io.cu: lpr_fmt_f3 /
lpr_fmt_n3 / lpr_fmt_table, and lpr_tab_format for a tableau that lives on the device)."""
import ctypes as C

import numpy as np

from . import _native as N


def _fmt(fn, x):
    buf = C.create_string_buffer(400)
    N.check(fn(float(x), buf, 400))
    return buf.value.decode("ascii")


def F3(x):
    """$"{x:F3}" of the .NET Framework"""
    return _fmt(N.lib().lpr_fmt_f3, x)


def _labels(rowLabels):
    if rowLabels is None:
        return None, 0
    arr = (C.c_char_p * max(1, len(rowLabels)))(*[str(s).encode("utf-8") for s in rowLabels])
    return arr, len(rowLabels)


class TableIterationFormater:
    @staticmethod
    def Format(tab, numOriginalVars, title, rowLabels=None):
        T = np.ascontiguousarray(np.asarray(tab, dtype=np.float64))
        if T.ndim != 2:
            raise ValueError("tab must be a 2-D table")
        arr, n = _labels(rowLabels)
        text, ln = N.vp(), C.c_int64()
        N.check(N.lib().lpr_fmt_table(N.pd(T), T.shape[0], T.shape[1], T.shape[1], int(numOriginalVars),
                                      str(title).encode("utf-8"), arr, n, C.byref(text), C.byref(ln)))
        return C.string_at(text.value, ln.value).decode("utf-8")

    @staticmethod
    def FormatDevice(device_tableau, numOriginalVars, title, rowLabels=None):
        """the same text for a tableau resident in HBM (row blocks streamed through pinned buffers)"""
        arr, n = _labels(rowLabels)
        text, ln = N.vp(), C.c_int64()
        N.check(N.lib().lpr_tab_format(device_tableau._h, int(numOriginalVars), str(title).encode("utf-8"), arr, n,
                                       C.byref(text), C.byref(ln)))
        return C.string_at(text.value, ln.value).decode("utf-8")


class NumFormat:
    EPS = 1e-12

    @staticmethod
    def N3(x):
        return _fmt(N.lib().lpr_fmt_n3, x)

    @staticmethod
    def Fixed(x, decimals):
        """$"{x:F<decimals>}" of the .NET Framework (lpr_fmt_fixed)"""
        buf = C.create_string_buffer(400)
        N.check(N.lib().lpr_fmt_fixed(float(x), int(decimals), buf, 400))
        return buf.value.decode("ascii")
