"""Tree-walking evaluator for the AST of csparse.py plus the slice of the .NET Base Class Library the reference uses.

TEST INFRASTRUCTURE ONLY (see oracle/csharp/__init__.py).

Value model
    int / long      -> Python int (bool is kept apart)          double / float -> Python float (IEEE binary64)
    string, char    -> str                                       null           -> None
    T[] / T[,]      -> CsArray (flat row-major storage)          List<T>        -> CsList
    Dictionary      -> CsDict      HashSet -> CsSet              Stack / Queue  -> CsStack / CsQueue
    (a, b) / Tuple  -> CsTuple (Item1.. and the declared names)  T? (Nullable)  -> the value or None
    class instance  -> CsObject    delegate / lambda / method group -> Python callable
    exceptions      -> CsException carrying the .NET type name and message (only these are visible to C# catch)

Typing: C# is statically typed and the evaluator is not, so the two places where the static type changes a VALUE are
applied from the declared types the parser keeps: (1) an int stored into a double variable / field / parameter /
return value / array or List<double> element becomes a double (ECMA-334 10.2.3 implicit numeric conversion), so that
a later "/" is a floating-point division; (2) int / int is integer division truncating toward zero (12.10.3).

BCL rules restated here, with the rule each follows (.NET Framework 4.7.2 reference source):
  * Math.Round(double): clr/src/classlibnative/float/floatdouble.cpp COMDouble::Round -- x if integral, else
    floor(x + 0.5), minus 1 when that is a tie at an odd value, sign of x copied.
  * Math.Round(x, digits[, mode]): mscorlib Math.InternalRound -- for |x| < 1e16: x * 10^digits, Round() (ToEven)
    or modf + bump when |fraction| >= 0.5 (AwayFromZero), then / 10^digits.
  * double -> string: Number.FormatDouble -- 15 significant digits first (DoubleToNumber, precision 15), then the
    format's rounding on that DIGIT STRING, half up (RoundNumber), a result that rounds to zero loses its sign;
    "G": scientific when the exponent is >= 15 or < -5; "R": 15 digits if they round-trip, else 17.
  * List<T>.IndexOf / Contains: EqualityComparer<T>.Default (double.Equals: NaN equals NaN, -0.0 equals 0.0).
  * Enumerable.Min / Max (double): the NaN rules of System.Linq.Enumerable; an empty source throws
    InvalidOperationException; Sum adds in sequence order.
  * Enumerable.OrderBy: stable.  List<T>.Sort / Array.Sort: ArraySortHelper introspective sort (insertion sort up
    to 16 elements, median-of-three quicksort, heapsort at the depth limit) -- NOT stable, restated literally.
  * List<T> indexer / RemoveAt / Insert out of range: ArgumentOutOfRangeException; array index: IndexOutOfRange;
    a foreach over a List<T> that is modified meanwhile: InvalidOperationException.
  * (int)double: truncation toward zero; NaN or out of range gives int.MinValue (x64 cvttsd2si).
  * [DllImport] static extern methods call `Interpreter.native` (pinvoke.py); SafeHandle, IDisposable and `using`
    release handles as the CLR does (minus finalizers: there is no garbage collector to run them).

Known simplifications (none is reached by the reference's solver paths; tests/test_csharp_interpreter.py states the
rules that are): int arithmetic does not wrap at 32 bits; the current culture is the invariant one ("." decimal point,
"," group separator -- a machine set to a decimal-comma culture would print the reference's "{x:F3}" differently);
exception text has no stack trace; overloads are resolved by argument count and a type score, not by the full better-
conversion rules; value tuples are immutable; static constructors, operators, events, generics declarations, async,
unsafe code and goto are not implemented (the parser rejects what it does not know); Math.Pow / Exp / Log go to the
C library, whose last bit may differ from the CLR's.
"""
import functools
import math
import os
import re

from .csparse import parse_source

INT_MIN = -2147483648

EXC_BASE = {
    "Exception": None, "SystemException": "Exception", "ApplicationException": "Exception",
    "ArgumentException": "SystemException", "ArgumentNullException": "ArgumentException",
    "ArgumentOutOfRangeException": "ArgumentException", "InvalidOperationException": "SystemException",
    "IndexOutOfRangeException": "SystemException", "NullReferenceException": "SystemException",
    "FormatException": "SystemException", "OverflowException": "ArithmeticException",
    "ArithmeticException": "SystemException", "DivideByZeroException": "ArithmeticException",
    "InvalidCastException": "SystemException", "KeyNotFoundException": "SystemException",
    "NotSupportedException": "SystemException", "NotImplementedException": "SystemException",
    "IOException": "SystemException", "FileNotFoundException": "IOException",
    "RankException": "SystemException", "ArrayTypeMismatchException": "SystemException",
}
EXC_NS = {"KeyNotFoundException": "System.Collections.Generic.", "IOException": "System.IO.",
          "FileNotFoundException": "System.IO."}


class CsException(Exception):
    """a .NET exception in flight"""

    def __init__(self, tname, message=None, inner=None, param=None):
        self.tname = tname
        defaults = {
            "NullReferenceException": "Object reference not set to an instance of an object.",
            "IndexOutOfRangeException": "Index was outside the bounds of the array.",
            "ArgumentOutOfRangeException": "Index was out of range. Must be non-negative and less than the size of "
                                           "the collection.",
            "InvalidOperationException": "Operation is not valid due to the current state of the object.",
            "DivideByZeroException": "Attempted to divide by zero.",
            "FormatException": "Input string was not in a correct format.",
            "ArgumentNullException": "Value cannot be null.",
            "KeyNotFoundException": "The given key was not present in the dictionary.",
        }
        if message is None:
            message = defaults.get(tname, f"Exception of type 'System.{tname}' was thrown.")
        if param is not None:
            message = f"{message}\r\nParameter name: {param}"
        self.message = message
        self.inner = inner
        Exception.__init__(self, f"{tname}: {message}")

    def is_a(self, tname):
        t = self.tname
        while t is not None:
            if t == tname:
                return True
            t = EXC_BASE.get(t, "Exception" if t != "Exception" else None)
        return False

    def cs_tostring(self):
        return f"{EXC_NS.get(self.tname, 'System.')}{self.tname}: {self.message}"


def null_ref():
    return CsException("NullReferenceException")


# ---------------------------------------------------------------------------------------------------------- values
class CsArray:
    __slots__ = ("elem", "dims", "data")

    def __init__(self, elem, dims, data=None, fill=None):
        self.elem = elem
        self.dims = tuple(dims)
        n = 1
        for d in self.dims:
            n *= d
        self.data = data if data is not None else [fill] * n

    def _off(self, idx):
        dims = self.dims
        if len(idx) != len(dims):
            raise CsException("RankException", "wrong number of indices")
        off = 0
        for i, d in zip(idx, dims):
            if i < 0 or i >= d:
                raise CsException("IndexOutOfRangeException")
            off = off * d + i
        return off

    def get(self, idx):
        return self.data[self._off(idx)]

    def set(self, idx, v):
        if self.elem == "double" and type(v) is int:
            v = float(v)
        self.data[self._off(idx)] = v


class CsList:
    __slots__ = ("items", "elem", "ver", "cap")

    def __init__(self, items=None, elem=None, cap=None):
        self.items = items if items is not None else []
        self.elem = elem
        self.ver = 0
        self.cap = len(self.items) if cap is None else cap      # length of the backing array (List.cs)

    def capacity(self):
        """List<T>.EnsureCapacity: 4 when empty, then doubling; what Sort's depth limit is computed from"""
        n = len(self.items)
        while self.cap < n:
            self.cap = 4 if self.cap == 0 else self.cap * 2
        return self.cap

    def co(self, v):
        if self.elem == "double" and type(v) is int:
            return float(v)
        if type(self.elem) is tuple and type(v) is CsTuple:      # List<(T1 a, T2 b)>: the element type names the fields
            return coerce(v, self.elem)
        return v


class CsSeq(list):
    """an evaluated IEnumerable<T> (LINQ results are produced eagerly)"""
    elem = None


class CsTuple:
    __slots__ = ("vals", "names")

    def __init__(self, vals, names=None):
        self.vals = list(vals)
        self.names = names or [None] * len(self.vals)

    def __eq__(self, o):
        return isinstance(o, CsTuple) and len(o.vals) == len(self.vals) and all(cs_equals(a, b) for a, b in zip(self.vals, o.vals))

    def __hash__(self):
        return hash(tuple(self.vals))

    def member(self, name):
        if name in self.names:
            return self.vals[self.names.index(name)]
        m = re.fullmatch(r"Item(\d+)", name)
        if m and 1 <= int(m.group(1)) <= len(self.vals):
            return self.vals[int(m.group(1)) - 1]
        raise AttributeError(f"tuple has no element {name}")


class CsDict:
    __slots__ = ("d",)

    def __init__(self):
        self.d = {}


class CsSet:
    __slots__ = ("s",)

    def __init__(self, items=()):
        self.s = dict.fromkeys(items)


class CsStack:
    __slots__ = ("items", "elem")

    def __init__(self, items=None, elem=None):
        self.items = items or []
        self.elem = elem


class CsQueue:
    __slots__ = ("items", "elem")

    def __init__(self, items=None, elem=None):
        self.items = items or []
        self.elem = elem


class CsAnon:
    def __init__(self, fields):
        self.f = fields


class CsEnum:
    __slots__ = ("tname", "name", "value")

    def __init__(self, tname, name, value):
        self.tname, self.name, self.value = tname, name, value

    def __eq__(self, o):
        return isinstance(o, CsEnum) and o.tname == self.tname and o.value == self.value

    def __hash__(self):
        return hash((self.tname, self.value))


class CsStringBuilder:
    __slots__ = ("parts",)

    def __init__(self, s=""):
        self.parts = [s] if s else []


class CsStringWriter(CsStringBuilder):
    pass


class TypeVal:
    """a type used as a value: the left side of a static member access"""
    __slots__ = ("name", "cls")

    def __init__(self, name, cls=None):
        self.name, self.cls = name, cls


class NamespaceVal:
    __slots__ = ("path",)

    def __init__(self, path):
        self.path = path


class CsClass:
    def __init__(self, decl, outer, interp):
        _, self.name, self.mods, self.bases, members, self.kind, self.ns = decl
        self.outer = outer
        self.base_names = [b[1].split(".")[-1] for b in self.bases if b[0] == "type"]
        self.fields = {}        # name -> (type, init, static)
        self.field_order = []
        self.props = {}         # name -> decl
        self.methods = {}       # name -> [decl]
        self.ctors = []
        self.nested = {}
        self.statics = {}
        self.static_ready = False
        self.is_static = "static" in self.mods
        for m in members:
            k = m[0]
            if k == "field":
                _, mods, ty, decls = m
                st = "static" in mods or "const" in mods
                for nm, init in decls:
                    self.fields[nm] = (ty, init, st)
                    self.field_order.append(nm)
            elif k == "property":
                self.props[m[3]] = m
                if m[4] == "auto" or (m[4] is None and m[5] is None):
                    self.fields[m[3]] = (m[2], m[6], "static" in m[1])
                    self.field_order.append(m[3])
            elif k == "method":
                self.methods.setdefault(m[3], []).append(m)
            elif k == "ctor":
                self.ctors.append(m)
            elif k == "class":
                c = CsClass(m, self, interp)
                self.nested[c.name] = c
                interp.register(c)
            elif k == "enum":
                interp.enums[m[1]] = {nm: i for i, (nm, _) in enumerate(m[2])}

    def chain(self):
        c = self
        while c is not None:
            yield c
            c = c.outer


class CsObject:
    __slots__ = ("cls", "f")

    def __init__(self, cls):
        self.cls = cls
        self.f = {}


class MethodGroup:
    __slots__ = ("interp", "cls", "name", "this")

    def __init__(self, interp, cls, name, this):
        self.interp, self.cls, self.name, self.this = interp, cls, name, this

    def __call__(self, *args):
        return self.interp.invoke(self.cls, self.name, self.this, list(args), {})


class CsLambda:
    __slots__ = ("interp", "params", "body", "env")

    def __init__(self, interp, params, body, env):
        self.interp, self.params, self.body, self.env = interp, params, body, env

    def __call__(self, *args):
        env = Env(self.env, self.env.this, self.env.cls)
        for p, a in zip(self.params, args):
            env.vars[p] = a
        body = self.body
        if body[0] == "block":
            sig = self.interp.exec_block(body[1], env)
            if sig is not None and sig[0] == "ret":
                return sig[1]
            return None
        return self.interp.ev(body, env)


class Ref:
    """an lvalue passed by out / ref"""
    __slots__ = ("get", "set")

    def __init__(self, get, set):
        self.get, self.set = get, set


class Env:
    __slots__ = ("vars", "types", "parent", "this", "cls")

    def __init__(self, parent, this, cls):
        self.vars = {}
        self.types = {}
        self.parent = parent
        self.this = this
        self.cls = cls

    def find(self, name):
        e = self
        while e is not None:
            if name in e.vars:
                return e
            e = e.parent
        return None


BREAK = ("brk",)
CONTINUE = ("cnt",)
_MISSING = object()


# ------------------------------------------------------------------------------------------------- number formatting
def _number(x):
    """DoubleToNumber(x, precision 15): (negative, digits without trailing zeros, scale) with |x| = 0.digits * 10^scale"""
    if x == 0:
        return False, "", 0
    mant, ex = f"{abs(x):.14e}".split("e")
    return x < 0, mant.replace(".", "").rstrip("0"), int(ex) + 1


def _round_number(neg, digits, scale, pos):
    """Number.RoundNumber: keep pos digits, half up on the digit string"""
    i = min(max(pos, 0), len(digits))
    if pos >= 0 and i == pos and i < len(digits) and digits[i] >= "5":
        while i > 0 and digits[i - 1] == "9":
            i -= 1
        if i > 0:
            digits = digits[:i - 1] + chr(ord(digits[i - 1]) + 1)
        else:
            scale += 1
            digits = "1"
            i = 1
    else:
        if pos < 0:
            i = 0
        while i > 0 and digits[i - 1] == "0":
            i -= 1
    digits = digits[:i]
    if i == 0:
        scale = 0
        neg = False
    return neg, digits, scale


def _fixed(neg, digits, scale, decimals, group=False):
    ip = digits[:scale].ljust(scale, "0") if scale > 0 else "0"
    if group:
        out = []
        while len(ip) > 3:
            out.insert(0, ip[-3:]); ip = ip[:-3]
        out.insert(0, ip)
        ip = ",".join(out)
    s = ("-" if neg else "") + ip
    if decimals > 0:
        fp = ("0" * (-scale) if scale < 0 else "") + digits[max(scale, 0):]
        s += "." + fp[:decimals].ljust(decimals, "0")
    return s


def _general(neg, digits, scale, maxdigits=15):
    if not digits:
        return "0"
    sign = "-" if neg else ""
    if scale > maxdigits or scale < -3:
        m = digits[0] + ("." + digits[1:] if len(digits) > 1 else "")
        ex = scale - 1
        return f"{sign}{m}E{'+' if ex >= 0 else '-'}{abs(ex):02d}"
    if scale > 0:
        ip = digits[:scale].ljust(scale, "0")
        fp = digits[scale:]
    else:
        ip, fp = "0", "0" * (-scale) + digits
    return sign + ip + ("." + fp if fp else "")


def _custom(x_number, fmt):
    """the subset of custom numeric format strings made of 0 # . , and literal text (one section)"""
    neg, digits, scale = x_number
    m = re.search(r"[0#][0#,]*(?:\.[0#]*)?|\.[0#]+", fmt)
    if not m:
        return fmt
    pat = m.group(0)
    prefix, suffix = fmt[:m.start()], fmt[m.end():]
    percent = "%" in fmt
    if percent and digits:
        scale += 2
    ipat, _, fpat = pat.partition(".")
    group = "," in ipat
    ipat = ipat.replace(",", "")
    min_int = ipat.count("0") and (len(ipat) - ipat.index("0"))
    max_frac = len(fpat)
    min_frac = fpat.rfind("0") + 1
    if digits:
        neg, digits, scale = _round_number(neg, digits, scale, scale + max_frac)
    ip = digits[:scale].ljust(scale, "0") if scale > 0 else ""
    ip = ip.rjust(min_int, "0")
    if group:
        out = []
        while len(ip) > 3:
            out.insert(0, ip[-3:]); ip = ip[:-3]
        out.insert(0, ip)
        ip = ",".join(out)
    fp = ("0" * (-scale) if scale < 0 else "") + digits[max(scale, 0):]
    fp = fp[:max_frac].rstrip("0")
    fp = fp.ljust(min_frac, "0")
    return ("-" if neg else "") + prefix + ip + ("." + fp if fp else "") + suffix


def format_double(x, fmt=None):
    if x != x:
        return "NaN"
    if x == math.inf:
        return "Infinity"
    if x == -math.inf:
        return "-Infinity"
    if not fmt:
        fmt = "G"
    m = re.fullmatch(r"([A-Za-z])(\d*)", fmt)
    if m:
        c, prec = m.group(1).upper(), m.group(2)
        if c == "R":
            s = _general(*_number(x))
            try:
                if float(s) == x:
                    return s
            except ValueError:
                pass
            mant, ex = f"{abs(x):.16e}".split("e")
            return _general(x < 0, mant.replace(".", "").rstrip("0"), int(ex) + 1, 17)
        neg, digits, scale = _number(x)
        if c == "G":
            p = int(prec) if prec and int(prec) > 0 else 15
            if p < 15:
                neg, digits, scale = _round_number(neg, digits, scale, p)
                return _general(neg, digits, scale, p)
            if p > 15:
                mant, ex = f"{abs(x):.{p - 1}e}".split("e")
                return _general(x < 0, mant.replace(".", "").rstrip("0"), int(ex) + 1, p)
            return _general(neg, digits, scale)
        if c in "FN":
            d = int(prec) if prec else 2
            neg, digits, scale = _round_number(neg, digits, scale, scale + d)
            return _fixed(neg, digits, scale, d, group=(c == "N"))
        if c == "E":
            d = int(prec) if prec else 6
            if digits:
                neg, digits, scale = _round_number(neg, digits, scale, d + 1)
            mant = (digits or "0").ljust(d + 1, "0")
            ex = scale - 1 if digits else 0
            e = m.group(1)  # keeps the case of the format letter
            return f"{'-' if neg else ''}{mant[0]}{'.' + mant[1:] if d else ''}{e}{'+' if ex >= 0 else '-'}{abs(ex):03d}"
        if c == "P":
            d = int(prec) if prec else 2
            if digits:
                scale += 2
            neg, digits, scale = _round_number(neg, digits, scale, scale + d)
            return _fixed(neg, digits, scale, d, group=True) + " %"
        if c == "C":
            d = int(prec) if prec else 2
            neg, digits, scale = _round_number(neg, digits, scale, scale + d)
            body = _fixed(False, digits, scale, d, group=True)
            return f"(¤{body})" if neg else f"¤{body}"
        raise CsException("FormatException", "Format specifier was invalid.")
    return _custom(_number(x), fmt)


def format_int(v, fmt=None):
    if not fmt:
        return str(v)
    m = re.fullmatch(r"([A-Za-z])(\d*)", fmt)
    if m:
        c, prec = m.group(1).upper(), m.group(2)
        if c == "D":
            s = str(abs(v)).rjust(int(prec) if prec else 0, "0")
            return ("-" if v < 0 else "") + s
        if c in "GR":
            return str(v)
        if c == "X":
            s = format(v & 0xFFFFFFFF if v < 0 else v, "X" if m.group(1) == "X" else "x")
            return s.rjust(int(prec) if prec else 0, "0")
        digits = str(abs(v)).rstrip("0") if v else ""
        num = (v < 0, digits, len(str(abs(v))) if v else 0)
        if c in "FN":
            d = int(prec) if prec else 2
            return _fixed(num[0], num[1], num[2], d, group=(c == "N"))
        raise CsException("FormatException", "Format specifier was invalid.")
    digits = str(abs(v)).rstrip("0") if v else ""
    return _custom((v < 0, digits, len(str(abs(v))) if v else 0), fmt)


def cs_tostring(v, fmt=None):
    if v is None:
        return ""
    t = type(v)
    if t is str:
        return v
    if t is bool:
        return "True" if v else "False"
    if t is float:
        return format_double(v, fmt)
    if t is int:
        return format_int(v, fmt)
    if t is CsEnum:
        return v.name
    if t is CsTuple:
        return "(" + ", ".join(cs_tostring(x) for x in v.vals) + ")"
    if t is CsStringBuilder or t is CsStringWriter:
        return "".join(v.parts)
    if isinstance(v, CsException):
        return v.cs_tostring()
    if t is CsList:
        return "System.Collections.Generic.List`1[" + {"double": "System.Double", "int": "System.Int32", "string": "System.String"}.get(v.elem, "T") + "]"
    if t is CsArray:
        return {"double": "System.Double", "int": "System.Int32", "string": "System.String"}.get(v.elem, "System.Object") + "[" + "," * (len(v.dims) - 1) + "]"
    if t is CsAnon:
        return "{ " + ", ".join(f"{k} = {cs_tostring(x)}" for k, x in v.f.items()) + " }"
    if t is CsObject:
        return (v.cls.ns + "." if v.cls.ns else "") + v.cls.name
    if t is CsDateTime:
        return v.text          # the injected clock already has the "yyyy-MM-dd HH:mm:ss" form the reference asks for
    return str(v)


def format_composite(fmt, args):
    """String.Format: {index[,alignment][:format]}"""
    out, i, n = [], 0, len(fmt)
    while i < n:
        ch = fmt[i]
        if ch == "{":
            if fmt[i + 1:i + 2] == "{":
                out.append("{"); i += 2; continue
            j = fmt.index("}", i)
            hole = fmt[i + 1:j]
            spec = None
            if ":" in hole:
                hole, spec = hole.split(":", 1)
            align = None
            if "," in hole:
                hole, a = hole.split(",", 1)
                align = int(a)
            k = int(hole)
            if k >= len(args):
                raise CsException("FormatException", "Index (zero based) must be greater than or equal to zero and less "
                                                     "than the size of the argument list.")
            out.append(pad(cs_tostring(args[k], spec), align))
            i = j + 1
        elif ch == "}":
            if fmt[i + 1:i + 2] == "}":
                out.append("}"); i += 2; continue
            raise CsException("FormatException")
        else:
            out.append(ch); i += 1
    return "".join(out)


def pad(s, align):
    if align is None:
        return s
    return s.rjust(align) if align >= 0 else s.ljust(-align)


_NUMRE = re.compile(r"^[\t\n\v\f\r ]*[+-]?(?:(?:\d[\d,]*)(?:\.\d*)?|\.\d+)(?:[eE][+-]?\d+)?[\t\n\v\f\r ]*$")


def parse_double(s):
    """double.Parse(s, InvariantCulture): NumberStyles.Float | AllowThousands"""
    if s is None:
        raise CsException("ArgumentNullException", param="s")
    t = s.strip("\t\n\v\f\r ")
    if t in ("NaN", "Infinity", "-Infinity"):
        return float(t.replace("Infinity", "inf").replace("NaN", "nan"))
    if not _NUMRE.match(s):
        raise CsException("FormatException")
    v = float(t.replace(",", ""))
    if v in (math.inf, -math.inf):
        raise CsException("OverflowException", "Value was either too large or too small for a Double.")
    return v


def parse_int(s):
    if s is None:
        raise CsException("ArgumentNullException", param="s")
    if not re.fullmatch(r"\s*[+-]?\d+\s*", s):
        raise CsException("FormatException")
    v = int(s)
    if not INT_MIN <= v <= 2147483647:
        raise CsException("OverflowException", "Value was either too large or too small for an Int32.")
    return v


# ------------------------------------------------------------------------------------------------------ BCL helpers
def math_round(x):
    """COMDouble::Round"""
    if x != x or x in (math.inf, -math.inf):
        return x
    if abs(x) < 9.3e18 and x == float(int(x)):
        return x
    t = x + 0.5
    f = float(math.floor(t))
    if f == t and math.fmod(t, 2.0) != 0:
        f -= 1.0
    return math.copysign(f, x)


_POW10 = [1e0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7, 1e8, 1e9, 1e10, 1e11, 1e12, 1e13, 1e14, 1e15]


def math_round_digits(x, digits, away=False):
    """Math.InternalRound"""
    if digits < 0 or digits > 15:
        raise CsException("ArgumentOutOfRangeException", "Rounding digits must be between 0 and 15, inclusive.", param="digits")
    if x != x:
        return x
    if abs(x) < 1e16:
        p = _POW10[digits]
        x = x * p
        if away:
            fr, ip = math.modf(x)
            if abs(fr) >= 0.5:
                ip += 1.0 if fr > 0 else -1.0
            x = ip
        else:
            x = math_round(x)
        x = x / p
    return x


def fdiv(a, b):
    try:
        return a / b
    except ZeroDivisionError:
        if a != a or a == 0:
            return math.nan
        neg = (math.copysign(1.0, a) < 0) != (math.copysign(1.0, b) < 0)
        return -math.inf if neg else math.inf


def to_int32(x):
    if type(x) is int:
        return x
    if x != x or x in (math.inf, -math.inf):
        return INT_MIN
    v = int(x)
    if not INT_MIN <= v <= 2147483647:
        return INT_MIN
    return v


def cs_equals(a, b):
    """EqualityComparer<T>.Default.Equals"""
    ta, tb = type(a), type(b)
    if ta is float or tb is float:
        if (ta is float or ta is int) and (tb is float or tb is int):
            return a == b or (a != a and b != b)
        return False
    if ta in (int, str, bool, CsEnum, CsTuple) or a is None:
        return a == b
    return a is b


def cs_compare(a, b):
    """Comparer<T>.Default.Compare for numbers and strings (double.CompareTo puts NaN first)"""
    if type(a) is float or type(b) is float:
        if a < b:
            return -1
        if a > b:
            return 1
        if a == b:
            return 0
        if a != a:
            return 0 if b != b else -1
        return 1
    if a is None:
        return 0 if b is None else -1
    if b is None:
        return 1
    if isinstance(a, CsTuple):
        for x, y in zip(a.vals, b.vals):
            c = cs_compare(x, y)
            if c:
                return c
        return 0
    return -1 if a < b else (1 if a > b else 0)


def introsort(keys, cmp, capacity=None):
    """ArraySortHelper<T>.IntrospectiveSort (mscorlib, .NET Framework 4.5+); `capacity` = length of the array that is
    sorted in place (a List<T>'s backing array), from which the depth limit is computed"""
    n = len(keys)
    if n < 2:
        return

    def swap_if_greater(a, b):
        if a != b and cmp(keys[a], keys[b]) > 0:
            keys[a], keys[b] = keys[b], keys[a]

    def insertion(lo, hi):
        for i in range(lo, hi):
            j = i
            t = keys[i + 1]
            while j >= lo and cmp(t, keys[j]) < 0:
                keys[j + 1] = keys[j]
                j -= 1
            keys[j + 1] = t

    def down_heap(i, m, lo):
        d = keys[lo + i - 1]
        while i <= m // 2:
            child = 2 * i
            if child < m and cmp(keys[lo + child - 1], keys[lo + child]) < 0:
                child += 1
            if not cmp(d, keys[lo + child - 1]) < 0:
                break
            keys[lo + i - 1] = keys[lo + child - 1]
            i = child
        keys[lo + i - 1] = d

    def heapsort(lo, hi):
        m = hi - lo + 1
        for i in range(m // 2, 0, -1):
            down_heap(i, m, lo)
        for i in range(m, 1, -1):
            keys[lo], keys[lo + i - 1] = keys[lo + i - 1], keys[lo]
            down_heap(1, i - 1, lo)

    def partition(lo, hi):
        mid = lo + (hi - lo) // 2
        swap_if_greater(lo, mid)
        swap_if_greater(lo, hi)
        swap_if_greater(mid, hi)
        pivot = keys[mid]
        keys[mid], keys[hi - 1] = keys[hi - 1], keys[mid]
        left, right = lo, hi - 1
        while left < right:
            left += 1
            while cmp(keys[left], pivot) < 0:
                left += 1
            right -= 1
            while cmp(pivot, keys[right]) < 0:
                right -= 1
            if left >= right:
                break
            keys[left], keys[right] = keys[right], keys[left]
        keys[left], keys[hi - 1] = keys[hi - 1], keys[left]
        return left

    def intro(lo, hi, depth):
        while hi > lo:
            size = hi - lo + 1
            if size <= 16:
                if size == 1:
                    return
                if size == 2:
                    swap_if_greater(lo, hi)
                    return
                if size == 3:
                    swap_if_greater(lo, hi - 1)
                    swap_if_greater(lo, hi)
                    swap_if_greater(hi - 1, hi)
                    return
                insertion(lo, hi)
                return
            if depth == 0:
                heapsort(lo, hi)
                return
            depth -= 1
            p = partition(lo, hi)
            intro(p + 1, hi, depth)
            hi = p - 1

    log2, m = 0, max(n, capacity or 0)
    while m >= 1:
        log2 += 1
        m //= 2
    intro(0, n - 1, 2 * log2)


def iterate(x):
    t = type(x)
    if t is CsList:
        ver, items, i = x.ver, x.items, 0
        while i < len(items):
            yield items[i]
            if x.ver != ver:
                raise CsException("InvalidOperationException", "Collection was modified; enumeration operation may not execute.")
            i += 1
        return
    if t is CsArray:
        yield from list(x.data)
        return
    if t is CsSeq or t is list or t is tuple:
        yield from x
        return
    if t is str:
        yield from x
        return
    if t is CsDict:
        for k, v in list(x.d.items()):
            yield CsTuple([k, v], ["Key", "Value"])
        return
    if t is CsSet:
        yield from list(x.s)
        return
    if t is CsStack:
        yield from reversed(list(x.items))
        return
    if t is CsQueue:
        yield from list(x.items)
        return
    if x is None:
        raise null_ref()
    raise TypeError(f"not enumerable: {t.__name__}")


def seq(items, elem=None):
    s = CsSeq(items)
    s.elem = elem
    return s


def elem_of(x):
    return getattr(x, "elem", None)


def guess_elem(items, elem):
    if elem is not None:
        return elem
    if items and all(type(v) is float for v in items):
        return "double"
    return None


def no_elements():
    return CsException("InvalidOperationException", "Sequence contains no elements")


def linq(name, src, args):
    """System.Linq.Enumerable extension methods; None when `name` is not one"""
    if src is None:
        raise CsException("ArgumentNullException", param="source")
    if name == "ToList":
        items = list(iterate(src))
        return CsList(items, guess_elem(items, elem_of(src)))
    if name == "ToArray":
        items = list(iterate(src))
        return CsArray(guess_elem(items, elem_of(src)), (len(items),), items)
    if name == "Select":
        f = args[0]
        if getattr(f, "params", None) is not None and len(f.params) == 2:
            return seq([f(v, i) for i, v in enumerate(iterate(src))])
        return seq([f(v) for v in iterate(src)])
    if name == "Where":
        f = args[0]
        return seq([v for v in iterate(src) if f(v)], elem_of(src))
    if name == "Any":
        if args:
            return any(args[0](v) for v in iterate(src))
        for _ in iterate(src):
            return True
        return False
    if name == "All":
        return all(args[0](v) for v in iterate(src))
    if name == "Count" or name == "LongCount":
        if args:
            return sum(1 for v in iterate(src) if args[0](v))
        return sum(1 for _ in iterate(src))
    if name in ("Min", "Max"):
        items = list(iterate(src))
        if args:
            items = [args[0](v) for v in items]
        if not items:
            raise no_elements()
        best = items[0]
        if name == "Min":
            for v in items[1:]:
                if type(v) is float or type(best) is float:
                    if v < best or v != v:          # Enumerable.Min(IEnumerable<double>)
                        best = v
                elif cs_compare(v, best) < 0:
                    best = v
        else:
            for v in items[1:]:
                if type(v) is float or type(best) is float:
                    if v > best or best != best:    # Enumerable.Max(IEnumerable<double>)
                        best = v
                elif cs_compare(v, best) > 0:
                    best = v
        return best
    if name == "Sum":
        items = iterate(src)
        if args:
            items = (args[0](v) for v in items)
        items = list(items)
        if elem_of(src) == "double" or any(type(v) is float for v in items):
            s = 0.0
            for v in items:
                s += v
            return s
        return sum(items)
    if name == "Average":
        items = list(iterate(src))
        if args:
            items = [args[0](v) for v in items]
        if not items:
            raise no_elements()
        s = 0.0
        for v in items:
            s += v
        return s / len(items)
    if name in ("First", "FirstOrDefault", "Last", "LastOrDefault", "Single", "SingleOrDefault"):
        items = list(iterate(src))
        if args:
            items = [v for v in items if args[0](v)]
        if not items:
            if name.endswith("OrDefault"):
                return default_for_elem(elem_of(src))
            raise no_elements() if not args else CsException("InvalidOperationException", "Sequence contains no matching element")
        if name.startswith("Single") and len(items) > 1:
            raise CsException("InvalidOperationException", "Sequence contains more than one element")
        return items[0] if name.startswith(("First", "Single")) else items[-1]
    if name == "ElementAt":
        items = list(iterate(src))
        if not 0 <= args[0] < len(items):
            raise CsException("ArgumentOutOfRangeException", param="index")
        return items[args[0]]
    if name == "Take":
        return seq(list(iterate(src))[:max(args[0], 0)], elem_of(src))
    if name == "Skip":
        return seq(list(iterate(src))[max(args[0], 0):], elem_of(src))
    if name == "TakeWhile":
        out = []
        for v in iterate(src):
            if not args[0](v):
                break
            out.append(v)
        return seq(out, elem_of(src))
    if name == "SkipWhile":
        items = list(iterate(src))
        i = 0
        while i < len(items) and args[0](items[i]):
            i += 1
        return seq(items[i:], elem_of(src))
    if name == "Reverse" and type(src) is not CsList:
        return seq(list(iterate(src))[::-1], elem_of(src))
    if name == "Concat":
        return seq(list(iterate(src)) + list(iterate(args[0])), elem_of(src))
    if name == "Distinct":
        out = []
        for v in iterate(src):
            if not any(cs_equals(v, w) for w in out):
                out.append(v)
        return seq(out, elem_of(src))
    if name == "Contains" and type(src) not in (CsList, str, CsSet, CsDict):
        return any(cs_equals(v, args[0]) for v in iterate(src))
    if name == "DefaultIfEmpty":
        items = list(iterate(src))
        if not items:
            items = [args[0] if args else default_for_elem(elem_of(src))]
            if elem_of(src) == "double" and type(items[0]) is int:
                items[0] = float(items[0])
        return seq(items, elem_of(src))
    if name == "Zip":
        f = args[1]
        return seq([f(a, b) for a, b in zip(iterate(src), iterate(args[0]))])
    if name in ("OrderBy", "OrderByDescending", "ThenBy", "ThenByDescending"):
        items = list(iterate(src))
        f = args[0]
        keyed = [(f(v), v) for v in items]
        sign = -1 if name.endswith("Descending") else 1
        prior = getattr(src, "order", None) if name.startswith("Then") else None

        def cmp(a, b):
            if prior is not None:
                c = prior(a[1], b[1])
                if c:
                    return c
            return sign * cs_compare(a[0], b[0])
        keyed.sort(key=functools.cmp_to_key(cmp))     # Python's sort is stable, like EnumerableSorter
        out = seq([v for _, v in keyed], elem_of(src))
        fkey = f

        def order(a, b, prior=prior, sign=sign, fkey=fkey):
            if prior is not None:
                c = prior(a, b)
                if c:
                    return c
            return sign * cs_compare(fkey(a), fkey(b))
        out.order = order
        return out
    if name == "SelectMany":
        out = []
        for v in iterate(src):
            out.extend(iterate(args[0](v)))
        return seq(out)
    if name == "Aggregate":
        items = list(iterate(src))
        if len(args) == 1:
            if not items:
                raise no_elements()
            acc = items[0]
            for v in items[1:]:
                acc = args[0](acc, v)
            return acc
        acc = args[0]
        for v in items:
            acc = args[1](acc, v)
        return args[2](acc) if len(args) > 2 else acc
    if name == "SequenceEqual":
        a, b = list(iterate(src)), list(iterate(args[0]))
        return len(a) == len(b) and all(cs_equals(x, y) for x, y in zip(a, b))
    if name == "ToDictionary":
        d = CsDict()
        for v in iterate(src):
            k = args[0](v)
            if k in d.d:
                raise CsException("ArgumentException", "An item with the same key has already been added.")
            d.d[k] = args[1](v) if len(args) > 1 else v
        return d
    if name == "ToHashSet":
        return CsSet(iterate(src))
    if name == "Cast" or name == "OfType" or name == "AsEnumerable":
        return seq(list(iterate(src)), elem_of(src))
    if name == "GroupBy":
        groups = {}
        for v in iterate(src):
            groups.setdefault(args[0](v), []).append(v)
        out = []
        for k, vs in groups.items():
            g = seq(vs)
            g.Key = k
            out.append(g)
        return seq(out)
    return _MISSING


def default_for_elem(elem):
    if elem == "double":
        return 0.0
    if elem == "int":
        return 0
    if elem == "bool":
        return False
    return None


def type_name(ty):
    """short element-type tag used for coercion: 'double', 'int', ... or None"""
    if ty is None or ty[0] != "type" or ty[3] or ty[4]:
        return None
    return {"float": "double", "Double": "double", "System.Double": "double", "Int32": "int", "long": "int",
            "String": "string", "Byte": "byte"}.get(ty[1], ty[1])


def default_value(ty):
    if ty is None:
        return None
    if ty[0] == "tupletype":
        if ty[2] or ty[3]:
            return None
        return CsTuple([default_value(t) for t, _ in ty[1]], [n for _, n in ty[1]])
    if ty[3] or ty[4]:
        return None
    n = ty[1]
    if n in ("int", "long", "short", "byte", "uint", "ulong", "ushort", "sbyte", "Int32", "Int64", "IntPtr"):
        return 0
    if n in ("double", "float", "decimal", "Double"):
        return 0.0
    if n == "bool":
        return False
    if n == "char":
        return "\0"
    return None


def coerce(v, ty):
    """implicit conversion to the declared type `ty` where it changes the value"""
    if ty is None:
        return v
    if ty[0] == "type":
        if type(v) is int and not ty[3] and ty[1] in ("double", "float", "decimal", "Double"):
            return float(v)
        return v
    if ty[0] == "tupletype" and type(v) is CsTuple and not ty[2]:
        elems = ty[1]
        if len(elems) == len(v.vals):
            return CsTuple([coerce(x, t) for x, (t, _) in zip(v.vals, elems)],
                           [n if n is not None else o for (_, n), o in zip(elems, v.names)])
    return v


# --------------------------------------------------------------------------------------------------- the interpreter
class Interpreter:
    def __init__(self, now="2025-01-01 00:00:00"):
        self.classes = {}      # simple name -> [CsClass]
        self.enums = {}
        self.aliases = {}
        self.console = []      # everything written to Console.Out
        self.stdin = []        # lines Console.ReadLine() returns
        self.files = {}        # virtual file system: path -> text (writes land here)
        self.now = now
        self.steps = 0
        self.max_steps = None
        self.native = None     # P/Invoke target: an object with call(name, params, values, return_type)
        self._dispatch = {name[2:]: getattr(self, name) for name in dir(self) if name.startswith("e_")}
        self.out_writer = None  # Console.SetOut(user TextWriter)
        self.base_dir = "C:\\LPR_381_Group_V22\\bin\\Debug\\"   # AppDomain.CurrentDomain.BaseDirectory
        self.path_sep = "\\"

    # ---- loading
    def register(self, c):
        self.classes.setdefault(c.name, []).append(c)

    def load_source(self, src, name="<cs>"):
        unit = parse_source(src, name)
        for u in unit[1]:
            if u[0] == "using_alias":
                self.aliases[u[1]] = u[2].split(".")[-1]
        for d in unit[2]:
            if d[0] == "class":
                self.register(CsClass(d, None, self))
            elif d[0] == "enum":
                self.enums[d[1]] = {nm: i for i, (nm, _) in enumerate(d[2])}

    def load_file(self, path):
        with open(path, encoding="utf-8-sig") as f:
            self.load_source(f.read(), os.path.basename(path))

    def find_class(self, name, ctx=None):
        name = name.split(".")[-1]
        name = self.aliases.get(name, name)
        if ctx is not None:
            for c in ctx.chain():
                if name in c.nested:
                    return c.nested[name]
                if c.name == name:
                    return c
        cs = self.classes.get(name)
        return cs[0] if cs else None

    # ---- public API for the golden-vector scripts
    def new(self, cls_name, *args, **named):
        cls = self.find_class(cls_name)
        if cls is None:
            raise KeyError(cls_name)
        return self.construct(cls, list(args), named)

    def call(self, obj, method, *args, **named):
        return self.invoke(obj.cls, method, obj, list(args), named)

    def call_static(self, cls_name, method, *args, **named):
        return self.invoke(self.find_class(cls_name), method, None, list(args), named)

    def get(self, obj, name):
        return self.get_member(obj, name, None)

    def console_text(self):
        return "".join(self.console)

    # ---- objects
    def ensure_static(self, cls):
        if cls.static_ready:
            return
        cls.static_ready = True
        env = Env(None, None, cls)
        for nm in cls.field_order:
            ty, init, st = cls.fields[nm]
            if st or cls.is_static:
                cls.statics[nm] = coerce(self.ev(init, env), ty) if init is not None else default_value(ty)

    def construct(self, cls, args, named):
        self.ensure_static(cls)
        obj = CsObject(cls)
        env = Env(None, obj, cls)
        if "SafeHandle" in cls.base_names:       # System.Runtime.InteropServices.SafeHandle: protected IntPtr handle
            obj.f["handle"] = 0
            obj.f["$released"] = False
        for nm in cls.field_order:
            ty, init, st = cls.fields[nm]
            if not st:
                obj.f[nm] = default_value(ty)
        for nm in cls.field_order:
            ty, init, st = cls.fields[nm]
            if not st and init is not None:
                obj.f[nm] = coerce(self.ev(init, env), ty)
        if cls.ctors:
            ctor = self.pick_overload(cls.ctors, 3, args, named, f"{cls.name}..ctor")
            self.run_method(cls, ctor, 3, ctor[4], None, obj, args, named)
        elif args or named:
            raise TypeError(f"{cls.name} has no constructor taking arguments")
        return obj

    def pick_overload(self, decls, pidx, args, named, what):
        n = len(args) + len(named)
        best = None
        for d in decls:
            ps = d[pidx]
            if n > len(ps) and not any(p[3] == "params" for p in ps):
                continue
            required = sum(1 for p in ps if p[2] is None and p[3] != "params")
            if n < required:
                continue
            if any(k not in [p[1] for p in ps] for k in named):
                continue
            score = 0
            for a, p in zip(args, ps):
                tn = type_name(p[0])
                if tn == "double" and type(a) is float:
                    score += 2
                elif tn == "int" and type(a) is int:
                    score += 2
                elif tn == "string" and type(a) is str:
                    score += 2
                elif tn == "double" and type(a) is int:
                    score += 1
                elif tn in ("double", "int", "string", "bool") and a is not None:
                    score -= 4
            score -= abs(len(ps) - n)
            if best is None or score > best[0]:
                best = (score, d)
        if best is None:
            raise TypeError(f"no overload of {what} takes {n} arguments")
        return best[1]

    def invoke(self, cls, name, this, args, named):
        for c in cls.chain():
            if name in c.methods:
                self.ensure_static(c)
                m = self.pick_overload(c.methods[name], 4, args, named, f"{c.name}.{name}")
                if "static" in m[1] or c.is_static:
                    this_arg = None
                else:
                    this_arg = this if (this is not None and this.cls is c) else None
                    if this_arg is None:
                        raise TypeError(f"{c.name}.{name} needs an instance")
                return self.run_method(c, m, 4, m[5], m[2], this_arg, args, named)
        raise AttributeError(f"{cls.name} has no method {name}")

    def run_method(self, cls, decl, pidx, body, rettype, this, args, named):
        env = Env(None, this, cls)
        ps = decl[pidx]
        for i, (ty, nm, default, mod) in enumerate(ps):
            if mod == "params":
                rest = args[i:]
                if len(rest) == 1 and type(rest[0]) is CsArray:
                    v = rest[0]
                else:
                    v = CsArray(type_name(ty[:3] + ([], False)), (len(rest),), list(rest))
            elif i < len(args):
                v = args[i]
            elif nm in named:
                v = named[nm]
            elif default is not None:
                v = self.ev(default, env)
            else:
                raise TypeError(f"missing argument {nm}")
            if mod in ("out", "ref"):
                env.vars[nm] = v      # a Ref: reads and writes go through it
                env.types[nm] = "ref"
                continue
            v = coerce(v, ty)
            env.vars[nm] = v
            if ty[0] == "type" and ty[1] in ("double", "float") and not ty[3]:
                env.types[nm] = ty
        if body is None:
            if "extern" in decl[1]:          # [DllImport] static extern: platform invoke
                if self.native is None:
                    raise RuntimeError(f"no native library bound for P/Invoke {decl[3]}")
                return self.native.call(decl[3], ps, [env.vars[p[1]] for p in ps], rettype)
            return None
        if body[0] == "exprbody":
            return coerce(self.ev(body[1], env), rettype)
        sig = self.exec_block(body[1], env)
        if sig is not None and sig[0] == "ret":
            return coerce(sig[1], rettype)
        return None

    def safehandle_call(self, obj, name, args):
        """the members of SafeHandle a derived handle class inherits"""
        if name == "SetHandle":
            obj.f["handle"] = args[0]
            return None
        if name == "DangerousGetHandle":
            return obj.f["handle"]
        if name in ("Dispose", "Close"):
            if not obj.f["$released"]:
                obj.f["$released"] = True
                if not self.get_member(obj, "IsInvalid", None):
                    self.invoke(obj.cls, "ReleaseHandle", obj, [], {})
            return None
        if name == "SetHandleAsInvalid":
            obj.f["$released"] = True
            return None
        raise AttributeError(f"SafeHandle has no member {name}")

    def dispose(self, v):
        if type(v) is CsObject:
            if "Dispose" in v.cls.methods:
                self.invoke(v.cls, "Dispose", v, [], {})
            elif "SafeHandle" in v.cls.base_names:
                self.safehandle_call(v, "Dispose", [])
        elif v is not None and type(v) in (CsStringWriter, CsStreamWriter):
            bcl_instance_call(self, v, "Dispose", [], {})

    # ---- statements
    def exec_block(self, stmts, env):
        for s in stmts:
            sig = self.exec(s, env)
            if sig is not None:
                return sig
        return None

    def declare(self, env, name, value, ty):
        env.vars[name] = value
        if ty is not None and ty[0] == "type" and ty[1] in ("double", "float") and not ty[3]:
            env.types[name] = ty
        else:
            env.types.pop(name, None)

    def exec(self, s, env):
        k = s[0]
        if k == "expr":
            self.ev(s[1], env)
            return None
        if k == "local":
            ty = s[1]
            for nm, init in s[2]:
                v = coerce(self.ev(init, env), ty) if init is not None else default_value(ty)
                self.declare(env, nm, v, ty)
            return None
        if k == "if":
            if self.ev(s[1], env):
                return self.exec(s[2], env)
            if s[3] is not None:
                return self.exec(s[3], env)
            return None
        if k == "block":
            return self.exec_block(s[1], env)
        if k == "for":
            # scopes as in ECMA-334 7.7: the loop variable lives once per loop, what the body declares once per
            # iteration (so that a lambda created in the body captures that iteration's variables)
            loop = Env(env, env.this, env.cls)
            for init in s[1]:
                self.exec(init, loop)
            cond, iters, body = s[2], s[3], s[4]
            while cond is None or self.ev(cond, loop):
                sig = self.exec(body, Env(loop, env.this, env.cls))
                if sig is not None:
                    if sig is BREAK:
                        break
                    if sig is not CONTINUE:
                        return sig
                for it in iters:
                    self.ev(it, loop)
                self.tick()
            return None
        if k == "while":
            while self.ev(s[1], env):
                sig = self.exec(s[2], Env(env, env.this, env.cls))
                if sig is not None:
                    if sig is BREAK:
                        break
                    if sig is not CONTINUE:
                        return sig
                self.tick()
            return None
        if k == "dowhile":
            while True:
                sig = self.exec(s[1], Env(env, env.this, env.cls))
                if sig is not None:
                    if sig is BREAK:
                        break
                    if sig is not CONTINUE:
                        return sig
                if not self.ev(s[2], env):
                    break
                self.tick()
            return None
        if k == "foreach":
            _, ty, names, e, body = s
            for v in iterate(self.ev(e, env)):
                scope = Env(env, env.this, env.cls)      # a fresh iteration variable every time round (C# 5)
                if isinstance(names, list):
                    self.deconstruct(names, v, scope, declare=True)
                else:
                    self.declare(scope, names, coerce(v, ty) if ty and ty[1] != "var" else v, ty)
                sig = self.exec(body, scope)
                if sig is not None:
                    if sig is BREAK:
                        break
                    if sig is not CONTINUE:
                        return sig
                self.tick()
            return None
        if k == "return":
            return ("ret", self.ev(s[1], env) if s[1] is not None else None)
        if k == "break":
            return BREAK
        if k == "continue":
            return CONTINUE
        if k == "deconstruct_decl":
            self.deconstruct(s[1], self.ev(s[2], env), env, declare=True)
            return None
        if k == "throw":
            if s[1] is None:
                raise env.find("$caught").vars["$caught"]
            v = self.ev(s[1], env)
            if v is None:
                raise null_ref()
            raise v
        if k == "try":
            _, body, catches, fin = s
            try:
                try:
                    return self.exec(body, env)
                except CsException as ex:
                    for ty, nm, when, blk in catches:
                        if ty is None or ex.is_a(ty[1].split(".")[-1]):
                            if nm is not None:
                                env.vars[nm] = ex
                            if when is not None and not self.ev(when, env):
                                continue
                            saved = env.vars.get("$caught", _MISSING)
                            env.vars["$caught"] = ex
                            try:
                                return self.exec(blk, env)
                            finally:
                                if saved is _MISSING:
                                    env.vars.pop("$caught", None)
                                else:
                                    env.vars["$caught"] = saved
                    raise
            finally:
                if fin is not None:
                    sig = self.exec(fin, env)
                    if sig is not None:
                        raise RuntimeError("control flow out of a finally block")
        if k == "switch":
            v = self.ev(s[1], env)
            chosen = None
            for labels, stmts in s[2]:
                for lab in labels:
                    if lab[0] == "case" and cs_equals(self.ev(lab[1], env), v):
                        chosen = stmts
                        break
                if chosen is not None:
                    break
            if chosen is None:
                for labels, stmts in s[2]:
                    if any(lab[0] == "default" for lab in labels):
                        chosen = stmts
            if chosen is not None:
                sig = self.exec_block(chosen, env)
                if sig is BREAK:
                    return None
                return sig
            return None
        if k == "using":
            res = s[1]
            if res[0] == "local":
                self.exec(res, env)
                held = [env.vars[nm] for nm, _ in res[2]]
            else:
                held = [self.ev(res[1], env)]
            try:
                return self.exec(s[2], env)
            finally:
                for v in reversed(held):
                    self.dispose(v)
        if k == "localfunc":
            m = s[1]
            interp = self

            def fn(*args, _m=m, _env=env):
                e2 = Env(_env, _env.this, _env.cls)
                for (ty, nm, default, mod), a in zip(_m[4], args):
                    e2.vars[nm] = coerce(a, ty)
                    if ty[0] == "type" and ty[1] in ("double", "float") and not ty[3]:
                        e2.types[nm] = ty
                for (ty, nm, default, mod) in _m[4][len(args):]:
                    e2.vars[nm] = coerce(interp.ev(default, e2), ty)
                body = _m[5]
                if body[0] == "exprbody":
                    return coerce(interp.ev(body[1], e2), _m[2])
                sig = interp.exec_block(body[1], e2)
                return coerce(sig[1], _m[2]) if sig is not None and sig[0] == "ret" else None
            env.vars[m[3]] = fn
            return None
        if k == "empty":
            return None
        raise NotImplementedError(f"statement {k}")

    def tick(self):
        self.steps += 1
        if self.max_steps is not None and self.steps > self.max_steps:
            raise RuntimeError("step budget exceeded")

    def deconstruct(self, names, value, env, declare):
        if type(value) is not CsTuple:
            if value is None:
                raise null_ref()
            raise TypeError("cannot deconstruct a non-tuple")
        if len(names) != len(value.vals):
            raise TypeError("deconstruction arity mismatch")
        for nm, v in zip(names, value.vals):
            if nm is None:
                continue
            if isinstance(nm, list):
                self.deconstruct(nm, v, env, declare)
            elif declare:
                env.vars[nm] = v
                env.types.pop(nm, None)
            else:
                self.assign(("name", nm), v, env)

    # ---- names and members
    def lookup(self, name, env):
        e = env.find(name)
        if e is not None:
            v = e.vars[name]
            if type(v) is Ref:
                return v.get()
            return v
        this, cls = env.this, env.cls
        if this is not None:
            if name in this.f:
                return this.f[name]
            c = this.cls
            if name in c.props:
                return self.get_property(c, c.props[name], this)
        if cls is not None:
            for c in cls.chain():
                if name in c.fields and (c.fields[name][2] or c.is_static):
                    self.ensure_static(c)
                    return c.statics[name]
                if name in c.props:
                    return self.get_property(c, c.props[name], this if this is not None and this.cls is c else None)
                if name in c.methods:
                    return MethodGroup(self, c, name, this if this is not None and this.cls is c else None)
                if name in c.nested:
                    return TypeVal(name, c.nested[name])
        c = self.find_class(name, cls)
        if c is not None:
            return TypeVal(c.name, c)
        if name in self.enums or name in BCL_TYPES:
            return TypeVal(name)
        if name in ("System", "LPR_381_Group_V22"):
            return NamespaceVal(name)
        raise NameError(f"C# name {name!r} is not defined")

    def get_property(self, cls, decl, this):
        getter = decl[4]
        env = Env(None, this, cls)
        if getter == "auto":
            return this.f[decl[3]] if this is not None else cls.statics[decl[3]]
        if getter[0] == "exprbody":
            return coerce(self.ev(getter[1], env), decl[2])
        sig = self.exec_block(getter[1], env)
        return coerce(sig[1], decl[2]) if sig is not None else None

    def set_property(self, cls, decl, this, value):
        setter = decl[5]
        value = coerce(value, decl[2])
        if setter == "auto" or (setter is None and decl[4] == "auto"):
            if this is not None:
                this.f[decl[3]] = value
            else:
                cls.statics[decl[3]] = value
            return
        if setter is None:
            raise TypeError(f"property {decl[3]} has no setter")
        env = Env(None, this, cls)
        env.vars["value"] = value
        if setter[0] == "exprbody":
            self.ev(setter[1], env)
        else:
            self.exec_block(setter[1], env)

    def get_member(self, obj, name, env):
        t = type(obj)
        if t is CsObject:
            if name in obj.f:
                return obj.f[name]
            c = obj.cls
            if name in c.props:
                return self.get_property(c, c.props[name], obj)
            if name in c.methods:
                return MethodGroup(self, c, name, obj)
            if name in c.fields:
                self.ensure_static(c)
                return c.statics[name]
            raise AttributeError(f"{c.name} has no member {name}")
        if t is TypeVal:
            c = obj.cls
            if c is not None:
                self.ensure_static(c)
                if name in c.statics:
                    return c.statics[name]
                if name in c.props:
                    return self.get_property(c, c.props[name], None)
                if name in c.methods:
                    return MethodGroup(self, c, name, None)
                if name in c.nested:
                    return TypeVal(name, c.nested[name])
                raise AttributeError(f"{c.name} has no static member {name}")
            if obj.name in self.enums:
                return CsEnum(obj.name, name, self.enums[obj.name][name])
            return bcl_static_member(self, obj.name, name)
        if obj is None:
            raise null_ref()
        if t is NamespaceVal:
            c = self.find_class(name)
            if c is not None:
                return TypeVal(c.name, c)
            if name in self.enums or name in BCL_TYPES or name in ENUMS:
                return TypeVal(name)
            return NamespaceVal(obj.path + "." + name)
        return bcl_instance_member(self, obj, name)

    # ---- assignment
    def assign(self, target, value, env):
        k = target[0]
        if k == "paren":
            return self.assign(target[1], value, env)
        if k == "name":
            name = target[1]
            e = env.find(name)
            if e is not None:
                cur = e.vars[name]
                if type(cur) is Ref:
                    cur.set(value)
                    return value
                ty = e.types.get(name)
                if ty is not None:
                    value = coerce(value, ty)
                e.vars[name] = value
                return value
            this, cls = env.this, env.cls
            if this is not None:
                c = this.cls
                if name in c.props and c.props[name][4] != "auto":
                    self.set_property(c, c.props[name], this, value)
                    return value
                if name in this.f:
                    this.f[name] = coerce(value, c.fields[name][0])
                    return value
            if cls is not None:
                for c in cls.chain():
                    if name in c.fields and (c.fields[name][2] or c.is_static):
                        self.ensure_static(c)
                        c.statics[name] = coerce(value, c.fields[name][0])
                        return value
                    if name in c.props:
                        self.set_property(c, c.props[name], None, value)
                        return value
            raise NameError(f"cannot assign to {name!r}")
        if k == "member":
            obj = self.ev(target[1], env)
            name = target[2]
            if type(obj) is CsObject:
                c = obj.cls
                if name in c.props and c.props[name][4] != "auto":
                    self.set_property(c, c.props[name], obj, value)
                elif name in obj.f:
                    obj.f[name] = coerce(value, c.fields[name][0])
                elif name in c.fields:
                    c.statics[name] = coerce(value, c.fields[name][0])
                else:
                    raise AttributeError(f"{c.name} has no member {name}")
                return value
            if type(obj) is TypeVal and obj.cls is not None:
                c = obj.cls
                self.ensure_static(c)
                if name in c.props and c.props[name][4] != "auto":
                    self.set_property(c, c.props[name], None, value)
                else:
                    c.statics[name] = coerce(value, c.fields[name][0])
                return value
            if obj is None:
                raise null_ref()
            if type(obj) is TypeVal and obj.name == "Console" and name in ("OutputEncoding", "InputEncoding", "Title", "ForegroundColor", "BackgroundColor"):
                return value
            raise TypeError(f"cannot assign member {name} of {type(obj).__name__}")
        if k == "index":
            obj = self.ev(target[1], env)
            idx = [self.ev(i, env) for i in target[2]]
            self.set_index(obj, idx, value)
            return value
        if k == "tuple":
            names = []
            for _, e in target[1]:
                names.append(e)
            if type(value) is not CsTuple or len(value.vals) != len(names):
                raise TypeError("tuple assignment arity mismatch")
            for e, v in zip(names, value.vals):
                if e[0] == "name" and e[1] == "_" and env.find("_") is None:
                    continue
                self.assign(e, v, env)
            return value
        raise TypeError(f"cannot assign to {k}")

    def get_index(self, obj, idx):
        t = type(obj)
        if t is CsList:
            i = idx[0]
            if type(i) is not int or i < 0 or i >= len(obj.items):
                raise CsException("ArgumentOutOfRangeException", param="index")
            return obj.items[i]
        if t is CsArray:
            return obj.get(idx)
        if t is CsDict:
            try:
                return obj.d[idx[0]]
            except KeyError:
                raise CsException("KeyNotFoundException") from None
        if t is str:
            i = idx[0]
            if i < 0 or i >= len(obj):
                raise CsException("IndexOutOfRangeException")
            return obj[i]
        if t is CsSeq:      # IReadOnlyList / IList views
            i = idx[0]
            if i < 0 or i >= len(obj):
                raise CsException("ArgumentOutOfRangeException", param="index")
            return obj[i]
        if obj is None:
            raise null_ref()
        raise TypeError(f"cannot index {t.__name__}")

    def set_index(self, obj, idx, value):
        t = type(obj)
        if t is CsList:
            i = idx[0]
            if type(i) is not int or i < 0 or i >= len(obj.items):
                raise CsException("ArgumentOutOfRangeException", param="index")
            obj.items[i] = obj.co(value)
            obj.ver += 1
            return
        if t is CsArray:
            obj.set(idx, value)
            return
        if t is CsDict:
            obj.d[idx[0]] = value
            return
        if obj is None:
            raise null_ref()
        raise TypeError(f"cannot index-assign {t.__name__}")

    # ---- expressions
    def ev(self, e, env):
        return self._dispatch[e[0]](e, env)

    def e_lit(self, e, env):
        return e[1]

    def e_charlit(self, e, env):
        return e[1]

    def e_paren(self, e, env):
        return self.ev(e[1], env)

    def e_name(self, e, env):
        # the hot path: locals first
        name = e[1]
        en = env
        while en is not None:
            vs = en.vars
            if name in vs:
                v = vs[name]
                if type(v) is Ref:
                    return v.get()
                return v
            en = en.parent
        return self.lookup(name, env)

    def e_this(self, e, env):
        return env.this

    def e_predef(self, e, env):
        return TypeVal(e[1])

    def e_istr(self, e, env):
        out = []
        for part in e[1]:
            if type(part) is str:
                out.append(part)
            else:
                ex, align, fmt = part
                s = cs_tostring(self.ev(ex, env), fmt)
                out.append(pad(s, self.ev(align, env) if align is not None else None))
        return "".join(out)

    def e_member(self, e, env):
        obj = self.ev(e[1], env)
        if obj is None:
            if e[3] or chain_has_nullcond(e[1]):
                return None
            return nullable_member(e[2])
        return self.get_member(obj, e[2], env)

    def e_index(self, e, env):
        obj = self.ev(e[1], env)
        if obj is None and (e[3] or chain_has_nullcond(e[1])):
            return None
        idx = [self.ev(i, env) for i in e[2]]
        return self.get_index(obj, idx)

    def eval_args(self, arglist, env):
        args, named = [], {}
        for name, ex, mod in arglist:
            if mod in ("out", "ref"):
                if ex[0] == "outvar":
                    nm = ex[2]
                    if nm != "_":
                        self.declare(env, nm, default_value(ex[1]), ex[1])
                    target = ("name", nm)
                else:
                    target = ex
                if target[0] == "name" and target[1] == "_" and env.find("_") is None:
                    v = Ref(lambda: None, lambda x: None)
                else:
                    v = Ref(lambda t=target: self.ev(t, env), lambda x, t=target: self.assign(t, x, env))
            else:
                v = self.ev(ex, env)
            if name is None:
                args.append(v)
            else:
                named[name] = v
        return args, named

    def e_call(self, e, env):
        f = e[1]
        if f[0] == "member":
            obj = self.ev(f[1], env)
            name = f[2]
            if obj is None:
                if f[3] or chain_has_nullcond(f[1]):
                    return None
                # an extension method called on null reaches the method: Enumerable.* throws ArgumentNullException
                args, named = self.eval_args(e[2], env)
                if name in LINQ_NAMES:
                    raise CsException("ArgumentNullException", param="source")
                raise null_ref()
            args, named = self.eval_args(e[2], env)
            t = type(obj)
            if t is CsObject:
                c = obj.cls
                if name in c.methods:
                    return self.invoke(c, name, obj, args, named)
                if "SafeHandle" in c.base_names and name not in obj.f:
                    return self.safehandle_call(obj, name, args)
                fld = self.get_member(obj, name, env)   # a delegate-typed field
                return fld(*args)
            if t is NamespaceVal:
                raise NameError(f"{obj.path}.{name} is not a type")
            if t is TypeVal:
                if obj.cls is not None:
                    if name in obj.cls.methods:
                        return self.invoke(obj.cls, name, None, args, named)
                    return self.get_member(obj, name, env)(*args)
                return bcl_static_call(self, obj.name, name, args, named)
            return bcl_instance_call(self, obj, name, args, named)
        if f[0] == "name":
            name = f[1]
            en = env.find(name)
            if en is not None:
                args, named = self.eval_args(e[2], env)
                fn = en.vars[name]
                return fn(*args)
            args, named = self.eval_args(e[2], env)
            this, cls = env.this, env.cls
            if cls is not None:
                for c in cls.chain():
                    if name in c.methods:
                        return self.invoke(c, name, this if (this is not None and this.cls is c) else None, args, named)
            if this is not None and "SafeHandle" in this.cls.base_names and name in ("SetHandle", "DangerousGetHandle", "Dispose"):
                return self.safehandle_call(this, name, args)
            v = self.lookup(name, env)
            return v(*args)
        if f[0] == "base":
            return None
        fn = self.ev(f, env)
        args, named = self.eval_args(e[2], env)
        if fn is None:
            raise null_ref()
        return fn(*args)

    def e_new(self, e, env):
        _, ty, arglist, init = e
        args, named = self.eval_args(arglist, env)
        if ty[0] == "tupletype":
            raise NotImplementedError("new of a tuple type")
        name = ty[1].split(".")[-1]
        name = self.aliases.get(name, name)
        cls = self.find_class(name, env.cls)
        if cls is not None:
            obj = self.construct(cls, args, named)
        else:
            obj = bcl_construct(self, name, ty, args, named)
        if init is not None:
            self.apply_init(obj, init, env)
        return obj

    def apply_init(self, obj, init, env):
        if init[0] == "collinit":
            for kind, item in init[1]:
                if kind == "one":
                    v = self.ev(item, env)
                    if type(obj) is CsList:
                        obj.items.append(obj.co(v)); obj.ver += 1
                    elif type(obj) is CsSet:
                        obj.s[v] = None
                    elif type(obj) is CsStack or type(obj) is CsQueue:
                        obj.items.append(v)
                    else:
                        bcl_instance_call(self, obj, "Add", [v], {})
                else:
                    vals = [self.ev(x, env) for x in item]
                    bcl_instance_call(self, obj, "Add", vals, {})
        else:
            for item in init[1]:
                if item[0] == "prop":
                    val = item[2]
                    if val[0] == "nestedinit":
                        self.apply_init(self.get_member(obj, item[1], env), val[1], env)
                    else:
                        v = self.ev(val, env)
                        c = obj.cls
                        if item[1] in c.props and c.props[item[1]][4] != "auto":
                            self.set_property(c, c.props[item[1]], obj, v)
                        else:
                            obj.f[item[1]] = coerce(v, c.fields[item[1]][0])
                else:
                    key = [self.ev(x, env) for x in item[1]]
                    self.set_index(obj, key, self.ev(item[2], env))

    def e_newarr(self, e, env):
        _, elem_ty, dims, init, rank = e
        elem = type_name(elem_ty) if elem_ty is not None else None
        if dims is not None:
            d = [self.ev(x, env) for x in dims]
            for x in d:
                if x < 0:
                    raise CsException("OverflowException", "Arithmetic operation resulted in an overflow.")
            arr = CsArray(elem, d, fill=default_value(elem_ty))
            if init is not None:
                self.fill_array(arr, init, env)
            return arr
        # dimensions from the initializer
        shape = []
        node = init
        r = rank or (len(elem_ty[3]) and elem_ty[3][0]) or 1
        if elem_ty is not None and rank is None:
            # "T[] x = { ... }" / "T[,] x = { {..}, {..} }": the declared type carries the rank
            r = elem_ty[3][0] if elem_ty[3] else 1
            elem_ty = elem_ty[:3] + (list(elem_ty[3][1:]), elem_ty[4])
            elem = type_name(elem_ty)
        for _ in range(r):
            shape.append(len(node))
            node = node[0] if node and isinstance(node[0], list) else None
        arr = CsArray(elem, shape, fill=default_value(elem_ty) if elem_ty is not None else None)
        self.fill_array(arr, init, env)
        if elem is None:
            arr.elem = guess_elem(arr.data, None)
        return arr

    def fill_array(self, arr, init, env):
        flat = []

        def walk(node, depth):
            if depth == len(arr.dims):
                flat.append(self.ev(node, env))
                return
            if not isinstance(node, list) or len(node) != arr.dims[depth]:
                raise TypeError("array initializer shape mismatch")
            for x in node:
                walk(x, depth + 1)
        walk(init, 0)
        if arr.elem == "double":
            flat = [float(v) if type(v) is int else v for v in flat]
        arr.data = flat

    def e_anon(self, e, env):
        return CsAnon({nm: self.ev(x, env) for nm, x in e[1]})

    def e_tuple(self, e, env):
        return CsTuple([self.ev(x, env) for _, x in e[1]], [n for n, _ in e[1]])

    def e_lambda(self, e, env):
        return CsLambda(self, e[1], e[2], env)

    def e_cond(self, e, env):
        a, b = e[2], e[3]
        v = self.ev(a if self.ev(e[1], env) else b, env)
        if type(v) is int and (is_double_expr(a) or is_double_expr(b)):
            return float(v)
        return v

    def e_coalesce(self, e, env):
        v = self.ev(e[1], env)
        return v if v is not None else self.ev(e[2], env)

    def e_throwexpr(self, e, env):
        raise self.ev(e[1], env)

    def e_assign(self, e, env):
        _, op, target, rhs = e
        if op == "=":
            return self.assign(target, self.ev(rhs, env), env)
        return self.compound(target, op[:-1], lambda: self.ev(rhs, env), env, post=False)

    def compound(self, target, op, rhs, env, post):
        """read-modify-write with the target's sub-expressions evaluated once"""
        k = target[0]
        if k == "paren":
            return self.compound(target[1], op, rhs, env, post)
        if k == "index":
            obj = self.ev(target[1], env)
            idx = [self.ev(i, env) for i in target[2]]
            old = self.get_index(obj, idx)
            new = binop(op, old, rhs())
            if type(old) is int and type(new) is float and op != "??":
                pass
            self.set_index(obj, idx, new)
            return old if post else new
        if k == "member":
            obj = self.ev(target[1], env)
            if obj is None:
                raise null_ref()
            old = self.get_member(obj, target[2], env)
            new = binop(op, old, rhs())
            tmp = Env(env, env.this, env.cls)
            tmp.vars["$obj"] = obj
            self.assign(("member", ("name", "$obj"), target[2], False), new, tmp)
            return old if post else new
        old = self.ev(target, env)
        new = binop(op, old, rhs())
        self.assign(target, new, env)
        return old if post else new

    def e_postfix(self, e, env):
        return self.compound(e[2], "+" if e[1] == "++" else "-", lambda: 1, env, post=True)

    def e_prefix(self, e, env):
        return self.compound(e[2], "+" if e[1] == "++" else "-", lambda: 1, env, post=False)

    def e_unary(self, e, env):
        v = self.ev(e[2], env)
        op = e[1]
        if op == "-":
            return None if v is None else -v
        if op == "!":
            return not v
        if op == "+":
            return v
        return ~v

    def e_binary(self, e, env):
        op = e[1]
        if op == "&&":
            return bool(self.ev(e[2], env)) and bool(self.ev(e[3], env))
        if op == "||":
            return bool(self.ev(e[2], env)) or bool(self.ev(e[3], env))
        return binop(op, self.ev(e[2], env), self.ev(e[3], env))

    def e_cast(self, e, env):
        v = self.ev(e[2], env)
        ty = e[1]
        if ty[0] == "type" and not ty[3]:
            n = ty[1]
            if n in ("int", "long", "short", "byte", "uint", "ulong"):
                if v is None:
                    if ty[4]:
                        return None
                    raise CsException("InvalidOperationException", "Nullable object must have a value.")
                if type(v) is CsEnum:
                    return v.value
                if type(v) is str:
                    return ord(v)
                if n == "long" and type(v) is float:
                    if v != v or abs(v) >= 9.3e18:
                        return -9223372036854775808
                    return int(v)
                return to_int32(v)
            if n in ("double", "float", "decimal"):
                if v is None:
                    if ty[4]:
                        return None
                    raise CsException("InvalidOperationException", "Nullable object must have a value.")
                return float(v)
            if n == "char" and type(v) is int:
                return chr(v)
        return v

    def e_is(self, e, env):
        v = self.ev(e[1], env)
        ok = v is not None and type_matches(v, e[2])
        if ok and e[3] is not None:
            env.vars[e[3]] = v
        return ok

    def e_as(self, e, env):
        v = self.ev(e[1], env)
        return v if v is not None and type_matches(v, e[2]) else None

    def e_typeof(self, e, env):
        return TypeVal(e[1][1])

    def e_default(self, e, env):
        return default_value(e[1]) if e[1] is not None else None

    def e_base(self, e, env):
        return env.this

    def e_outvar(self, e, env):
        raise TypeError("out variable outside an argument list")


def chain_has_nullcond(e):
    """`a?.B().C`: when a is null the WHOLE chain is null (ECMA-334 12.8.8), not only `a?.B()`"""
    while True:
        k = e[0]
        if k == "member" or k == "index":
            if e[3]:
                return True
            e = e[1]
        elif k == "call":
            e = e[1]
        elif k == "paren":
            return False
        else:
            return False


def is_double_expr(e):
    k = e[0]
    if k == "lit":
        return type(e[1]) is float
    if k == "unary" or k == "paren":
        return is_double_expr(e[-1])
    if k == "member":
        return e[1][0] == "predef" and e[1][1] in ("double", "float")
    return False


def type_matches(v, ty):
    n = ty[1].split(".")[-1] if ty[0] == "type" else None
    t = type(v)
    if ty[0] == "type" and ty[3]:
        return t is CsArray
    if n in ("double", "float"):
        return t is float
    if n in ("int", "long"):
        return t is int
    if n == "string":
        return t is str
    if n == "bool":
        return t is bool
    if n == "object":
        return True
    if t is CsObject:
        return v.cls.name == n
    if isinstance(v, CsException):
        return v.is_a(n)
    if n in ("List", "IList", "IEnumerable", "IReadOnlyList", "ICollection"):
        return t in (CsList, CsSeq, CsArray)
    return False


def binop(op, a, b):
    ta, tb = type(a), type(b)
    if op == "+":
        if ta is str or tb is str:
            return cs_tostring(a) + cs_tostring(b)
        if a is None or b is None:
            return None
        return a + b
    if op == "-":
        if a is None or b is None:
            return None
        return a - b
    if op == "*":
        if a is None or b is None:
            return None
        return a * b
    if op == "/":
        if a is None or b is None:
            return None
        if ta is int and tb is int:
            if b == 0:
                raise CsException("DivideByZeroException")
            q = abs(a) // abs(b)
            return q if (a < 0) == (b < 0) else -q
        return fdiv(a, b)
    if op == "%":
        if a is None or b is None:
            return None
        if ta is int and tb is int:
            if b == 0:
                raise CsException("DivideByZeroException")
            r = abs(a) % abs(b)
            return r if a >= 0 else -r
        try:
            return math.fmod(a, b)
        except ValueError:
            return math.nan
    if op == "==":
        return eq_op(a, b)
    if op == "!=":
        return not eq_op(a, b)
    if op in ("<", ">", "<=", ">="):
        if a is None or b is None:
            return False
        if ta is CsEnum:
            a, b = a.value, b.value
        if op == "<":
            return a < b
        if op == ">":
            return a > b
        if op == "<=":
            return a <= b
        return a >= b
    if op == "&":
        return (a and b) if ta is bool else a & b
    if op == "|":
        return (a or b) if ta is bool else a | b
    if op == "^":
        return (a != b) if ta is bool else a ^ b
    if op == "<<":
        return a << (b & 31)
    if op == ">>":
        return a >> (b & 31)
    if op == "??":
        return a if a is not None else b
    raise NotImplementedError(op)


def eq_op(a, b):
    ta, tb = type(a), type(b)
    if a is None or b is None:
        return a is b
    if ta in (int, float, str, bool) and tb in (int, float, str, bool):
        return a == b
    if ta is CsEnum or ta is CsTuple:
        return a == b
    return a is b


# ----------------------------------------------------------------------------------------------------- BCL surface
BCL_TYPES = {"AppDomain", "IntPtr", "Marshal", "Math", "Console", "Enumerable", "String", "Array", "Tuple", "Convert", "CultureInfo", "File", "Path",
             "Environment", "DateTime", "MidpointRounding", "StringSplitOptions", "NumberStyles", "Double", "Int32",
             "Encoding", "StringComparison", "Directory", "ValueTuple", "ConsoleColor", "Char", "GC", "Boolean"}
LINQ_NAMES = {"ToList", "ToArray", "Select", "Where", "Any", "All", "Count", "Min", "Max", "Sum", "Average", "First",
              "FirstOrDefault", "Last", "LastOrDefault", "Single", "SingleOrDefault", "Take", "Skip", "OrderBy",
              "OrderByDescending", "Zip", "DefaultIfEmpty", "Distinct", "Concat", "Reverse", "SelectMany", "Aggregate",
              "SequenceEqual", "ElementAt", "Contains", "ThenBy", "ThenByDescending", "ToDictionary", "Cast", "GroupBy",
              "TakeWhile", "SkipWhile", "ToHashSet", "OfType", "AsEnumerable", "LongCount"}

ENUMS = {
    "MidpointRounding": {"ToEven": 0, "AwayFromZero": 1},
    "StringSplitOptions": {"None": 0, "RemoveEmptyEntries": 1},
    "NumberStyles": {"None": 0, "Float": 167, "Any": 511, "Integer": 7, "AllowThousands": 64, "AllowDecimalPoint": 32,
                     "Number": 111, "AllowLeadingSign": 4},
    "StringComparison": {"CurrentCulture": 0, "CurrentCultureIgnoreCase": 1, "InvariantCulture": 2,
                         "InvariantCultureIgnoreCase": 3, "Ordinal": 4, "OrdinalIgnoreCase": 5},
    "ConsoleColor": {n: i for i, n in enumerate(
        ["Black", "DarkBlue", "DarkGreen", "DarkCyan", "DarkRed", "DarkMagenta", "DarkYellow", "Gray", "DarkGray",
         "Blue", "Green", "Cyan", "Red", "Magenta", "Yellow", "White"])},
}

CULTURE = object()


def bcl_static_member(interp, tname, name):
    if tname in ENUMS:
        return CsEnum(tname, name, ENUMS[tname][name])
    if tname in ("double", "Double", "float"):
        consts = {"MaxValue": 1.7976931348623157e308, "MinValue": -1.7976931348623157e308,
                  "PositiveInfinity": math.inf, "NegativeInfinity": -math.inf, "NaN": math.nan,
                  "Epsilon": 5e-324}
        if name in consts:
            return consts[name]
    if tname in ("int", "Int32"):
        if name == "MaxValue":
            return 2147483647
        if name == "MinValue":
            return INT_MIN
    if tname == "long":
        if name == "MaxValue":
            return 9223372036854775807
        if name == "MinValue":
            return -9223372036854775808
    if tname == "Math":
        if name == "PI":
            return math.pi
        if name == "E":
            return math.e
    if tname in ("string", "String") and name == "Empty":
        return ""
    if tname == "CultureInfo":
        return CULTURE
    if tname == "Environment" and name == "NewLine":
        return "\r\n"
    if tname == "DateTime" and name in ("Now", "UtcNow", "Today"):
        return CsDateTime(interp.now)
    if tname == "Console" and name == "Out":
        return ConsoleOut(interp)
    if tname == "Encoding":
        return CsEncoding(name)
    if tname == "IntPtr" and name == "Zero":
        return 0
    if tname == "AppDomain" and name == "CurrentDomain":
        return CsAppDomain()
    # a method group: Select(Math.Abs), Select(NumFormat.N3) ...
    return lambda *args: bcl_static_call(interp, tname, name, list(args), {})


class ConsoleOut:
    def __init__(self, interp):
        self.interp = interp


class CsDirectoryInfo:
    __slots__ = ("path",)

    def __init__(self, path):
        self.path = path


class CsAppDomain:
    pass


class CsEncoding:
    __slots__ = ("name",)

    def __init__(self, name):
        self.name = name

    def __eq__(self, o):
        return (isinstance(o, CsEncoding) and o.name == self.name) or o == self.name

    def __hash__(self):
        return hash(self.name)


class CsDateTime:
    def __init__(self, text):
        self.text = text


def _split(s, seps, remove_empty, count=None):
    if not seps:
        parts = re.split(r"\s", s)
    else:
        pat = "|".join(re.escape(x) for x in sorted(seps, key=len, reverse=True))
        parts = re.split(pat, s) if pat else [s]
    if remove_empty:
        parts = [p for p in parts if p != ""]
    return CsArray("string", (len(parts),), parts)


def bcl_static_call(interp, tname, name, args, named):
    if tname == "Math":
        a = args
        if name == "Abs":
            if a[0] == INT_MIN and type(a[0]) is int:
                raise CsException("OverflowException", "Negating the minimum value of a twos complement number is invalid.")
            return abs(a[0])
        if name == "Max":
            x, y = a
            if type(x) is float or type(y) is float:
                x, y = float(x), float(y)
                if x != x:
                    return x
                if y != y:
                    return y
                if x == y == 0:
                    return y if math.copysign(1.0, x) < 0 else x
            return x if x > y else y
        if name == "Min":
            x, y = a
            if type(x) is float or type(y) is float:
                x, y = float(x), float(y)
                if x != x:
                    return x
                if y != y:
                    return y
                if x == y == 0:
                    return x if math.copysign(1.0, x) < 0 else y
            return x if x < y else y
        if name == "Floor":
            x = float(a[0])
            return x if (x != x or x in (math.inf, -math.inf)) else float(math.floor(x)) if x != 0 else x
        if name == "Ceiling":
            x = float(a[0])
            if x != x or x in (math.inf, -math.inf) or x == 0:
                return x
            r = float(math.ceil(x))
            return math.copysign(0.0, x) if r == 0 else r
        if name == "Truncate":
            x = float(a[0])
            return x if (x != x or x in (math.inf, -math.inf)) else math.copysign(float(math.trunc(x)), x)
        if name == "Round":
            x = float(a[0])
            if len(a) == 1:
                return math_round(x)
            if len(a) == 2:
                if type(a[1]) is CsEnum:
                    return math_round_digits(x, 0, a[1].value == 1)
                return math_round_digits(x, a[1])
            return math_round_digits(x, a[1], a[2].value == 1)
        if name == "Sqrt":
            x = float(a[0])
            return math.sqrt(x) if x >= 0 else (math.nan if x == x else x)
        if name == "Pow":
            try:
                return math.pow(float(a[0]), float(a[1]))
            except OverflowError:
                return math.inf
            except ValueError:
                return math.nan
        if name == "Sign":
            if a[0] != a[0]:
                raise CsException("ArithmeticException", "Function does not accept floating point Not-a-Number values.")
            return (a[0] > 0) - (a[0] < 0)
        if name == "Exp":
            try:
                return math.exp(a[0])
            except OverflowError:
                return math.inf
        if name in ("Log", "Log10"):
            x = float(a[0])
            if x == 0:
                return -math.inf
            if x < 0 or x != x:
                return math.nan
            if name == "Log10":
                return math.log10(x)
            return math.log(x) if len(a) == 1 else math.log(x) / math.log(a[1])
        if name in ("Sin", "Cos", "Tan", "Atan", "Asin", "Acos", "Sinh", "Cosh", "Tanh"):
            return getattr(math, name.lower())(a[0])
        if name == "Atan2":
            return math.atan2(a[0], a[1])
    if tname == "Console":
        if name in ("WriteLine", "Write"):
            if not args:
                s = ""
            elif len(args) > 1 and type(args[0]) is str:
                rest = args[1:]
                if len(rest) == 1 and type(rest[0]) is CsArray:
                    rest = rest[0].data
                s = format_composite(args[0], rest)
            else:
                s = cs_tostring(args[0])
            w = interp.out_writer
            if w is None:
                interp.console.append(s + ("\r\n" if name == "WriteLine" else ""))
                return None
            # Console.SetOut(user TextWriter): TextWriter funnels everything it does not override into Write(char)
            def overload(mname, ptype):
                for d in w.cls.methods.get(mname, []):
                    if len(d[4]) == 1 and d[4][0][0][1] == ptype:
                        return d
                return None

            def write_text(text):
                d = overload("Write", "string")
                if d is not None:
                    interp.run_method(w.cls, d, 4, d[5], d[2], w, [text], {})
                    return
                d = overload("Write", "char")
                for ch in text:
                    interp.run_method(w.cls, d, 4, d[5], d[2], w, [ch], {})
            if name == "WriteLine" and args:
                d = overload("WriteLine", "string")
                if d is not None:
                    interp.run_method(w.cls, d, 4, d[5], d[2], w, [s], {})
                    return None
                write_text(s)
                write_text("\r\n")
                return None
            if name == "WriteLine":
                d = overload("Write", "char")
                if d is not None:
                    for ch in "\r\n":
                        interp.run_method(w.cls, d, 4, d[5], d[2], w, [ch], {})
                else:
                    write_text("\r\n")
                return None
            write_text(s)
            return None
        if name == "SetOut":
            interp.out_writer = args[0] if type(args[0]) is CsObject else None
            return None
        if name == "ReadLine":
            return interp.stdin.pop(0) if interp.stdin else None
        if name in ("ReadKey", "Clear", "ResetColor", "Beep", "SetCursorPosition"):
            return None
    if tname in ("string", "String"):
        if name == "Join":
            sep = args[0]
            if len(args) == 2 and type(args[1]) is not str:
                items = iterate(args[1])
            else:
                items = args[1:]
            return (sep or "").join(cs_tostring(v) for v in items)
        if name == "Format":
            fmt = args[0]
            if type(fmt) is not str and fmt is CULTURE:
                args = args[1:]
                fmt = args[0]
            rest = args[1:]
            if len(rest) == 1 and type(rest[0]) is CsArray:
                rest = rest[0].data
            return format_composite(fmt, rest)
        if name == "IsNullOrWhiteSpace":
            return args[0] is None or args[0].strip() == ""
        if name == "IsNullOrEmpty":
            return args[0] is None or args[0] == ""
        if name == "Concat":
            if len(args) == 1 and type(args[0]) is not str:
                return "".join(cs_tostring(v) for v in iterate(args[0]))
            return "".join(cs_tostring(v) for v in args)
        if name == "Equals":
            if len(args) > 2 and type(args[2]) is CsEnum and args[2].value in (1, 3, 5):
                return args[0] is not None and args[1] is not None and args[0].lower() == args[1].lower()
            return args[0] == args[1]
        if name == "Compare":
            return cs_compare(args[0], args[1])
    if tname in ("double", "Double", "float"):
        if name == "Parse":
            return parse_double(args[0])
        if name == "TryParse":
            try:
                v = parse_double(args[0])
            except CsException:
                args[-1].set(0.0)
                return False
            args[-1].set(v)
            return True
        if name == "IsNaN":
            return args[0] != args[0]
        if name == "IsInfinity":
            return args[0] in (math.inf, -math.inf)
        if name == "IsPositiveInfinity":
            return args[0] == math.inf
        if name == "IsNegativeInfinity":
            return args[0] == -math.inf
    if tname in ("int", "Int32", "long"):
        if name == "Parse":
            return parse_int(args[0])
        if name == "TryParse":
            try:
                v = parse_int(args[0])
            except CsException:
                args[-1].set(0)
                return False
            args[-1].set(v)
            return True
    if tname in ("bool", "Boolean") and name == "Parse":
        return args[0].strip().lower() == "true"
    if tname in ("char", "Char"):
        c = args[0]
        if name == "IsDigit":
            return c.isdigit()
        if name == "IsLetter":
            return c.isalpha()
        if name == "IsWhiteSpace":
            return c.isspace()
        if name == "IsLetterOrDigit":
            return c.isalnum()
        if name == "ToUpper":
            return c.upper()
        if name == "ToLower":
            return c.lower()
    if tname == "Enumerable":
        if name == "Range":
            if args[1] < 0:
                raise CsException("ArgumentOutOfRangeException", param="count")
            return seq(list(range(args[0], args[0] + args[1])), "int")
        if name == "Repeat":
            if args[1] < 0:
                raise CsException("ArgumentOutOfRangeException", param="count")
            return seq([args[0]] * args[1], "double" if type(args[0]) is float else None)
        if name == "Empty":
            return seq([])
        r = linq(name, args[0], args[1:])
        if r is not _MISSING:
            return r
    if tname == "Array":
        if name == "Copy":
            src, dst, n = args[0], args[1], args[-1]
            so = do = 0
            if len(args) == 5:
                so, do = args[1], args[3]
                dst = args[2]
            if src is None or dst is None:
                raise CsException("ArgumentNullException")
            if n < 0 or so + n > len(src.data) or do + n > len(dst.data):
                raise CsException("ArgumentException", "Source array was not long enough. Check srcIndex and length, "
                                                       "and the array's lower bounds.")
            dst.data[do:do + n] = src.data[so:so + n]
            return None
        if name == "Sort":
            arr = args[0]
            cmpf = args[1] if len(args) > 1 else cs_compare
            introsort(arr.data, cmpf)
            return None
        if name == "IndexOf":
            for i, v in enumerate(args[0].data):
                if cs_equals(v, args[1]):
                    return i
            return -1
        if name == "Reverse":
            args[0].data.reverse()
            return None
        if name == "Clear":
            arr, start, n = args
            for i in range(start, start + n):
                arr.data[i] = default_for_elem(arr.elem)
            return None
        if name == "Empty":
            return CsArray(None, (0,), [])
        if name == "Fill":
            args[0].data[:] = [args[1]] * len(args[0].data)
            return None
    if tname in ("Tuple", "ValueTuple") and name == "Create":
        return CsTuple(args)
    if tname == "Convert":
        v = args[0]
        if name == "ToDouble":
            return parse_double(v) if type(v) is str else float(v)
        if name == "ToInt32":
            if type(v) is str:
                return parse_int(v)
            if type(v) is float:
                r = math_round(v)
                if not INT_MIN <= r <= 2147483647:
                    raise CsException("OverflowException")
                return int(r)
            return int(v)
        if name == "ToString":
            return cs_tostring(v)
    if tname == "File":
        path = args[0]
        if name == "Exists":
            return path in interp.files or os.path.isfile(path)
        if name in ("ReadAllLines", "ReadAllText", "ReadLines"):
            if path in interp.files:
                text = interp.files[path]
            elif os.path.isfile(path):
                with open(path, encoding="utf-8-sig", newline="") as f:
                    text = f.read()
            else:
                raise CsException("FileNotFoundException", f"Could not find file '{path}'.")
            if name == "ReadAllText":
                return text
            if text.startswith("﻿"):
                text = text[1:]
            lines = re.split(r"\r\n|\r|\n", text)
            if lines and lines[-1] == "":
                lines.pop()
            return CsArray("string", (len(lines),), lines)
        # File.WriteAllText / AppendAllText(path, text, Encoding.UTF8): the StreamWriter emits the UTF-8 preamble when
        # it starts at stream position 0 (a new or empty file), never in the middle of an existing file
        bom = "\ufeff" if len(args) > 2 and args[2] == "UTF8" else ""
        if name == "WriteAllText":
            interp.files[path] = bom + args[1]
            return None
        if name == "AppendAllText":
            old = interp.files.get(path, "")
            interp.files[path] = old + (bom if old == "" else "") + args[1]
            return None
        if name in ("WriteAllLines", "AppendAllLines"):
            text = "".join(cs_tostring(v) + "\r\n" for v in iterate(args[1]))
            interp.files[path] = (interp.files.get(path, "") if name.startswith("Append") else "") + text
            return None
        if name == "Delete":
            interp.files.pop(path, None)
            return None
    if tname == "Path":
        if name == "Combine":
            return interp.path_sep.join(a.rstrip("\\/") if i < len(args) - 1 else a for i, a in enumerate(args))
        if name == "GetFileName":
            return re.split(r"[\\/]", args[0])[-1]
        if name == "GetFileNameWithoutExtension":
            return os.path.splitext(re.split(r"[\\/]", args[0])[-1])[0]
        if name == "GetExtension":
            return os.path.splitext(args[0])[1]
        if name == "GetDirectoryName":
            return re.sub(r"[\\/][^\\/]*$", "", args[0]) if re.search(r"[\\/]", args[0]) else ""
        if name == "GetFullPath":
            return args[0]
    if tname == "Directory":
        if name == "GetParent":
            p = args[0]
            if p.endswith(("\\", "/")):        # a trailing separator names the directory itself
                return CsDirectoryInfo(p[:-1])
            return CsDirectoryInfo(re.sub(r"[\\/][^\\/]*$", "", p))
        if name in ("CreateDirectory",):
            return None
        if name == "Exists":
            return True
        if name == "GetCurrentDirectory":
            return "."
    if tname == "Marshal":
        import ctypes
        if name == "PtrToStringAnsi":
            if not args[0]:
                return None
            raw = ctypes.string_at(args[0], args[1]) if len(args) > 1 else ctypes.string_at(args[0])
            return raw.decode("utf-8")
        if name == "Copy":          # Marshal.Copy(IntPtr source, byte[] destination, int startIndex, int length)
            src, dst, start, n = args
            dst.data[start:start + n] = list(ctypes.string_at(src, n))
            return None
    if tname == "IntPtr":
        raise NotImplementedError(f"IntPtr.{name}")
    if tname == "Environment" and name == "Exit":
        raise SystemExit(args[0])
    if tname == "GC":
        return None
    raise NotImplementedError(f"BCL static {tname}.{name}")


def bcl_construct(interp, name, ty, args, named):
    targs = ty[2]
    if name in ("List", "IList"):
        elem = type_name(targs[0]) if targs else None
        if targs and targs[0][0] == "tupletype":
            elem = targs[0]
        if args and type(args[0]) is not int:
            if args[0] is None:
                raise CsException("ArgumentNullException", param="collection")
            lst = CsList(list(iterate(args[0])), elem)
            if elem == "double":
                lst.items = [float(v) if type(v) is int else v for v in lst.items]
            return lst
        if args and args[0] < 0 or named.get("capacity", 0) < 0:
            raise CsException("ArgumentOutOfRangeException", param="capacity")
        return CsList([], elem, cap=(args[0] if args else named.get("capacity", 0)))
    if name == "Dictionary" or name == "SortedDictionary":
        d = CsDict()
        if args and type(args[0]) is CsDict:
            d.d = dict(args[0].d)
        return d
    if name == "HashSet":
        return CsSet(iterate(args[0])) if args and type(args[0]) is not int else CsSet()
    if name == "Stack":
        return CsStack(list(iterate(args[0])) if args and type(args[0]) is not int else None,
                       targs[0] if targs and targs[0][0] == "tupletype" else None)
    if name == "Queue":
        return CsQueue(list(iterate(args[0])) if args and type(args[0]) is not int else None,
                       targs[0] if targs and targs[0][0] == "tupletype" else None)
    if name == "StringBuilder":
        return CsStringBuilder(args[0] if args and type(args[0]) is str else "")
    if name == "StringWriter":
        return CsStringWriter()
    if name in ("string", "String"):
        if len(args) == 2:
            return args[0] * args[1]
        if type(args[0]) is CsArray:
            return "".join(args[0].data)
    if name in EXC_BASE or name.endswith("Exception"):
        msg = args[0] if args else None
        inner = args[1] if len(args) > 1 and isinstance(args[1], CsException) else None
        if name == "ArgumentNullException" and len(args) == 1:
            return CsException(name, None, param=args[0])
        if name in ("ArgumentException", "ArgumentOutOfRangeException") and len(args) == 2 and type(args[1]) is str:
            if name == "ArgumentOutOfRangeException":
                return CsException(name, args[1], param=args[0])
            return CsException(name, args[0], param=args[1])
        if name == "ArgumentOutOfRangeException" and len(args) == 1:
            return CsException(name, "Specified argument was out of the range of valid values.", param=args[0])
        return CsException(name, msg, inner)
    if name in ("Tuple", "ValueTuple"):
        return CsTuple(args)
    if name == "Random":
        raise NotImplementedError("System.Random is not modelled (its sequence is an implementation detail)")
    if name == "StreamWriter":
        return CsStreamWriter(interp, args[0], bool(args[1]) if len(args) > 1 and type(args[1]) is bool else False)
    if name in ("UTF8Encoding",):
        return "UTF8"
    raise NotImplementedError(f"BCL type {name}")


class CsStreamWriter(CsStringBuilder):
    __slots__ = ("interp", "path")

    def __init__(self, interp, path, append):
        CsStringBuilder.__init__(self)
        self.interp, self.path = interp, path
        if not append:
            interp.files[path] = ""


def bcl_instance_member(interp, obj, name):
    t = type(obj)
    if name == "Count":
        if t is CsList or t is CsStack or t is CsQueue:
            return len(obj.items)
        if t is CsDict:
            return len(obj.d)
        if t is CsSet:
            return len(obj.s)
        if t is CsSeq:
            return len(obj)
    if name == "Length":
        if t is CsArray:
            return len(obj.data)
        if t is str:
            return len(obj)
        if t is CsStringBuilder:
            return sum(len(p) for p in obj.parts)
    if t is CsArray and name == "Rank":
        return len(obj.dims)
    if t is CsTuple:
        return obj.member(name)
    if t is CsAnon:
        return obj.f[name]
    if t is float or t is int or t is bool:
        if name == "Value":
            return obj
        if name == "HasValue":
            return True
    if t is CsDict:
        if name == "Keys":
            return seq(list(obj.d.keys()))
        if name == "Values":
            return seq(list(obj.d.values()))
    if isinstance(obj, CsException):
        if name == "Message":
            return obj.message
        if name == "InnerException":
            return obj.inner
        if name == "StackTrace":
            return ""
        if name == "ParamName":
            return None
    if t is CsList and name == "Capacity":
        return len(obj.items)
    if t is CsSeq and name == "Key":
        return obj.Key
    if t is CsDateTime:
        return obj
    if t is ConsoleOut and name == "Encoding":
        return CsEncoding("UTF8")
    if t is CsAppDomain and name == "BaseDirectory":
        return interp.base_dir
    if t is CsDirectoryInfo:
        if name == "FullName":
            return obj.path
        if name == "Parent":
            return CsDirectoryInfo(re.sub(r"[\\/][^\\/]*$", "", obj.path))
        if name == "Name":
            return re.split(r"[\\/]", obj.path)[-1]
    # a method group on an instance: list.Add, sb.Append ...
    return lambda *args: bcl_instance_call(interp, obj, name, list(args), {})


def nullable_member(name):
    if name == "HasValue":
        return False
    if name == "Value":
        raise CsException("InvalidOperationException", "Nullable object must have a value.")
    raise null_ref()


def bcl_instance_call(interp, obj, name, args, named):
    t = type(obj)
    if t is CsList:
        items = obj.items
        if name == "Add":
            items.append(obj.co(args[0])); obj.ver += 1
            obj.capacity()
            return None
        if name == "AddRange":
            if args[0] is None:
                raise CsException("ArgumentNullException", param="collection")
            items.extend(obj.co(v) for v in list(iterate(args[0]))); obj.ver += 1
            return None
        if name == "Insert":
            i = args[0]
            if i < 0 or i > len(items):
                raise CsException("ArgumentOutOfRangeException", "Index must be within the bounds of the List.", param="index")
            items.insert(i, obj.co(args[1])); obj.ver += 1
            return None
        if name == "RemoveAt":
            i = args[0]
            if i < 0 or i >= len(items):
                raise CsException("ArgumentOutOfRangeException", param="index")
            del items[i]; obj.ver += 1
            return None
        if name == "Remove":
            for i, v in enumerate(items):
                if cs_equals(v, args[0]):
                    del items[i]; obj.ver += 1
                    return True
            return False
        if name == "RemoveAll":
            keep = [v for v in items if not args[0](v)]
            n = len(items) - len(keep)
            obj.items[:] = keep; obj.ver += 1
            return n
        if name == "RemoveRange":
            i, n = args
            if i < 0 or n < 0 or i + n > len(items):
                raise CsException("ArgumentException", "Offset and length were out of bounds for the array or count is "
                                                       "greater than the number of elements from index to the end of the "
                                                       "source collection.")
            del items[i:i + n]; obj.ver += 1
            return None
        if name == "Clear":
            items.clear(); obj.ver += 1
            return None
        if name == "Contains":
            return any(cs_equals(v, args[0]) for v in items)
        if name == "IndexOf":
            for i, v in enumerate(items):
                if cs_equals(v, args[0]):
                    return i
            return -1
        if name == "LastIndexOf":
            for i in range(len(items) - 1, -1, -1):
                if cs_equals(items[i], args[0]):
                    return i
            return -1
        if name == "Sort":
            if args and callable(args[0]):
                introsort(items, args[0], obj.capacity())
            else:
                introsort(items, cs_compare, obj.capacity())
            obj.ver += 1
            return None
        if name == "Reverse" and not args:
            items.reverse(); obj.ver += 1
            return None
        if name == "ToArray":
            return CsArray(obj.elem, (len(items),), list(items))
        if name == "AsReadOnly":
            return obj
        if name == "GetRange":
            i, n = args
            if i < 0 or n < 0 or i + n > len(items):
                raise CsException("ArgumentException", "Offset and length were out of bounds for the array or count is "
                                                       "greater than the number of elements from index to the end of the "
                                                       "source collection.")
            return CsList(items[i:i + n], obj.elem)
        if name == "ForEach":
            for v in list(items):
                args[0](v)
            return None
        if name == "Exists":
            return any(args[0](v) for v in items)
        if name == "TrueForAll":
            return all(args[0](v) for v in items)
        if name == "Find":
            for v in items:
                if args[0](v):
                    return v
            return default_for_elem(obj.elem)
        if name == "FindAll":
            return CsList([v for v in items if args[0](v)], obj.elem)
        if name == "FindIndex":
            for i, v in enumerate(items):
                if args[0](v):
                    return i
            return -1
        if name == "FindLastIndex":
            for i in range(len(items) - 1, -1, -1):
                if args[0](items[i]):
                    return i
            return -1
        if name == "CopyTo":
            dst = args[0]
            off = args[1] if len(args) > 1 else 0
            dst.data[off:off + len(items)] = items
            return None
        if name == "ConvertAll":
            return CsList([args[0](v) for v in items])
    elif t is CsArray:
        if name == "GetLength":
            d = args[0]
            if d < 0 or d >= len(obj.dims):
                raise CsException("IndexOutOfRangeException")
            return obj.dims[d]
        if name == "Clone":
            return CsArray(obj.elem, obj.dims, list(obj.data))
        if name == "GetUpperBound":
            return obj.dims[args[0]] - 1
        if name == "GetLowerBound":
            return 0
        if name == "CopyTo":
            dst, off = args
            dst.data[off:off + len(obj.data)] = obj.data
            return None
        if name == "Contains":
            return any(cs_equals(v, args[0]) for v in obj.data)
        if name == "GetValue":
            return obj.get(args)
        if name == "SetValue":
            obj.set(args[1:], args[0])
            return None
    elif t is str:
        r = str_call(obj, name, args)
        if r is not _MISSING:
            return r
    elif t is float or t is int:
        if name == "ToString":
            fmt = args[0] if args and type(args[0]) is str else None
            return cs_tostring(obj, fmt)
        if name == "CompareTo":
            return cs_compare(obj, float(args[0]) if t is float else args[0])
        if name == "Equals":
            return cs_equals(obj, args[0])
        if name == "GetValueOrDefault":
            return obj
        if name == "GetHashCode":
            return hash(obj) & 0x7FFFFFFF
    elif t is bool:
        if name == "ToString":
            return cs_tostring(obj)
        if name == "GetValueOrDefault":
            return obj
    elif t is CsStringBuilder or t is CsStringWriter or t is CsStreamWriter:
        if name in ("Append", "Write"):
            if len(args) == 2 and type(args[0]) is str and type(args[1]) is int and len(args[0]) == 1 and t is CsStringBuilder:
                obj.parts.append(args[0] * args[1])
            elif len(args) > 1 and type(args[0]) is str:
                obj.parts.append(format_composite(args[0], args[1:]))
            else:
                obj.parts.append(cs_tostring(args[0]))
            return obj
        if name in ("AppendLine", "WriteLine"):
            if len(args) > 1 and type(args[0]) is str:
                obj.parts.append(format_composite(args[0], args[1:]) + "\r\n")
            else:
                obj.parts.append((cs_tostring(args[0]) if args else "") + "\r\n")
            return obj
        if name == "AppendFormat":
            obj.parts.append(format_composite(args[0], args[1:]))
            return obj
        if name == "ToString":
            return "".join(obj.parts)
        if name == "Clear":
            obj.parts.clear()
            return obj
        if name == "Insert":
            s = "".join(obj.parts)
            obj.parts[:] = [s[:args[0]] + cs_tostring(args[1]) + s[args[0]:]]
            return obj
        if name == "GetStringBuilder":
            return obj
        if name in ("Flush", "Close", "Dispose"):
            if t is CsStreamWriter:
                obj.interp.files[obj.path] = obj.interp.files.get(obj.path, "") + "".join(obj.parts)
                obj.parts.clear()
            return None
    elif t is CsDict:
        d = obj.d
        if name == "ContainsKey":
            if args[0] is None:
                raise CsException("ArgumentNullException", param="key")
            return args[0] in d
        if name == "Add":
            if args[0] in d:
                raise CsException("ArgumentException", "An item with the same key has already been added.")
            d[args[0]] = args[1]
            return None
        if name == "Remove":
            return d.pop(args[0], _MISSING) is not _MISSING
        if name == "TryGetValue":
            if args[0] in d:
                args[1].set(d[args[0]])
                return True
            args[1].set(None)
            return False
        if name == "Clear":
            d.clear()
            return None
        if name == "ContainsValue":
            return any(cs_equals(v, args[0]) for v in d.values())
    elif t is CsSet:
        if name == "Add":
            if args[0] in obj.s:
                return False
            obj.s[args[0]] = None
            return True
        if name == "Contains":
            return args[0] in obj.s
        if name == "Remove":
            return obj.s.pop(args[0], _MISSING) is not _MISSING
        if name == "Clear":
            obj.s.clear()
            return None
        if name == "UnionWith":
            for v in iterate(args[0]):
                obj.s[v] = None
            return None
    elif t is CsStack:
        if name == "Push":
            obj.items.append(coerce(args[0], obj.elem) if obj.elem is not None else args[0])
            return None
        if name == "Pop" or name == "Peek":
            if not obj.items:
                raise CsException("InvalidOperationException", "Stack empty.")
            return obj.items.pop() if name == "Pop" else obj.items[-1]
        if name == "Clear":
            obj.items.clear()
            return None
        if name == "Contains":
            return any(cs_equals(v, args[0]) for v in obj.items)
    elif t is CsQueue:
        if name == "Enqueue":
            obj.items.append(coerce(args[0], obj.elem) if obj.elem is not None else args[0])
            return None
        if name == "Dequeue" or name == "Peek":
            if not obj.items:
                raise CsException("InvalidOperationException", "Queue empty.")
            return obj.items.pop(0) if name == "Dequeue" else obj.items[0]
        if name == "Clear":
            obj.items.clear()
            return None
    elif t is CsTuple:
        if name == "ToString":
            return cs_tostring(obj)
        if name == "Equals":
            return obj == args[0]
    elif isinstance(obj, CsException):
        if name == "ToString":
            return obj.cs_tostring()
    elif t is CsEnum:
        if name == "ToString":
            return obj.name
        if name == "Equals":
            return obj == args[0]
        if name == "HasFlag":
            return (obj.value & args[0].value) == args[0].value
    elif t is CsDateTime:
        if name == "ToString":
            return obj.text
    elif t is ConsoleOut:
        if name in ("Write", "WriteLine"):      # the original Console.Out: never through Console.SetOut's writer
            saved, interp.out_writer = interp.out_writer, None
            try:
                return bcl_static_call(interp, "Console", name, args, named)
            finally:
                interp.out_writer = saved
        return bcl_static_call(interp, "Console", name, args, named)
    elif t is CsEncoding:
        if name == "GetString":
            data = args[0].data
            if len(args) == 3:
                if args[1] < 0 or args[2] < 0 or args[1] + args[2] > len(data):
                    raise CsException("ArgumentOutOfRangeException", param="count")
                data = data[args[1]:args[1] + args[2]]
            return bytes(data).decode("utf-8", errors="replace")
        if name == "GetBytes":
            b = list(args[0].encode("utf-8"))
            return CsArray("byte", (len(b),), b)
        if name == "GetPreamble":
            b = [0xEF, 0xBB, 0xBF] if obj.name == "UTF8" else []
            return CsArray("byte", (len(b),), b)
    elif t is CsAnon:
        if name == "ToString":
            return cs_tostring(obj)
    elif callable(obj) and name == "Invoke":
        return obj(*args)
    if name == "ToString" and not args:
        return cs_tostring(obj)
    if name == "Equals" and len(args) == 1:
        return cs_equals(obj, args[0])
    if name == "GetHashCode":
        return id(obj) & 0x7FFFFFFF
    if name == "GetType":
        return TypeVal(t.__name__)
    if name == "Dispose":
        return None
    if name in LINQ_NAMES:
        r = linq(name, obj, args)
        if r is not _MISSING:
            return r
    raise NotImplementedError(f"BCL instance {t.__name__}.{name}")


def str_call(s, name, args):
    if name == "Trim":
        if args:
            chars = args[0].data if type(args[0]) is CsArray else args
            return s.strip("".join(chars))
        return s.strip()
    if name == "TrimEnd":
        if args:
            chars = args[0].data if type(args[0]) is CsArray else args
            return s.rstrip("".join(chars))
        return s.rstrip()
    if name == "TrimStart":
        if args:
            chars = args[0].data if type(args[0]) is CsArray else args
            return s.lstrip("".join(chars))
        return s.lstrip()
    if name == "Split":
        seps, opts, count = [], 0, None
        for a in args:
            if type(a) is CsArray:
                seps.extend(a.data)
            elif type(a) is str:
                seps.append(a)
            elif type(a) is CsEnum:
                opts = a.value
            elif type(a) is int:
                count = a
        return _split(s, seps, opts == 1, count)
    if name in ("ToLower", "ToLowerInvariant"):
        return s.lower()
    if name in ("ToUpper", "ToUpperInvariant"):
        return s.upper()
    if name == "StartsWith":
        if len(args) > 1 and type(args[1]) is CsEnum and args[1].value in (1, 3, 5):
            return s.lower().startswith(args[0].lower())
        return s.startswith(args[0])
    if name == "EndsWith":
        if len(args) > 1 and type(args[1]) is CsEnum and args[1].value in (1, 3, 5):
            return s.lower().endswith(args[0].lower())
        return s.endswith(args[0])
    if name == "Contains":
        return args[0] in s
    if name == "Replace":
        return s.replace(args[0], args[1])
    if name == "Substring":
        start = args[0]
        n = args[1] if len(args) > 1 else len(s) - start
        if start < 0 or n < 0 or start + n > len(s):
            raise CsException("ArgumentOutOfRangeException", param="startIndex" if start < 0 or start > len(s) else "length")
        return s[start:start + n]
    if name == "IndexOf":
        if len(args) > 1 and type(args[-1]) is CsEnum and args[-1].value in (1, 3, 5):
            return s.lower().find(args[0].lower())
        return s.find(args[0], *(args[1:2] if len(args) > 1 and type(args[1]) is int else []))
    if name == "LastIndexOf":
        return s.rfind(args[0])
    if name == "PadLeft":
        return s.rjust(args[0], args[1] if len(args) > 1 else " ")
    if name == "PadRight":
        return s.ljust(args[0], args[1] if len(args) > 1 else " ")
    if name == "Equals":
        if args[0] is None:
            return False
        if len(args) > 1 and type(args[1]) is CsEnum and args[1].value in (1, 3, 5):
            return s.lower() == args[0].lower()
        return s == args[0]
    if name == "ToString":
        return s
    if name == "CompareTo":
        return cs_compare(s, args[0])
    if name == "ToCharArray":
        return CsArray("char", (len(s),), list(s))
    if name == "Insert":
        return s[:args[0]] + args[1] + s[args[0]:]
    if name == "Remove":
        return s[:args[0]] + (s[args[0] + args[1]:] if len(args) > 1 else "")
    if name == "GetHashCode":
        return hash(s) & 0x7FFFFFFF
    return _MISSING


# ------------------------------------------------------------------------------------------- Python <-> C# values
def to_list(values, elem="double"):
    if elem == "double":
        return CsList([float(v) for v in values], "double")
    return CsList(list(values), elem)


def to_array(values, elem="double"):
    vals = [float(v) for v in values] if elem == "double" else list(values)
    return CsArray(elem, (len(vals),), vals)


def to_array2d(rows, elem="double"):
    r = len(rows)
    c = len(rows[0]) if r else 0
    flat = [float(v) if elem == "double" else v for row in rows for v in row]
    return CsArray(elem, (r, c), flat)


def from_cs(v):
    """C# value -> nested Python lists / scalars"""
    t = type(v)
    if t is CsList:
        return [from_cs(x) for x in v.items]
    if t is CsArray:
        if len(v.dims) == 1:
            return [from_cs(x) for x in v.data]
        r, c = v.dims
        return [[from_cs(v.data[i * c + j]) for j in range(c)] for i in range(r)]
    if t is CsSeq:
        return [from_cs(x) for x in v]
    if t is CsTuple:
        return tuple(from_cs(x) for x in v.vals)
    return v
