"""Lexer and recursive-descent parser for the C# 7.3 subset used by Storm-Tarran/LPR_381_Group_V22.

TEST INFRASTRUCTURE ONLY (see oracle/csharp/__init__.py).  The grammar follows the C# language specification
(ECMA-334): operator precedence table of section 12.4.2, the cast-expression disambiguation rule of 12.9.7, the
generic "<" disambiguation of 6.2.5 (a type-argument list is accepted in an expression only when "(" follows),
interpolated strings (12.8.3), tuples and deconstruction, expression-bodied members, named / optional arguments.
The AST is made of plain tuples, `(kind, ...)`; csrun.py documents the kinds it evaluates.
"""
import re

KEYWORDS = {
    "abstract", "as", "base", "bool", "break", "byte", "case", "catch", "char", "checked", "class", "const",
    "continue", "decimal", "default", "delegate", "do", "double", "else", "enum", "event", "explicit", "extern",
    "false", "finally", "fixed", "float", "for", "foreach", "goto", "if", "implicit", "in", "int", "interface",
    "internal", "is", "lock", "long", "namespace", "new", "null", "object", "operator", "out", "override",
    "params", "private", "protected", "public", "readonly", "ref", "return", "sbyte", "sealed", "short",
    "sizeof", "stackalloc", "static", "string", "struct", "switch", "this", "throw", "true", "try", "typeof",
    "uint", "ulong", "unchecked", "unsafe", "ushort", "using", "virtual", "void", "volatile", "while",
}
PREDEFINED = {"bool", "byte", "char", "decimal", "double", "float", "int", "long", "object", "sbyte", "short",
              "string", "uint", "ulong", "ushort", "void"}
MODIFIERS = {"public", "private", "protected", "internal", "static", "readonly", "sealed", "abstract", "virtual",
             "override", "const", "extern", "unsafe", "volatile"}

OPS = ["??=", "<<=", ">>=", "=>", "?.", "??", "++", "--", "&&", "||", "==", "!=", "<=", ">=", "+=", "-=", "*=", "/=",
       "%=", "&=", "|=", "^=", "<<", "+", "-", "*", "/", "%", "&", "|", "^", "!", "~", "=", "<", ">", "?", ":", ".",
       ",", ";", "(", ")", "[", "]", "{", "}"]

_ID = re.compile(r"[A-Za-z_\u0080-\uffff][A-Za-z0-9_\u0080-\uffff]*")
_NUM = re.compile(r"0[xX][0-9a-fA-F_]+[uUlL]*|(?:\d[\d_]*)?\.\d[\d_]*(?:[eE][+-]?\d+)?[dDfFmM]?|\d[\d_]*[eE][+-]?\d+[dDfFmM]?"
                  r"|\d[\d_]*[dDfFmMuUlL]*")
_ESC = {"n": "\n", "t": "\t", "r": "\r", "0": "\0", "\\": "\\", '"': '"', "'": "'", "a": "\a", "b": "\b",
        "f": "\f", "v": "\v"}


class CsSyntaxError(Exception):
    pass


class Tok:
    __slots__ = ("kind", "val", "line", "adj")

    def __init__(self, kind, val, line, adj=False):
        self.kind, self.val, self.line, self.adj = kind, val, line, adj

    def __repr__(self):
        return f"Tok({self.kind},{self.val!r},l{self.line})"


def _unescape(s, i, quote):
    """regular string / char literal body starting at s[i]; returns (text, index after the closing quote)"""
    out = []
    while True:
        ch = s[i]
        if ch == quote:
            return "".join(out), i + 1
        if ch == "\\":
            e = s[i + 1]
            if e == "u":
                out.append(chr(int(s[i + 2:i + 6], 16))); i += 6
            elif e == "x":
                m = re.match(r"[0-9a-fA-F]{1,4}", s[i + 2:])
                out.append(chr(int(m.group(0), 16))); i += 2 + len(m.group(0))
            else:
                out.append(_ESC[e]); i += 2
        else:
            out.append(ch); i += 1


def _interp_parts(s, i, verbatim):
    """body of an interpolated string starting after the opening quote; returns (parts, index after the quote)"""
    parts, lit = [], []
    while True:
        ch = s[i]
        if ch == '"':
            if verbatim and s[i + 1:i + 2] == '"':
                lit.append('"'); i += 2; continue
            if lit:
                parts.append("".join(lit))
            return parts, i + 1
        if ch == "{":
            if s[i + 1] == "{":
                lit.append("{"); i += 2; continue
            if lit:
                parts.append("".join(lit)); lit = []
            # find the end of the hole: track nesting and nested string / char literals
            j, depth = i + 1, 0
            split_comma = split_colon = None
            while True:
                c = s[j]
                if c == '"' or (c in "$@" and s[j + 1] in '"$@'):
                    k = j
                    while s[k] in "$@":
                        k += 1
                    if "$" in s[j:k]:
                        _, j = _interp_parts(s, k + 1, "@" in s[j:k])
                    elif "@" in s[j:k]:
                        k += 1
                        while not (s[k] == '"' and s[k + 1] != '"'):
                            k += 2 if s[k] == '"' else 1
                        j = k + 1
                    else:
                        _, j = _unescape(s, k + 1, '"')
                    continue
                if c == "'":
                    _, j = _unescape(s, j + 1, "'"); continue
                if c in "([{":
                    depth += 1
                elif c in ")]":
                    depth -= 1
                elif c == "}":
                    if depth == 0:
                        break
                    depth -= 1
                elif depth == 0 and c == "," and split_comma is None and split_colon is None:
                    split_comma = j
                elif depth == 0 and c == ":" and split_colon is None:
                    split_colon = j
                j += 1
            end = j
            fmt = None
            expr_end = end
            if split_colon is not None:
                fmt = s[split_colon + 1:end]; expr_end = split_colon
            align = None
            if split_comma is not None:
                align = s[split_comma + 1:expr_end]; expr_end = split_comma
            parts.append((s[i + 1:expr_end], align, fmt))
            i = end + 1
            continue
        if ch == "}":
            if s[i + 1] == "}":
                lit.append("}"); i += 2; continue
            raise CsSyntaxError("stray } in interpolated string")
        if ch == "\\" and not verbatim:
            e = s[i + 1]
            if e == "u":
                lit.append(chr(int(s[i + 2:i + 6], 16))); i += 6
            else:
                lit.append(_ESC[e]); i += 2
            continue
        lit.append(ch); i += 1


def tokenize(s):
    toks, i, n, line = [], 0, len(s), 1
    if s.startswith("﻿"):
        i = 1
    while i < n:
        ch = s[i]
        if ch == "\n":
            line += 1; i += 1; continue
        if ch in " \t\r\f\v":
            i += 1; continue
        if ch == "/" and s[i + 1:i + 2] == "/":
            j = s.find("\n", i)
            i = n if j < 0 else j
            continue
        if ch == "/" and s[i + 1:i + 2] == "*":
            j = s.index("*/", i + 2)
            line += s.count("\n", i, j); i = j + 2; continue
        if ch == "#":  # preprocessor line (#region ...)
            j = s.find("\n", i)
            i = n if j < 0 else j
            continue
        if ch == '"':
            text, j = _unescape(s, i + 1, '"')
            toks.append(Tok("str", text, line)); line += s.count("\n", i, j); i = j; continue
        if ch in "$@" and i + 1 < n and (s[i + 1] == '"' or (s[i + 1] in "$@" and s[i + 2:i + 3] == '"')):
            k = i
            while s[k] in "$@":
                k += 1
            pre = s[i:k]
            if "$" in pre:
                parts, j = _interp_parts(s, k + 1, "@" in pre)
                toks.append(Tok("istr", parts, line))
            else:
                out, j = [], k + 1
                while True:
                    if s[j] == '"':
                        if s[j + 1:j + 2] == '"':
                            out.append('"'); j += 2; continue
                        j += 1; break
                    out.append(s[j]); j += 1
                toks.append(Tok("str", "".join(out), line))
            line += s.count("\n", i, j); i = j; continue
        if ch == "'":
            text, j = _unescape(s, i + 1, "'")
            toks.append(Tok("char", text, line)); i = j; continue
        if ch == "@" and _ID.match(s, i + 1):
            m = _ID.match(s, i + 1)
            toks.append(Tok("id", m.group(0), line)); i = m.end(); continue
        m = _ID.match(s, i)
        if m:
            w = m.group(0)
            toks.append(Tok("kw" if w in KEYWORDS else "id", w, line)); i = m.end(); continue
        if ch.isdigit() or (ch == "." and s[i + 1:i + 2].isdigit()):
            m = _NUM.match(s, i)
            t = m.group(0).replace("_", "")
            low = t.lower()
            if low.startswith("0x"):
                toks.append(Tok("int", int(low.rstrip("ul"), 16), line))
            elif low[-1] in "dfm" or "." in low or "e" in low:
                toks.append(Tok("real", float(low.rstrip("dfm")), line))
            else:
                toks.append(Tok("int", int(low.rstrip("ul")), line))
            i = m.end(); continue
        for op in OPS:
            if s.startswith(op, i):
                adj = bool(toks) and toks[-1].kind == "op" and toks[-1].val == ">" and s[i - 1] == ">"
                toks.append(Tok("op", op, line, adj)); i += len(op); break
        else:
            raise CsSyntaxError(f"line {line}: unexpected character {ch!r}")
    toks.append(Tok("eof", None, line))
    return toks


BINARY_PREC = [
    ["||"], ["&&"], ["|"], ["^"], ["&"], ["==", "!="], ["<", ">", "<=", ">=", "is", "as"], ["<<", ">>"],
    ["+", "-"], ["*", "/", "%"],
]
ASSIGN_OPS = {"=", "+=", "-=", "*=", "/=", "%=", "&=", "|=", "^=", "<<=", ">>=", "??="}


class Parser:
    def __init__(self, src, name="<cs>"):
        self.name = name
        self.toks = tokenize(src)
        self.p = 0

    # ------------------------------------------------------------------ token helpers
    def peek(self, k=0):
        return self.toks[min(self.p + k, len(self.toks) - 1)]

    def at(self, val, k=0):
        t = self.peek(k)
        return t.kind in ("op", "kw") and t.val == val

    def at_id(self, k=0):
        return self.peek(k).kind == "id"

    def accept(self, val):
        if self.at(val):
            self.p += 1
            return True
        return False

    def expect(self, val):
        if not self.accept(val):
            t = self.peek()
            raise CsSyntaxError(f"{self.name}:{t.line}: expected {val!r}, found {t.val!r}")

    def ident(self):
        t = self.peek()
        if t.kind != "id":
            raise CsSyntaxError(f"{self.name}:{t.line}: expected identifier, found {t.val!r}")
        self.p += 1
        return t.val

    def fail(self, msg):
        t = self.peek()
        raise CsSyntaxError(f"{self.name}:{t.line}: {msg} (at {t.val!r})")

    # ------------------------------------------------------------------ types
    def parse_type(self):
        if self.at("("):
            self.p += 1
            elems = []
            while True:
                t = self.parse_type()
                nm = self.ident() if self.at_id() else None
                elems.append((t, nm))
                if not self.accept(","):
                    break
            self.expect(")")
            if len(elems) < 2:
                self.fail("tuple type needs two elements")
            base = ["tupletype", elems, [], False]
        else:
            t = self.peek()
            if t.kind == "kw" and t.val in PREDEFINED:
                self.p += 1
                base = ["type", t.val, [], [], False]
            elif t.kind == "id":
                self.p += 1
                name = t.val
                args = []
                if self.at("<"):
                    args = self.type_args()
                while self.at(".") and self.at_id(1):
                    self.p += 1
                    name = name + "." + self.ident()
                    if self.at("<"):
                        args = self.type_args()
                base = ["type", name, args, [], False]
            else:
                self.fail("expected type")
        # nullable / array suffixes
        while True:
            if self.at("?") and not base[-1]:
                self.p += 1
                base[-1] = True
            elif self.at("[") and (self.at("]", 1) or self.at(",", 1)):
                self.p += 1
                rank = 1
                while self.accept(","):
                    rank += 1
                self.expect("]")
                base[-2].append(rank)
            else:
                break
        return tuple(base)

    def type_args(self):
        self.expect("<")
        args = [self.parse_type()]
        while self.accept(","):
            args.append(self.parse_type())
        self.expect(">")
        return args

    def try_type(self):
        save = self.p
        try:
            return self.parse_type()
        except CsSyntaxError:
            self.p = save
            return None

    # ------------------------------------------------------------------ compilation unit
    def parse_unit(self):
        usings, classes = [], []
        self.namespace_body(usings, classes, "", top=True)
        return ("unit", usings, classes)

    def namespace_body(self, usings, classes, ns, top=False):
        while not self.at("}") and self.peek().kind != "eof":
            if self.at("using"):
                self.p += 1
                if self.accept("static"):
                    usings.append(("using_static", self.qualified()))
                elif self.at_id() and self.at("=", 1):
                    alias = self.ident(); self.p += 1
                    usings.append(("using_alias", alias, self.qualified()))
                else:
                    usings.append(("using", self.qualified()))
                self.expect(";")
            elif self.at("namespace"):
                self.p += 1
                name = self.qualified()
                self.expect("{")
                self.namespace_body(usings, classes, name)
                self.expect("}")
            elif self.at("["):
                self.skip_attribute()
            else:
                classes.append(self.type_decl(ns))
        if top and self.peek().kind != "eof":
            self.fail("unexpected }")

    def qualified(self):
        name = self.ident()
        while self.accept("."):
            name += "." + self.ident()
        return name

    def skip_attribute(self):
        self.expect("[")
        depth = 1
        while depth:
            t = self.peek(); self.p += 1
            if t.kind == "op" and t.val == "[":
                depth += 1
            elif t.kind == "op" and t.val == "]":
                depth -= 1

    def modifiers(self):
        mods = []
        while True:
            t = self.peek()
            if t.kind == "kw" and t.val in MODIFIERS:
                mods.append(t.val); self.p += 1
            elif t.kind == "id" and t.val in ("partial", "async") and self.peek(1).kind in ("kw", "id"):
                mods.append(t.val); self.p += 1
            else:
                return mods

    def type_decl(self, ns, mods=None):
        while self.at("["):
            self.skip_attribute()
        mods = self.modifiers() if mods is None else mods
        t = self.peek()
        if t.val == "enum":
            self.p += 1
            name = self.ident()
            if self.accept(":"):
                self.parse_type()
            self.expect("{")
            members = []
            while not self.at("}"):
                mn = self.ident()
                val = self.expr() if self.accept("=") else None
                members.append((mn, val))
                if not self.accept(","):
                    break
            self.expect("}")
            self.accept(";")
            return ("enum", name, members, ns)
        if t.val not in ("class", "struct", "interface"):
            self.fail("expected type declaration")
        self.p += 1
        name = self.ident()
        if self.at("<"):
            self.type_args()
        bases = []
        if self.accept(":"):
            bases.append(self.parse_type())
            while self.accept(","):
                bases.append(self.parse_type())
        self.expect("{")
        members = []
        while not self.at("}"):
            members.append(self.member(name))
        self.expect("}")
        self.accept(";")
        return ("class", name, mods, bases, members, t.val, ns)

    def params(self):
        self.expect("(")
        ps = []
        while not self.at(")"):
            while self.at("["):
                self.skip_attribute()
            mod = None
            if self.peek().val in ("out", "ref", "params", "this", "in"):
                mod = self.peek().val; self.p += 1
            ty = self.parse_type()
            nm = self.ident()
            default = self.expr() if self.accept("=") else None
            ps.append((ty, nm, default, mod))
            if not self.accept(","):
                break
        self.expect(")")
        return ps

    def member(self, cls_name):
        while self.at("["):
            self.skip_attribute()
        mods = self.modifiers()
        t = self.peek()
        if t.val in ("class", "struct", "interface", "enum"):
            return self.type_decl("", mods)
        # constructor: Name (
        if t.kind == "id" and t.val == cls_name and self.at("(", 1):
            self.p += 1
            ps = self.params()
            init = None
            if self.accept(":"):
                which = self.peek().val; self.p += 1
                init = (which, self.arguments())
            body = self.block()
            return ("ctor", mods, cls_name, ps, body, init)
        ty = self.parse_type()
        if self.at("operator"):
            self.fail("operator overloads are not supported")
        name = self.ident()
        if self.at("<") and not self.at("=", 1):
            self.type_args()
        if self.at("("):
            ps = self.params()
            if self.accept("=>"):
                e = self.expr(); self.expect(";")
                body = ("exprbody", e)
            elif self.accept(";"):
                body = None
            else:
                body = self.block()
            return ("method", mods, ty, name, ps, body)
        if self.at("{") or self.at("=>"):
            # property
            getter = setter = None
            init = None
            if self.accept("=>"):
                getter = ("exprbody", self.expr()); self.expect(";")
            else:
                self.expect("{")
                while not self.at("}"):
                    self.modifiers()
                    acc = self.ident()
                    if self.accept(";"):
                        body = "auto"
                    elif self.accept("=>"):
                        body = ("exprbody", self.expr()); self.expect(";")
                    else:
                        body = self.block()
                    if acc == "get":
                        getter = body
                    elif acc == "set":
                        setter = body
                    else:
                        self.fail("expected get or set")
                self.expect("}")
                if self.accept("="):
                    init = self.expr(); self.expect(";")
            return ("property", mods, ty, name, getter, setter, init)
        # field(s)
        decls = []
        while True:
            init = None
            if self.accept("="):
                init = self.array_init_or_expr(ty)
            decls.append((name, init))
            if not self.accept(","):
                break
            name = self.ident()
        self.expect(";")
        return ("field", mods, ty, decls)

    # ------------------------------------------------------------------ statements
    def block(self):
        self.expect("{")
        stmts = []
        while not self.at("}"):
            stmts.append(self.statement())
        self.expect("}")
        return ("block", stmts)

    def array_init_or_expr(self, ty):
        if self.at("{"):
            return ("newarr", ty, None, self.array_initializer(), None)
        return self.expr()

    def is_local_decl(self):
        save = self.p
        try:
            if self.at("const"):
                return True
            if self.peek().kind == "id" and self.peek().val == "var" and self.at("(", 1):
                return True
            t = self.try_type()
            if t is None:
                return False
            if not self.at_id():
                return False
            nxt = self.peek(1)
            return nxt.kind in ("op", "kw") and nxt.val in ("=", ";", ",", "in", ")")
        finally:
            self.p = save

    def local_decl(self):
        self.accept("const")
        if self.peek().val == "var" and self.at("(", 1):
            self.p += 1
            names = self.deconstruct_names()
            self.expect("=")
            e = self.expr()
            return ("deconstruct_decl", names, e)
        ty = self.parse_type()
        decls = []
        while True:
            nm = self.ident()
            init = self.array_init_or_expr(ty) if self.accept("=") else None
            decls.append((nm, init))
            if not self.accept(","):
                break
        return ("local", ty, decls)

    def deconstruct_names(self):
        self.expect("(")
        names = []
        while True:
            if self.at("("):
                names.append(self.deconstruct_names())
            else:
                nm = self.ident()
                names.append(None if nm == "_" else nm)
            if not self.accept(","):
                break
        self.expect(")")
        return names

    def is_local_function(self):
        save = self.p
        try:
            while self.peek().val in ("static", "async", "unsafe"):
                self.p += 1
            t = self.try_type()
            if t is None or not self.at_id() or not self.at("(", 1):
                return False
            # find the matching ")" and look for "{" or "=>"
            self.p += 1
            depth = 0
            while True:
                tk = self.peek()
                if tk.kind == "eof":
                    return False
                if tk.kind == "op" and tk.val == "(":
                    depth += 1
                elif tk.kind == "op" and tk.val == ")":
                    depth -= 1
                    if depth == 0:
                        self.p += 1
                        break
                self.p += 1
            return self.at("{") or self.at("=>")
        finally:
            self.p = save

    def statement(self):
        t = self.peek()
        if t.kind == "op":
            if t.val == "{":
                return self.block()
            if t.val == ";":
                self.p += 1
                return ("empty",)
        if t.kind == "kw":
            v = t.val
            if v == "if":
                self.p += 1
                self.expect("("); c = self.expr(); self.expect(")")
                a = self.statement()
                b = self.statement() if self.accept("else") else None
                return ("if", c, a, b)
            if v == "while":
                self.p += 1
                self.expect("("); c = self.expr(); self.expect(")")
                return ("while", c, self.statement())
            if v == "do":
                self.p += 1
                body = self.statement()
                self.expect("while"); self.expect("("); c = self.expr(); self.expect(")"); self.expect(";")
                return ("dowhile", body, c)
            if v == "for":
                self.p += 1
                self.expect("(")
                inits = []
                if not self.at(";"):
                    if self.is_local_decl():
                        inits.append(self.local_decl())
                    else:
                        inits.append(("expr", self.expr()))
                        while self.accept(","):
                            inits.append(("expr", self.expr()))
                self.expect(";")
                cond = None if self.at(";") else self.expr()
                self.expect(";")
                iters = []
                if not self.at(")"):
                    iters.append(self.expr())
                    while self.accept(","):
                        iters.append(self.expr())
                self.expect(")")
                return ("for", inits, cond, iters, self.statement())
            if v == "foreach":
                self.p += 1
                self.expect("(")
                if self.peek().val == "var" and self.at("(", 1):
                    self.p += 1
                    names = self.deconstruct_names(); ty = None
                else:
                    ty = self.parse_type()
                    names = self.ident()
                self.expect("in")
                e = self.expr()
                self.expect(")")
                return ("foreach", ty, names, e, self.statement())
            if v == "return":
                self.p += 1
                e = None if self.at(";") else self.expr()
                self.expect(";")
                return ("return", e)
            if v == "break":
                self.p += 1; self.expect(";")
                return ("break",)
            if v == "continue":
                self.p += 1; self.expect(";")
                return ("continue",)
            if v == "throw":
                self.p += 1
                e = None if self.at(";") else self.expr()
                self.expect(";")
                return ("throw", e)
            if v == "try":
                self.p += 1
                body = self.block()
                catches, fin = [], None
                while self.accept("catch"):
                    ty = nm = when = None
                    if self.accept("("):
                        ty = self.parse_type()
                        if self.at_id():
                            nm = self.ident()
                        self.expect(")")
                    if self.peek().kind == "id" and self.peek().val == "when":
                        self.p += 1
                        self.expect("("); when = self.expr(); self.expect(")")
                    catches.append((ty, nm, when, self.block()))
                if self.accept("finally"):
                    fin = self.block()
                return ("try", body, catches, fin)
            if v == "switch":
                self.p += 1
                self.expect("("); e = self.expr(); self.expect(")")
                self.expect("{")
                sections = []
                while not self.at("}"):
                    labels = []
                    while self.at("case") or self.at("default"):
                        if self.accept("default"):
                            labels.append(("default",))
                        else:
                            self.p += 1
                            labels.append(("case", self.expr()))
                        self.expect(":")
                    stmts = []
                    while not (self.at("case") or self.at("default") or self.at("}")):
                        stmts.append(self.statement())
                    sections.append((labels, stmts))
                self.expect("}")
                return ("switch", e, sections)
            if v == "using":
                self.p += 1
                self.expect("(")
                res = self.local_decl() if self.is_local_decl() else ("expr", self.expr())
                self.expect(")")
                return ("using", res, self.statement())
            if v == "lock":
                self.p += 1
                self.expect("("); self.expr(); self.expect(")")
                return self.statement()
            if v in ("checked", "unchecked") and self.at("{", 1):
                self.p += 1
                return self.block()
        if self.is_local_function():
            mods = []
            while self.peek().val in ("static", "async", "unsafe"):
                mods.append(self.peek().val); self.p += 1
            ty = self.parse_type()
            nm = self.ident()
            ps = self.params()
            if self.accept("=>"):
                body = ("exprbody", self.expr()); self.expect(";")
            else:
                body = self.block()
            return ("localfunc", ("method", mods, ty, nm, ps, body))
        if self.is_local_decl():
            d = self.local_decl()
            self.expect(";")
            return d
        e = self.expr()
        self.expect(";")
        return ("expr", e)

    # ------------------------------------------------------------------ expressions
    def expr(self):
        return self.assignment()

    def is_lambda_start(self):
        if self.at_id() and self.at("=>", 1):
            return True
        if not self.at("("):
            return False
        k, depth = 0, 0
        while True:
            t = self.peek(k)
            if t.kind == "eof":
                return False
            if t.kind == "op" and t.val == "(":
                depth += 1
            elif t.kind == "op" and t.val == ")":
                depth -= 1
                if depth == 0:
                    return self.at("=>", k + 1)
            k += 1

    def lambda_expr(self):
        params = []
        if self.at_id():
            params.append(self.ident())
        else:
            self.expect("(")
            while not self.at(")"):
                if self.peek().val in ("out", "ref", "in"):
                    self.p += 1
                if self.at_id() and (self.at(",", 1) or self.at(")", 1)):
                    params.append(self.ident())
                else:
                    self.parse_type()
                    params.append(self.ident())
                if not self.accept(","):
                    break
            self.expect(")")
        self.expect("=>")
        body = self.block() if self.at("{") else self.expr()
        return ("lambda", params, body)

    def assignment(self):
        if self.is_lambda_start():
            return self.lambda_expr()
        left = self.conditional()
        t = self.peek()
        if t.kind == "op" and t.val in ASSIGN_OPS:
            self.p += 1
            right = self.assignment()
            return ("assign", t.val, left, right)
        if t.kind == "op" and t.val == ">" and self.at(">=", 1) and self.peek(1).adj:
            self.p += 2
            return ("assign", ">>=", left, self.assignment())
        return left

    def conditional(self):
        c = self.coalesce()
        if self.accept("?"):
            a = self.assignment() if not self.at("throw") else self.throw_expr()
            self.expect(":")
            b = self.assignment() if not self.at("throw") else self.throw_expr()
            return ("cond", c, a, b)
        return c

    def throw_expr(self):
        self.expect("throw")
        return ("throwexpr", self.expr())

    def coalesce(self):
        left = self.binary(0)
        if self.accept("??"):
            right = self.throw_expr() if self.at("throw") else self.coalesce()
            return ("coalesce", left, right)
        return left

    def binary(self, level):
        if level == len(BINARY_PREC):
            return self.unary()
        ops = BINARY_PREC[level]
        left = self.binary(level + 1)
        while True:
            t = self.peek()
            op = None
            if t.kind in ("op", "kw") and t.val in ops:
                op = t.val
                # ">" ">" adjacent is a right shift (the lexer never joins them: generics close with ">>")
                if op == ">" and self.at(">", 1) and self.peek(1).adj:
                    op = None
            elif ">>" in ops and t.kind == "op" and t.val == ">" and self.at(">", 1) and self.peek(1).adj:
                op = ">>"
                self.p += 1
            if op is None:
                return left
            self.p += 1
            if op == "is":
                if self.accept("null"):
                    left = ("binary", "==", left, ("lit", None))
                    continue
                if self.peek().kind == "id" and self.peek().val == "not" and self.at("null", 1):
                    self.p += 2
                    left = ("binary", "!=", left, ("lit", None))
                    continue
                ty = self.parse_type()
                nm = None
                if self.at_id() and not self.at("=>", 1):
                    nm = self.ident()
                left = ("is", left, ty, nm)
                continue
            if op == "as":
                left = ("as", left, self.parse_type())
                continue
            right = self.binary(level + 1)
            left = ("binary", op, left, right)

    def try_cast(self):
        """ECMA-334 12.9.7: "(" type ")" is a cast when the type is a keyword / array / nullable / generic type, or
        when the token after ")" is "~", "!", "(", an identifier, a literal, or a keyword other than as / is"""
        save = self.p
        self.p += 1
        ty = self.try_type()
        if ty is None or not self.accept(")"):
            self.p = save
            return None
        t = self.peek()
        definite = ty[0] == "tupletype" or ty[1] in PREDEFINED or ty[-2] or ty[-1] or (ty[0] == "type" and ty[2])
        if ty[0] == "tupletype":
            # "(a, b) = ..." and "(a, b)" as a tuple expression are far more common than tuple casts
            self.p = save
            return None
        follows = (t.kind in ("id", "int", "real", "str", "char", "istr")
                   or (t.kind == "op" and t.val in ("~", "!", "("))
                   or (t.kind == "kw" and t.val not in ("as", "is", "in")))
        if definite and t.kind == "op" and t.val in ("-", "+", "++", "--"):
            follows = True
        if not follows:
            self.p = save
            return None
        return ("cast", ty, self.unary())

    def unary(self):
        t = self.peek()
        if t.kind == "op":
            if t.val in ("-", "+", "!", "~"):
                self.p += 1
                return ("unary", t.val, self.unary())
            if t.val in ("++", "--"):
                self.p += 1
                return ("prefix", t.val, self.unary())
            if t.val == "(" and not self.is_lambda_start():
                c = self.try_cast()
                if c is not None:
                    return c
        return self.postfix(self.primary())

    def arguments(self):
        self.expect("(")
        args = []
        while not self.at(")"):
            name = mod = None
            if self.at_id() and self.at(":", 1):
                name = self.ident(); self.p += 1
            if self.peek().val in ("out", "ref", "in") and self.peek().kind == "kw":
                mod = self.peek().val; self.p += 1
                if mod == "out":
                    # out var x / out int x / out _
                    save = self.p
                    ty = self.try_type()
                    if ty is not None and self.at_id() and (self.at(",", 1) or self.at(")", 1)):
                        args.append((name, ("outvar", ty, self.ident()), mod))
                        if not self.accept(","):
                            break
                        continue
                    self.p = save
            args.append((name, self.expr(), mod))
            if not self.accept(","):
                break
        self.expect(")")
        return args

    def try_generic_call_args(self):
        save = self.p
        try:
            self.type_args()
        except CsSyntaxError:
            self.p = save
            return False
        if self.at("("):
            return True
        self.p = save
        return False

    def postfix(self, e):
        while True:
            t = self.peek()
            if t.kind != "op":
                return e
            v = t.val
            if v == "." or v == "?.":
                self.p += 1
                name = self.ident()
                if self.at("<"):
                    self.try_generic_call_args()
                e = ("member", e, name, v == "?.")
            elif v == "(":
                e = ("call", e, self.arguments())
            elif v == "[":
                self.p += 1
                idx = [self.expr()]
                while self.accept(","):
                    idx.append(self.expr())
                self.expect("]")
                e = ("index", e, idx, False)
            elif v == "?" and self.at("[", 1):
                self.p += 2
                idx = [self.expr()]
                while self.accept(","):
                    idx.append(self.expr())
                self.expect("]")
                e = ("index", e, idx, True)
            elif v in ("++", "--"):
                self.p += 1
                e = ("postfix", v, e)
            elif v == "!" and not self.at("=", 1) and self.peek(1).kind == "op" and self.peek(1).val in (".", ")", ";", ","):
                self.p += 1  # null-forgiving
            else:
                return e

    def array_initializer(self):
        self.expect("{")
        items = []
        while not self.at("}"):
            items.append(self.array_initializer() if self.at("{") else self.expr())
            if not self.accept(","):
                break
        self.expect("}")
        return items

    def object_or_collection_init(self):
        self.expect("{")
        if self.at("}"):
            self.p += 1
            return ("collinit", [])
        if (self.at_id() and self.at("=", 1)) or self.at("["):
            items = []
            while not self.at("}"):
                if self.accept("["):
                    key = [self.expr()]
                    while self.accept(","):
                        key.append(self.expr())
                    self.expect("]"); self.expect("=")
                    items.append(("idx", key, self.init_value()))
                else:
                    nm = self.ident(); self.expect("=")
                    items.append(("prop", nm, self.init_value()))
                if not self.accept(","):
                    break
            self.expect("}")
            return ("objinit", items)
        items = []
        while not self.at("}"):
            if self.at("{"):
                items.append(("multi", self.array_initializer()))
            else:
                items.append(("one", self.expr()))
            if not self.accept(","):
                break
        self.expect("}")
        return ("collinit", items)

    def init_value(self):
        if self.at("{"):
            return ("nestedinit", self.object_or_collection_init())
        return self.expr()

    def primary(self):
        t = self.peek()
        k, v = t.kind, t.val
        if k == "int" or k == "real" or k == "str":
            self.p += 1
            return ("lit", v)
        if k == "char":
            self.p += 1
            return ("charlit", v)
        if k == "istr":
            self.p += 1
            parts = []
            for part in v:
                if isinstance(part, str):
                    parts.append(part)
                else:
                    src, align, fmt = part
                    sub = Parser(src, self.name); e = sub.expr()
                    if sub.peek().kind != "eof":
                        sub.fail("trailing tokens in interpolation hole")
                    a = None
                    if align is not None:
                        sa = Parser(align, self.name); a = sa.expr()
                    parts.append((e, a, fmt))
            return ("istr", parts)
        if k == "id":
            self.p += 1
            if v == "nameof" and self.at("("):
                self.p += 1
                depth, words = 1, []
                while depth:
                    tk = self.peek(); self.p += 1
                    if tk.kind == "op" and tk.val == "(":
                        depth += 1
                    elif tk.kind == "op" and tk.val == ")":
                        depth -= 1
                    elif tk.kind in ("id", "kw"):
                        words.append(tk.val)
                return ("lit", words[-1])
            if self.at("<") and self.try_generic_call_args():
                pass
            return ("name", v)
        if k == "kw":
            if v == "true" or v == "false":
                self.p += 1
                return ("lit", v == "true")
            if v == "null":
                self.p += 1
                return ("lit", None)
            if v == "this":
                self.p += 1
                return ("this",)
            if v == "base":
                self.p += 1
                return ("base",)
            if v in PREDEFINED:
                self.p += 1
                return ("predef", v)
            if v == "typeof":
                self.p += 1
                self.expect("("); ty = self.parse_type(); self.expect(")")
                return ("typeof", ty)
            if v == "default":
                self.p += 1
                ty = None
                if self.accept("("):
                    ty = self.parse_type(); self.expect(")")
                return ("default", ty)
            if v in ("checked", "unchecked"):
                self.p += 1
                self.expect("("); e = self.expr(); self.expect(")")
                return e
            if v == "new":
                return self.new_expr()
            if v == "throw":
                return self.throw_expr()
        if k == "op" and v == "(":
            self.p += 1
            items = []
            while True:
                nm = None
                if self.at_id() and self.at(":", 1):
                    nm = self.ident(); self.p += 1
                # "(var a, var b) = ..." / "(int a, string b) = ..." declaration expressions
                items.append((nm, self.expr()))
                if not self.accept(","):
                    break
            self.expect(")")
            if len(items) == 1 and items[0][0] is None:
                return ("paren", items[0][1])
            return ("tuple", items)
        self.fail("unexpected token in expression")

    def new_expr(self):
        self.expect("new")
        if self.at("{"):
            # anonymous object
            self.p += 1
            items = []
            while not self.at("}"):
                if self.at_id() and self.at("=", 1):
                    nm = self.ident(); self.p += 1
                    items.append((nm, self.expr()))
                else:
                    e = self.expr()
                    nm = e[1] if e[0] == "name" else e[2]
                    items.append((nm, e))
                if not self.accept(","):
                    break
            self.expect("}")
            return ("anon", items)
        if self.at("["):
            # implicitly typed array: new[] { ... }
            self.p += 1
            rank = 1
            while self.accept(","):
                rank += 1
            self.expect("]")
            return ("newarr", None, None, self.array_initializer(), rank)
        # element type: parse without consuming a dimension list "[n, m]"
        save = self.p
        ty = self.parse_type()
        if self.at("["):
            # new T[n] / new T[n, m] / new T[n][] : dimensions given
            self.p += 1
            dims = [self.expr()]
            while self.accept(","):
                dims.append(self.expr())
            self.expect("]")
            extra = []
            while self.at("[") and (self.at("]", 1) or self.at(",", 1)):
                self.p += 1
                r = 1
                while self.accept(","):
                    r += 1
                self.expect("]")
                extra.append(r)
            elem = ty
            if extra:
                elem = ty[:-2] + (list(ty[-2]) + extra, ty[-1])
            init = self.array_initializer() if self.at("{") else None
            return ("newarr", elem, dims, init, len(dims))
        if ty[-2]:
            # new T[] { ... } / new T[,] { {..}, {..} }
            ranks = list(ty[-2])
            rank = ranks[0]
            elem = ty[:-2] + (ranks[1:], ty[-1])
            return ("newarr", elem, None, self.array_initializer(), rank)
        args = self.arguments() if self.at("(") else []
        init = self.object_or_collection_init() if self.at("{") else None
        return ("new", ty, args, init)


def parse_source(src, name="<cs>"):
    return Parser(src, name).parse_unit()
