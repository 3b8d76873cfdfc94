#!/usr/bin/env python3
"""Mechanical translator of the reference's Fortran-90 subset to C++  --  TEST INFRASTRUCTURE ONLY.

Purpose: there is no Fortran compiler in this image, so the reference cannot be run as it is.  This tool reads
core/module_noahmp_*.f90 where they lie under /root/reference, translates every subroutine of `noahmp_func`
statement by statement (no knowledge of the physics, no per-routine special cases) and writes ONE C++ file into
oracle/_ref/ (git-ignored: a derived copy of reference source never enters the repository).  Compiled with
`g++ -O1 -ffp-contract=off` it executes the reference's own statements in the reference's own order with the arithmetic
a Fortran compiler without fast-math would use:

  * `real` -> float, `real(r8)`/`double precision` -> double, default integer -> int, logical -> bool
  * mixed-mode expressions: the usual arithmetic conversions of C++ coincide with Fortran's (int -> real -> double);
    integer division truncates in both; assignment converts like INT()/REAL()
  * `x ** n` with integer n: repeated squaring as in libgcc's __powisf2 (what gfortran emits); real exponents: powf/pow
  * intrinsics call glibc's float functions (expf, logf, ...), as gfortran's runtime does
  * arguments by reference (scalars `T&` unless intent(in), arrays as pointer + the dummy's declared lower bound)
  * internal procedures (`contains` inside a subroutine) become lambdas capturing the host's variables
  * WRITE/PRINT are dropped; `call wrf_error_fatal`, `stop` raise a C++ exception that the C entry returns as an error

The golden vectors made with it (tools/make_reference_golden.py -> tests/golden/ref_*.npz) pin the hand-written oracle
to the reference's source text.  Nothing here is imported by the product.
"""
from __future__ import annotations

import json
import os
import re
import sys

CPP_KEYWORDS = {"int", "float", "double", "and", "or", "not", "xor", "new", "delete", "this", "class", "template", "union",
                "register", "auto", "switch", "case", "default", "char", "short", "long", "signed", "unsigned", "void",
                "const", "static", "extern", "struct", "enum", "typedef", "namespace", "using", "operator", "private",
                "public", "protected", "virtual", "friend", "inline", "bool", "true", "false", "try", "catch", "throw",
                "goto", "break", "continue", "return", "if", "else", "for", "while", "do", "sizeof", "volatile", "errno",
                "stdin", "stdout", "stderr", "NULL", "EOF", "main", "max", "min", "abs", "time", "index", "y0", "y1", "yn",
                "j0", "j1", "jn", "gamma", "exp", "log", "sqrt", "pow", "sin", "cos", "tan", "div", "free", "signal", "rand"}

SKIP_SUBS = {"bvocflux"}           # dead code in the reference (its only call is commented out, func:6749)
FATAL_CALLS = {"wrf_error_fatal", "wrf_error_fatal3", "stop_program", "wrf_abort"}
MESSAGE_CALLS = {"wrf_message", "wrf_debug"}


REAL = "float"    # C++ type of Fortran `real`; translate(..., real_type="double") builds the fp64 twin (SURVEY 8c iv)


def cname(n: str) -> str:
    return n + "_" if n in CPP_KEYWORDS else n


# ------------------------------------------------------------------------------------------------ source
def strip_comment(line: str) -> str:
    q = None
    for i, ch in enumerate(line):
        if q:
            if ch == q:
                q = None
        elif ch in "'\"":
            q = ch
        elif ch == "!":
            return line[:i]
    return line


def logical_lines(path):
    """(first line number, statement) with comments removed, continuations joined, lower-cased."""
    out, cur, start = [], "", 0
    for no, raw in enumerate(open(path, errors="replace"), 1):
        s = strip_comment(raw.rstrip("\n")).strip()
        if not s:
            continue
        if cur:
            if s.startswith("&"):
                s = s[1:].lstrip()
        else:
            start = no
        if s.endswith("&"):
            cur += s[:-1].rstrip() + " "
            continue
        cur += s
        for part in split_semicolons(cur):
            if part.strip():
                out.append((start, lower_outside_strings(part.strip())))
        cur = ""
    return out


def split_semicolons(s):
    parts, q, cur = [], None, ""
    for ch in s:
        if q:
            cur += ch
            if ch == q:
                q = None
        elif ch in "'\"":
            q = ch
            cur += ch
        elif ch == ";":
            parts.append(cur)
            cur = ""
        else:
            cur += ch
    parts.append(cur)
    return parts


def lower_outside_strings(s):
    out, q = [], None
    for ch in s:
        if q:
            out.append(ch)
            if ch == q:
                q = None
        elif ch in "'\"":
            q = ch
            out.append(ch)
        else:
            out.append(ch.lower())
    return "".join(out)


# ------------------------------------------------------------------------------------------------ model
class Var:
    def __init__(self, name, ftype, dims=None, intent=None, param=False, init=None, line=0):
        self.name, self.ftype, self.dims, self.intent, self.param, self.init, self.line = name, ftype, dims, intent, param, init, line
        self.is_arg = False

    @property
    def ctype(self):
        return {"real": REAL, "double": "double", "integer": "int", "logical": "bool", "character": "const char*"}[self.ftype]

    @property
    def rank(self):
        return len(self.dims) if self.dims else 0


class Sub:
    def __init__(self, name, args, line):
        self.name, self.args, self.line = name, args, line
        self.vars: dict[str, Var] = {}
        self.order: list[str] = []
        self.body: list[tuple[int, str]] = []
        self.internal: list["Sub"] = []
        self.host: "Sub | None" = None
        self.module = None


class Module:
    def __init__(self, name):
        self.name = name
        self.vars: dict[str, Var] = {}
        self.order: list[str] = []
        self.subs: list[Sub] = []


TYPE_RE = re.compile(r"^(real|integer|logical|double\s+precision|character)\s*(\([^)]*\))?\s*(.*)$")


def split_top(s, sep=","):
    """Split at top-level separators (outside parentheses and (/ /) constructors and strings)."""
    parts, depth, cur, q = [], 0, "", None
    i = 0
    while i < len(s):
        ch = s[i]
        if q:
            cur += ch
            if ch == q:
                q = None
        elif ch in "'\"":
            q = ch
            cur += ch
        elif ch == "(":
            depth += 1
            cur += ch
        elif ch == ")":
            depth -= 1
            cur += ch
        elif ch == sep and depth == 0:
            parts.append(cur.strip())
            cur = ""
        else:
            cur += ch
        i += 1
    if cur.strip() or parts:
        parts.append(cur.strip())
    return parts


def parse_decl(stmt, line):
    """Type declaration statement -> list of Var, or None."""
    m = TYPE_RE.match(stmt)
    if not m:
        return None
    base, kind, rest = m.group(1), m.group(2), m.group(3)
    after = stmt[len(base):len(base) + 1]
    if after not in (" ", "(", ",", ":"):
        return None     # an identifier that merely starts with a type keyword
    if "::" not in rest and Translator.find_assign(stmt) is not None:
        return None     # an assignment, not an old-style declaration
    ftype = "double" if base.startswith("double") else base
    if ftype == "real" and kind:
        k = kind.strip("() ").replace("kind=", "").strip()
        if k in ("r8", "8", "c_double"):
            ftype = "double"
    if "::" in rest:
        attrs, ents = rest.split("::", 1)
    else:
        attrs, ents = "", rest
    dims = intent = None
    param = False
    for a in split_top(attrs.strip().lstrip(",")):
        a = a.strip()
        if not a:
            continue
        if a.startswith("dimension"):
            dims = split_top(a[a.index("(") + 1:a.rindex(")")])
        elif a.startswith("intent"):
            intent = a[a.index("(") + 1:a.rindex(")")].strip().replace(" ", "")
        elif a == "parameter":
            param = True
        elif a in ("save", "allocatable", "target", "pointer", "public", "private", "optional"):
            if a in ("allocatable", "pointer", "optional"):
                raise NotImplementedError(f"line {line}: attribute {a}")
        else:
            raise NotImplementedError(f"line {line}: attribute {a!r}")
    out = []
    for e in split_top(ents):
        e = e.strip()
        if not e:
            continue
        init = None
        # entity: name[(dims)] [= init]
        m2 = re.match(r"^([a-z_][a-z0-9_]*)\s*(\(.*?\))?\s*(=\s*(.*))?$", e)
        if not m2:
            raise NotImplementedError(f"line {line}: entity {e!r}")
        # the lazy (.*?) may stop early for nested parens: redo with explicit scanning
        name = m2.group(1)
        restent = e[len(name):].strip()
        edims = dims
        if restent.startswith("("):
            depth, j = 0, 0
            for j, ch in enumerate(restent):
                depth += ch == "("
                depth -= ch == ")"
                if depth == 0:
                    break
            edims = split_top(restent[1:j])
            restent = restent[j + 1:].strip()
        if restent.startswith("="):
            init = restent[1:].strip()
        out.append(Var(name, ftype, edims, intent, param, init, line))
    return out


def parse_file(path, want_subs=True):
    stmts = logical_lines(path)
    mod = None
    stack: list[Sub] = []
    i = 0
    in_interface = False
    while i < len(stmts):
        line, s = stmts[i]
        i += 1
        if in_interface:
            if re.match(r"^end\s*interface", s):
                in_interface = False
            continue
        if re.match(r"^interface\b", s):
            in_interface = True
            continue
        m = re.match(r"^module\s+([a-z_][a-z0-9_]*)$", s)
        if m and not stack:
            mod = Module(m.group(1))
            continue
        if re.match(r"^end\s*module", s):
            continue
        m = re.match(r"^(?:recursive\s+)?subroutine\s+([a-z_][a-z0-9_]*)\s*(\((.*)\))?$", s)
        if m:
            args = [a.strip() for a in split_top(m.group(3) or "") if a.strip()]
            sub = Sub(m.group(1), args, line)
            sub.module = mod
            if stack:
                sub.host = stack[-1]
                stack[-1].internal.append(sub)
            else:
                mod.subs.append(sub)
            stack.append(sub)
            continue
        if re.match(r"^end\s*subroutine", s):
            stack.pop()
            continue
        if re.match(r"^(end\s*)?function\b", s) or re.match(r"^(real|integer|logical)[^:]*function\b", s):
            raise NotImplementedError(f"{path}:{line}: functions are not in the supported subset")
        if s == "contains":
            continue
        if re.match(r"^(implicit\s+none|private\b|public\b|save$|use\s)", s):
            continue
        scope = stack[-1] if stack else mod
        m = re.match(r"^parameter\s*\((.*)\)$", s)
        if m:
            for item in split_top(m.group(1)):
                n, e = item.split("=", 1)
                scope.vars[n.strip()].param = True
                scope.vars[n.strip()].init = e.strip()
            continue
        try:
            decl = parse_decl(s, line)
        except NotImplementedError:
            if stack and stack[-1].name in SKIP_SUBS:
                continue
            raise
        is_exec_started = bool(stack) and bool(stack[-1].body)
        if decl is not None:
            for v in decl:
                if isinstance(scope, Sub) and v.name in scope.args:
                    v.is_arg = True
                scope.vars[v.name] = v
                scope.order.append(v.name)
            continue
        if stack:
            stack[-1].body.append((line, s))
        # module-level executable statements do not exist
    if not want_subs:
        mod.subs = []
    return mod


# ------------------------------------------------------------------------------------------------ expressions
TOKEN_RE = re.compile(r"""
    (?P<dotop>\.(?:and|or|not|eqv|neqv|eq|ne|lt|le|gt|ge|true|false)\.)
  | (?P<num>(?:\d+\.\d*|\.\d+|\d+)(?:[ed][+-]?\d+)?(?:_[a-z0-9]+)?)
  | (?P<name>[a-z_][a-z0-9_]*)
  | (?P<str>'(?:[^']|'')*'|"(?:[^"]|"")*")
  | (?P<op>\*\*|==|/=|<=|>=|\(/|/\)|[-+*/()<>,:=%])
  | (?P<ws>\s+)
""", re.X)

INTRINSIC_1 = {"sqrt": "f_sqrt", "exp": "f_exp", "log": "f_log", "alog": "f_log", "log10": "f_log10", "alog10": "f_log10",
               "tanh": "f_tanh", "atan": "f_atan", "cos": "f_cos", "sin": "f_sin", "tan": "f_tan", "acos": "f_acos",
               "asin": "f_asin", "abs": "f_abs", "nint": "f_nint", "float": "F_REAL", "dble": "(double)", "sngl": "F_REAL",
               "aint": "f_aint", "anint": "f_anint", "floor": "f_floor", "ceiling": "f_ceiling", "isnan": "f_isnan"}


def tokenize(s):
    toks, pos = [], 0
    while pos < len(s):
        m = TOKEN_RE.match(s, pos)
        if not m:
            raise SyntaxError(f"cannot tokenize {s[pos:pos + 20]!r} in {s!r}")
        kind = m.lastgroup
        text = m.group(kind)
        if kind == "num":
            # "1.eq.2": the dot belongs to the operator
            m2 = re.match(r"^(\d+)\.(?=(?:and|or|not|eqv|neqv|eq|ne|lt|le|gt|ge)\.)", s[pos:])
            if m2:
                text = m2.group(1)
                toks.append(("num", text))
                pos += len(text)
                continue
        if kind != "ws":
            toks.append((kind, text))
        pos = m.end()
    return toks


class Section:
    """a(lo:hi) as an actual argument or in an array assignment."""

    def __init__(self, name, lo, hi):
        self.name, self.lo, self.hi = name, lo, hi


class ExprTranslator:
    def __init__(self, tr, sub):
        self.tr, self.sub = tr, sub
        self.elem_index = None          # name of the elemental loop index inside an array assignment

    def lookup(self, name):
        return self.tr.lookup(self.sub, name)

    def parse(self, s):
        self.toks = tokenize(s)
        self.i = 0
        e = self.p_or()
        if self.i != len(self.toks):
            raise SyntaxError(f"trailing tokens in {s!r}: {self.toks[self.i:]}")
        return e

    def peek(self):
        return self.toks[self.i] if self.i < len(self.toks) else (None, None)

    def take(self, text=None):
        k, t = self.peek()
        if text is not None and t != text:
            raise SyntaxError(f"expected {text!r}, got {t!r} in {self.toks}")
        self.i += 1
        return k, t

    def p_or(self):
        e = self.p_and()
        while self.peek()[1] in (".or.",):
            self.take()
            e = f"({e} || {self.p_and()})"
        while self.peek()[1] in (".eqv.", ".neqv."):
            op = self.take()[1]
            e = f"({e} {'==' if op == '.eqv.' else '!='} {self.p_and()})"
        return e

    def p_and(self):
        e = self.p_not()
        while self.peek()[1] == ".and.":
            self.take()
            e = f"({e} && {self.p_not()})"
        return e

    def p_not(self):
        if self.peek()[1] == ".not.":
            self.take()
            return f"(!{self.p_not()})"
        return self.p_rel()

    REL = {"==": "==", "/=": "!=", "<": "<", "<=": "<=", ">": ">", ">=": ">=", ".eq.": "==", ".ne.": "!=", ".lt.": "<",
           ".le.": "<=", ".gt.": ">", ".ge.": ">="}

    def p_rel(self):
        e = self.p_add()
        if self.peek()[1] in self.REL:
            op = self.REL[self.take()[1]]
            e = f"({e} {op} {self.p_add()})"
        return e

    def p_add(self):
        if self.peek()[1] in ("+", "-"):
            op = self.take()[1]
            t = self.p_mul()
            e = f"(-{t})" if op == "-" else t
        else:
            e = self.p_mul()
        while self.peek()[1] in ("+", "-"):
            op = self.take()[1]
            e = f"({e} {op} {self.p_mul()})"
        return e

    def p_mul(self):
        e = self.p_pow()
        while self.peek()[1] in ("*", "/"):
            op = self.take()[1]
            e = f"({e} {op} {self.p_pow()})"
        return e

    def p_pow(self):
        base = self.p_primary()
        if self.peek()[1] == "**":
            self.take()
            if self.peek()[1] in ("+", "-"):
                sgn = self.take()[1]
                ex = self.p_pow()
                ex = f"(-{ex})" if sgn == "-" else ex
            else:
                ex = self.p_pow()          # right associative
            return f"f_pow({base}, {ex})"
        return base

    def number(self, t):
        m = re.match(r"^((?:\d+\.\d*|\.\d+|\d+))(?:([ed])([+-]?\d+))?(?:_([a-z0-9]+))?$", t)
        mant, ech, ex, kind = m.groups()
        is_real = "." in mant or ech is not None
        if not is_real:
            return mant.lstrip("0") or "0"
        if "." not in mant:
            mant += ".0"
        if mant.endswith("."):
            mant += "0"
        if mant.startswith("."):
            mant = "0" + mant
        txt = mant + (f"e{ex}" if ex is not None else "")
        dbl = ech == "d" or kind in ("r8", "8")
        if dbl:
            return txt
        return txt + "f" if REAL == "float" else f"((double){txt}f)"

    def p_args(self):
        """Parse '(' a, b, lo:hi ')' -> list of str | Section-bounds tuples."""
        self.take("(")
        args = []
        if self.peek()[1] == ")":
            self.take()
            return args
        while True:
            lo = hi = None
            colon = False
            if self.peek()[0] == "name" and self.i + 1 < len(self.toks) and self.toks[self.i + 1][1] == "=":
                self.i += 2            # keyword argument (kind=...): positional from here on in this code base
            if self.peek()[1] != ":":
                lo = self.p_or()
            if self.peek()[1] == ":":
                self.take()
                colon = True
                if self.peek()[1] not in (",", ")"):
                    hi = self.p_or()
            args.append(("sec", lo, hi) if colon else lo)
            if self.peek()[1] == ",":
                self.take()
                continue
            self.take(")")
            return args

    def p_primary(self):
        k, t = self.peek()
        if k == "num":
            self.take()
            return self.number(t)
        if k == "dotop" and t in (".true.", ".false."):
            self.take()
            return "true" if t == ".true." else "false"
        if k == "str":
            self.take()
            return '"' + t[1:-1].replace("\\", "\\\\").replace('"', '\\"') + '"'
        if t == "(":
            self.take()
            e = self.p_or()
            self.take(")")
            return f"({e})"
        if t == "(/":
            self.take()
            items = []
            while self.peek()[1] != "/)":
                items.append(self.p_or())
                if self.peek()[1] == ",":
                    self.take()
            self.take("/)")
            return "{" + ", ".join(items) + "}"
        if k == "name":
            self.take()
            if self.peek()[1] == "(":
                v = self.lookup(t)
                if v is not None and v.rank > 0:
                    args = self.p_args()
                    if any(isinstance(a, tuple) for a in args):
                        if len(args) != 1:
                            raise NotImplementedError(f"multi-dimensional section of {t}")
                        _, lo, hi = args[0]
                        if self.elem_index is not None:
                            lo_e = lo if lo is not None else f"{cname(t)}.lo"
                            return f"{cname(t)}(({lo_e}) + {self.elem_index})"
                        return Section(t, lo, hi)
                    return f"{cname(t)}({', '.join(args)})"
                if v is not None:
                    sf = self.tr.stmt_func(self.sub, t)
                    if sf is None:
                        raise SyntaxError(f"{t} is a scalar but is subscripted")
                    args = self.p_args()
                    return f"{cname(t)}__sf({', '.join(args)})"
                return self.intrinsic(t, self.p_args())
            v = self.lookup(t)
            if v is None:
                raise NameError(f"unknown name {t!r} in {self.sub.name if self.sub else '<module>'}")
            if v.rank > 0:
                if self.elem_index is not None:
                    return f"{cname(t)}({cname(t)}.lo + {self.elem_index})"
                return Section(t, None, None)
            return cname(t)
        raise SyntaxError(f"unexpected token {t!r} in {self.toks}")

    def intrinsic(self, name, args):
        if any(isinstance(a, (tuple, Section)) for a in args):
            raise NotImplementedError(f"array argument to intrinsic {name}")
        if name in ("max", "min", "amax1", "amin1", "max0", "min0"):
            f = "f_max" if "max" in name else "f_min"
            e = args[0]
            for a in args[1:]:
                e = f"{f}({e}, {a})"
            return e
        if name in INTRINSIC_1 and len(args) == 1:
            return f"{INTRINSIC_1[name]}({args[0]})"
        if name == "int" and len(args) >= 1:
            return f"(int)({args[0]})"
        if name == "real":
            if len(args) == 2 and args[1].strip("()") in ("r8", "8"):
                return f"(double)({args[0]})"
            return f"F_REAL({args[0]})"
        if name == "mod" and len(args) == 2:
            return f"f_mod({args[0]}, {args[1]})"
        if name == "sign" and len(args) == 2:
            return f"f_sign({args[0]}, {args[1]})"
        if name == "atan2" and len(args) == 2:
            return f"f_atan2({args[0]}, {args[1]})"
        if name == "transfer":
            return "f_nan()"
        if name == "trim" or name == "adjustl":
            return args[0]
        raise NameError(f"unknown function {name!r}")


# ------------------------------------------------------------------------------------------------ translator
class Translator:
    def __init__(self, modules: dict[str, Module]):
        self.modules = modules
        self.subs: dict[str, Sub] = {}
        for m in modules.values():
            for s in m.subs:
                self.subs[s.name] = s
        self.out: list[str] = []
        self.warnings: list[str] = []

    # -- symbols
    def lookup(self, sub, name):
        s = sub
        while s is not None:
            if name in s.vars:
                return s.vars[name]
            s = s.host
        for m in self.modules.values():
            if name in m.vars:
                return m.vars[name]
        return None

    def find_sub(self, sub, name):
        s = sub
        while s is not None:
            for c in s.internal:
                if c.name == name:
                    return c
            s = s.host
        return self.subs.get(name)

    def stmt_func(self, sub, name):
        s = sub
        while s is not None:
            if name in getattr(s, "stmt_funcs", {}):
                return s.stmt_funcs[name]
            s = s.host
        return None

    # -- emit helpers
    def emit(self, s=""):
        self.out.append(s)

    def expr(self, sub, text, elem=None):
        et = ExprTranslator(self, sub)
        et.elem_index = elem
        return et.parse(text)

    def dims_of(self, sub, v):
        """[(lo_expr, hi_expr)] as C++."""
        res = []
        for d in v.dims:
            if ":" in d:
                lo, hi = d.split(":", 1)
                lo, hi = lo.strip() or "1", hi.strip()
            else:
                lo, hi = "1", d.strip()
            res.append((self.expr(sub, lo), self.expr(sub, hi)))
        return res

    # -- declarations
    def decl_local(self, sub, v, indent="  "):
        ct = v.ctype
        n = cname(v.name)
        if v.ftype == "character":
            return f'{indent}const char* {n} = ""; (void){n};'
        if v.rank == 0:
            init = ""
            if v.init is not None:
                init = f" = {self.expr(sub, v.init)}"
            elif not v.param:
                init = "{}"
            const = "const " if v.param else ""
            return f"{indent}{const}{ct} {n}{init}; (void){n};"
        d = self.dims_of(sub, v)
        if v.rank == 1:
            if v.init is not None:
                e = self.expr(sub, v.init)
                if isinstance(e, str) and e.startswith("{"):
                    return f"{indent}FArr<{ct}> {n}({d[0][0]}, {d[0][1]}, std::initializer_list<{ct}>{e});"
                return f"{indent}FArr<{ct}> {n}({d[0][0]}, {d[0][1]}, ({ct})({e}));"
            return f"{indent}FArr<{ct}> {n}({d[0][0]}, {d[0][1]});"
        if v.rank == 2:
            if v.init is not None:
                return f"{indent}FArr2<{ct}> {n}({d[0][0]}, {d[0][1]}, {d[1][0]}, {d[1][1]}, ({ct})({self.expr(sub, v.init)}));"
            return f"{indent}FArr2<{ct}> {n}({d[0][0]}, {d[0][1]}, {d[1][0]}, {d[1][1]});"
        raise NotImplementedError(f"rank {v.rank} array {v.name}")

    def signature(self, sub, as_lambda=False):
        ps = []
        for a in sub.args:
            v = sub.vars.get(a)
            if v is None:
                raise NameError(f"{sub.name}: dummy argument {a} is not declared")
            if v.rank > 0:
                ps.append(f"{v.ctype}* {cname(a)}__p")
            elif v.ftype == "character":
                ps.append(f"const char* {cname(a)}")
            elif v.intent == "in":
                ps.append(f"{v.ctype} {cname(a)}")
            else:
                ps.append(f"{v.ctype}& {cname(a)}")
        return ", ".join(ps)

    # -- statements
    def actual_args(self, sub, callee, argtexts, line):
        outs = []
        if len(argtexts) != len(callee.args):
            raise SyntaxError(f"line {line}: call {callee.name} with {len(argtexts)} args, expected {len(callee.args)}")
        for text, dn in zip(argtexts, callee.args):
            dv = callee.vars[dn]
            e = self.expr(sub, text)
            if dv.rank > 0:
                if isinstance(e, Section):
                    if e.lo is None:
                        outs.append(f"{cname(e.name)}.p")
                    else:
                        outs.append(f"&{cname(e.name)}({e.lo})")
                else:
                    # an element as the start of a sequence-associated array
                    m = re.match(r"^([a-z_][a-z0-9_]*)\((.*)\)$", e)
                    if not m:
                        raise SyntaxError(f"line {line}: array dummy {dn} of {callee.name} gets {e!r}")
                    outs.append(f"&{e}")
            else:
                if isinstance(e, Section):
                    raise SyntaxError(f"line {line}: scalar dummy {dn} of {callee.name} gets an array")
                outs.append(e)
        return outs

    def assignment(self, sub, lhs, rhs, line, ind):
        m = re.match(r"^([a-z_][a-z0-9_]*)\s*\(([a-z0-9_, ]*)\)$", lhs)
        if m:
            v = self.lookup(sub, m.group(1))
            if v is not None and v.rank == 0:
                # statement function: name(dummies) = expression
                params = [a.strip() for a in m.group(2).split(",") if a.strip()]
                if not hasattr(sub, "stmt_funcs"):
                    sub.stmt_funcs = {}
                sub.stmt_funcs[m.group(1)] = params
                ps = ", ".join(f"{self.lookup(sub, a).ctype} {cname(a)}" for a in params)
                self.emit(f"{ind}auto {cname(m.group(1))}__sf = [&]({ps}) -> {v.ctype} {{ return {self.expr(sub, rhs)}; }};")
                return
        lt = ExprTranslator(self, sub)
        l = lt.parse(lhs)
        if isinstance(l, Section):
            v = self.lookup(sub, l.name)
            if v.rank != 1:
                raise NotImplementedError(f"line {line}: whole-array assignment to rank-{v.rank} {l.name}")
            n = cname(l.name)
            lo = l.lo if l.lo is not None else f"{n}.lo"
            hi = l.hi if l.hi is not None else f"{n}.hi()"
            r = self.expr(sub, rhs, elem="k__")
            if isinstance(r, Section):
                raise NotImplementedError(f"line {line}: {rhs}")
            self.emit(f"{ind}{{ const int lo__ = {lo}, n__ = ({hi}) - lo__ + 1; std::vector<decltype({n}(lo__) + 0)> t__(n__ > 0 ? n__ : 0);")
            self.emit(f"{ind}  for (int k__ = 0; k__ < n__; ++k__) t__[k__] = {r};")
            self.emit(f"{ind}  for (int k__ = 0; k__ < n__; ++k__) {n}(lo__ + k__) = t__[k__]; }}")
            return
        r = self.expr(sub, rhs)
        if isinstance(r, Section):
            raise NotImplementedError(f"line {line}: array value assigned to scalar: {lhs} = {rhs}")
        self.emit(f"{ind}{l} = {r};")

    def translate_body(self, sub, ind="  "):
        loops = []   # stack of (kind, name)
        depth = 0
        for line, s in sub.body:
            pad = ind + "  " * depth
            try:
                depth = self.statement(sub, line, s, pad, loops, depth)
            except Exception as exc:
                raise type(exc)(f"{sub.name} (func line {line}): {s!r}: {exc}") from exc

    def statement(self, sub, line, s, pad, loops, depth):
        # drop I/O
        if re.match(r"^(write\s*\(|print\b|format\s*\(|\d+\s+format\s*\(|open\s*\(|close\s*\(|read\s*\()", s):
            return depth
        if s in ("continue",):
            return depth
        m = re.match(r"^if\s*\((.*)\)\s*then$", s)
        if m:
            self.emit(f"{pad}if ({self.expr(sub, m.group(1))}) {{")
            return depth + 1
        m = re.match(r"^else\s*if\s*\((.*)\)\s*then$", s)
        if m:
            self.emit(f"{pad[:-2]}}} else if ({self.expr(sub, m.group(1))}) {{")
            return depth
        if s == "else":
            self.emit(f"{pad[:-2]}}} else {{")
            return depth
        if re.match(r"^end\s*if$", s):
            self.emit(f"{pad[:-2]}}}")
            return depth - 1
        m = re.match(r"^(?:([a-z_][a-z0-9_]*)\s*:\s*)?do\s+while\s*\((.*)\)$", s)
        if m:
            loops.append(m.group(1))
            self.emit(f"{pad}while ({self.expr(sub, m.group(2))}) {{")
            return depth + 1
        m = re.match(r"^(?:([a-z_][a-z0-9_]*)\s*:\s*)?do\s+([a-z_][a-z0-9_]*)\s*=\s*(.*)$", s)
        if m:
            loops.append(m.group(1))
            var = cname(m.group(2))
            parts = split_top(m.group(3))
            a, b = self.expr(sub, parts[0]), self.expr(sub, parts[1])
            c = self.expr(sub, parts[2]) if len(parts) > 2 else "1"
            u = f"{len(self.out)}"
            self.emit(f"{pad}{{ const int b__{u} = {b}, c__{u} = {c};")
            self.emit(f"{pad}for ({var} = {a}; (c__{u} > 0) ? ({var} <= b__{u}) : ({var} >= b__{u}); {var} += c__{u}) {{")
            return depth + 1
        m = re.match(r"^(?:([a-z_][a-z0-9_]*)\s*:\s*)?do$", s)
        if m:
            loops.append(m.group(1))
            self.emit(f"{pad}for (;;) {{")
            return depth + 1
        m = re.match(r"^end\s*do(\s+[a-z_][a-z0-9_]*)?$", s)
        if m:
            name = loops.pop()
            # was it a counted loop? (opened with an extra brace)
            self.emit(f"{pad[:-2]}}}")
            # find the matching opener to see whether an extra scope brace is needed
            self.close_counted(pad[:-2])
            if name:
                self.emit(f"{pad[:-2]}exit_{name}__:;")
            return depth - 1
        m = re.match(r"^exit(\s+([a-z_][a-z0-9_]*))?$", s)
        if m:
            if m.group(2):
                self.emit(f"{pad}goto exit_{m.group(2)}__;")
            else:
                self.emit(f"{pad}break;")
            return depth
        if s == "cycle":
            self.emit(f"{pad}continue;")
            return depth
        if s == "return":
            self.emit(f"{pad}return;")
            return depth
        if re.match(r"^stop\b", s):
            self.emit(f'{pad}f_fatal("stop at func line {line}");')
            return depth
        m = re.match(r"^call\s+([a-z_][a-z0-9_]*)\s*(\((.*)\))?$", s)
        if m:
            name = m.group(1)
            if name in FATAL_CALLS:
                self.emit(f'{pad}f_fatal("{name} at func line {line}");')
                return depth
            if name in MESSAGE_CALLS:
                return depth
            callee = self.find_sub(sub, name)
            if callee is None:
                raise NameError(f"call to unknown subroutine {name}")
            args = self.actual_args(sub, callee, [a for a in split_top(m.group(3) or "") if a != ""], line)
            self.emit(f"{pad}{cname(name)}({', '.join(args)});")
            return depth
        m = re.match(r"^if\s*\(", s)
        if m:
            # one-line logical IF: find the matching parenthesis
            d, j = 0, 0
            for j in range(s.index("("), len(s)):
                d += s[j] == "("
                d -= s[j] == ")"
                if d == 0:
                    break
            cond, rest = s[s.index("(") + 1:j], s[j + 1:].strip()
            self.emit(f"{pad}if ({self.expr(sub, cond)}) {{")
            self.statement(sub, line, rest, pad + "  ", loops, depth)
            self.emit(f"{pad}}}")
            return depth
        # assignment
        eq = self.find_assign(s)
        if eq is not None:
            self.assignment(sub, s[:eq].strip(), s[eq + 1:].strip(), line, pad)
            return depth
        raise NotImplementedError(f"statement not in the supported subset: {s!r}")

    def close_counted(self, pad):
        """After a closing brace of a loop: if the loop header opened an extra `{ const int b__...` scope, close it."""
        # walk back to the matching opener
        depth = 0
        for k in range(len(self.out) - 1, -1, -1):
            t = self.out[k]
            depth += t.count("}") - t.count("{")
            if depth == 0:
                # line k opened the loop; the extra scope line is k-1
                if k > 0 and re.search(r"\{ const int b__\d+ = ", self.out[k - 1]) and self.out[k].lstrip().startswith("for ("):
                    self.emit(f"{pad}}}")
                return

    @staticmethod
    def find_assign(s):
        depth, q = 0, None
        for i, ch in enumerate(s):
            if q:
                if ch == q:
                    q = None
            elif ch in "'\"":
                q = ch
            elif ch == "(":
                depth += 1
            elif ch == ")":
                depth -= 1
            elif ch == "=" and depth == 0:
                if s[i + 1:i + 2] == "=" or s[i - 1] in "=/<>":
                    continue
                return i
        return None

    # -- a subroutine
    def emit_sub(self, sub, ind="", as_lambda=False):
        if sub.name in SKIP_SUBS:
            return
        sig = self.signature(sub)
        if as_lambda:
            self.emit(f"{ind}auto {cname(sub.name)} = [&]({sig}) -> void {{")
        else:
            self.emit(f"{ind}void {cname(sub.name)}({sig}) {{")
        # dummy arrays
        for a in sub.args:
            v = sub.vars[a]
            if v.rank == 1:
                d = self.dims_of(sub, v)
                self.emit(f"{ind}  FView<{v.ctype}> {cname(a)}({cname(a)}__p, {d[0][0]}, {d[0][1]});")
            elif v.rank == 2:
                d = self.dims_of(sub, v)
                self.emit(f"{ind}  FView2<{v.ctype}> {cname(a)}({cname(a)}__p, {d[0][0]}, {d[0][1]}, {d[1][0]}, {d[1][1]});")
        for n in sub.order:
            v = sub.vars[n]
            if v.is_arg:
                continue
            self.emit(self.decl_local(sub, v, ind + "  "))
        for c in sub.internal:
            self.emit_sub(c, ind + "  ", as_lambda=True)
        self.translate_body(sub, ind + "  ")
        self.emit(f"{ind}}}{';' if as_lambda else ''}")
        self.emit()

    # -- whole program
    def emit_module_vars(self, m):
        for n in m.order:
            v = m.vars[n]
            if v.ftype == "character":
                continue
            if v.param and v.ftype == "integer" and v.init and re.match(r"^c_[a-z0-9_]+$", v.init):
                continue            # kind constants (r4 = C_FLOAT ...)
            self.emit(self.decl_local(None, v, "").replace("(void)" + cname(n) + ";", ""))

    def registry(self):
        """C++ table of the settable module variables + JSON description for the Python side."""
        desc = {}
        self.emit("struct RegEntry { const char* name; void* ptr; int type; int count; };")
        self.emit("static RegEntry registry[] = {")
        for m in self.modules.values():
            for n in m.order:
                v = m.vars[n]
                if v.param or v.ftype not in ("real", "integer", "double"):
                    continue
                tcode = {"real": 0, "integer": 1, "double": 2}[v.ftype]
                if v.rank == 0:
                    self.emit(f'  {{"{n}", &{cname(n)}, {tcode}, 1}},')
                    desc[n] = {"module": m.name, "type": v.ftype, "dims": []}
                else:
                    self.emit(f'  {{"{n}", {cname(n)}.p, {tcode}, {cname(n)}.size()}},')
                    desc[n] = {"module": m.name, "type": v.ftype, "dims": v.dims}
        self.emit("  {nullptr, nullptr, 0, 0}};")
        return desc

    def c_entry(self, sub):
        """extern "C" int ref_<sub>(void** a): every argument by address."""
        self.emit(f'extern "C" int ref_{sub.name}(void** a) {{')
        self.emit("  try {")
        args = []
        for k, an in enumerate(sub.args):
            v = sub.vars[an]
            if v.rank > 0:
                args.append(f"({v.ctype}*)a[{k}]")
            elif v.ftype == "character":
                args.append('""')
            else:
                args.append(f"*({v.ctype}*)a[{k}]")
        self.emit(f"    {cname(sub.name)}({', '.join(args)});")
        self.emit("  } catch (const FFatal& e) { last_fatal = e.what; return 1; }")
        self.emit("  return 0;")
        self.emit("}")

    def sig_json(self, sub):
        return [{"name": a, "type": sub.vars[a].ftype, "rank": sub.vars[a].rank, "intent": sub.vars[a].intent,
                 "dims": sub.vars[a].dims} for a in sub.args]


RUNTIME = r'''// GENERATED by oracle/f90_to_cpp.py from the reference's Fortran -- do not edit, do not commit.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <initializer_list>
#include <limits>
#include <string>
#include <type_traits>
#include <vector>

#ifndef F_REAL_T
#define F_REAL_T float
#endif
#define F_REAL(x) ((F_REAL_T)(x))
struct FFatal { std::string what; };
static std::string last_fatal;
[[noreturn]] static void f_fatal(const char* msg) { throw FFatal{msg}; }

template <class T> struct FView {
  T* p; int lo; int n;
  FView() : p(nullptr), lo(1), n(0) {}
  FView(T* p_, int lo_, int hi_) : p(p_), lo(lo_), n(hi_ - lo_ + 1) {}
  T& operator()(int i) const {
#ifdef F_BOUNDS
    if (i < lo || i >= lo + n) { char b[96]; snprintf(b, sizeof b, "index %d outside %d:%d", i, lo, lo + n - 1); f_fatal(b); }
#endif
    return p[i - lo];
  }
  int hi() const { return lo + n - 1; }
  int size() const { return n; }
  void fill(T v) { for (int i = 0; i < n; ++i) p[i] = v; }
};
template <class T> struct FArr : FView<T> {
  std::vector<T> store;
  FArr(int lo_, int hi_) : store(hi_ >= lo_ ? hi_ - lo_ + 1 : 0) { this->p = store.data(); this->lo = lo_; this->n = (int)store.size(); }
  FArr(int lo_, int hi_, T v) : FArr(lo_, hi_) { this->fill(v); }
  FArr(int lo_, int hi_, std::initializer_list<T> l) : FArr(lo_, hi_) { int i = 0; for (T v : l) if (i < this->n) this->p[i++] = v; }
  FArr(const FArr&) = delete;
};
template <class T> struct FView2 {
  T* p; int lo1, n1, lo2, n2;
  FView2() : p(nullptr), lo1(1), n1(0), lo2(1), n2(0) {}
  FView2(T* p_, int l1, int h1, int l2, int h2) : p(p_), lo1(l1), n1(h1 - l1 + 1), lo2(l2), n2(h2 - l2 + 1) {}
  T& operator()(int i, int j) const {
#ifdef F_BOUNDS
    if (i < lo1 || i >= lo1 + n1 || j < lo2 || j >= lo2 + n2) f_fatal("2-d index out of bounds");
#endif
    return p[(i - lo1) + n1 * (j - lo2)];
  }
  int size() const { return n1 * n2; }
  void fill(T v) { for (int i = 0; i < n1 * n2; ++i) p[i] = v; }
};
template <class T> struct FArr2 : FView2<T> {
  std::vector<T> store;
  FArr2(int l1, int h1, int l2, int h2) : store((size_t)(h1 - l1 + 1) * (h2 - l2 + 1)) {
    this->p = store.data(); this->lo1 = l1; this->n1 = h1 - l1 + 1; this->lo2 = l2; this->n2 = h2 - l2 + 1; }
  FArr2(int l1, int h1, int l2, int h2, T v) : FArr2(l1, h1, l2, h2) { this->fill(v); }
  FArr2(const FArr2&) = delete;
};

// x ** n with integer n: libgcc's __powisf2 / __powidf2 / gfortran's pow_i4_i4
static inline float f_pow(float x, int m) {
  unsigned n = m < 0 ? -(unsigned)m : (unsigned)m; float y = (n % 2) ? x : 1.0f;
  while (n >>= 1) { x = x * x; if (n % 2) y *= x; }
  return m < 0 ? 1.0f / y : y;
}
static inline double f_pow(double x, int m) {
  unsigned n = m < 0 ? -(unsigned)m : (unsigned)m; double y = (n % 2) ? x : 1.0;
  while (n >>= 1) { x = x * x; if (n % 2) y *= x; }
  return m < 0 ? 1.0 / y : y;
}
static inline int f_pow(int x, int m) {
  if (m < 0) return (x == 1) ? 1 : (x == -1 ? ((m & 1) ? -1 : 1) : 0);
  int y = 1; unsigned n = (unsigned)m; for (;;) { if (n & 1) y *= x; n >>= 1; if (!n) break; x *= x; } return y;
}
static inline float f_pow(float x, float y) { return powf(x, y); }
static inline double f_pow(double x, double y) { return pow(x, y); }
static inline double f_pow(double x, float y) { return pow(x, (double)y); }
static inline double f_pow(float x, double y) { return pow((double)x, y); }
static inline float f_pow(int x, float y) { return powf((float)x, y); }

#define F_MATH1(name, ff, fd) static inline float name(float x) { return ff(x); } static inline double name(double x) { return fd(x); }
F_MATH1(f_sqrt, sqrtf, sqrt) F_MATH1(f_exp, expf, exp) F_MATH1(f_log, logf, log) F_MATH1(f_log10, log10f, log10)
F_MATH1(f_tanh, tanhf, tanh) F_MATH1(f_atan, atanf, atan) F_MATH1(f_cos, cosf, cos) F_MATH1(f_sin, sinf, sin)
F_MATH1(f_tan, tanf, tan) F_MATH1(f_acos, acosf, acos) F_MATH1(f_asin, asinf, asin) F_MATH1(f_aint, truncf, trunc)
F_MATH1(f_anint, roundf, round)
static inline float f_abs(float x) { return fabsf(x); }
static inline double f_abs(double x) { return fabs(x); }
static inline int f_abs(int x) { return x < 0 ? -x : x; }
static inline int f_nint(float x) { return (int)lroundf(x); }
static inline int f_nint(double x) { return (int)lround(x); }
static inline int f_floor(float x) { return (int)floorf(x); }
static inline int f_floor(double x) { return (int)floor(x); }
static inline int f_ceiling(float x) { return (int)ceilf(x); }
static inline int f_ceiling(double x) { return (int)ceil(x); }
static inline bool f_isnan(float x) { return x != x; }
static inline bool f_isnan(double x) { return x != x; }
static inline float f_atan2(float y, float x) { return atan2f(y, x); }
static inline double f_atan2(double y, double x) { return atan2(y, x); }
static inline double f_pow(int x, double y) { return pow((double)x, y); }
template <class A, class B> static inline typename std::common_type<A, B>::type f_max(A a, B b) {
  typedef typename std::common_type<A, B>::type T; return (T)a > (T)b ? (T)a : (T)b; }
template <class A, class B> static inline typename std::common_type<A, B>::type f_min(A a, B b) {
  typedef typename std::common_type<A, B>::type T; return (T)a < (T)b ? (T)a : (T)b; }
static inline int f_mod(int a, int b) { return a % b; }
static inline float f_mod(float a, float b) { return fmodf(a, b); }
static inline double f_mod(double a, double b) { return fmod(a, b); }
static inline float f_sign(float a, float b) { return copysignf(fabsf(a), b); }
static inline double f_sign(double a, double b) { return copysign(fabs(a), b); }
static inline int f_sign(int a, int b) { return b >= 0 ? (a < 0 ? -a : a) : -(a < 0 ? -a : a); }
static inline F_REAL_T f_nan() { return std::numeric_limits<F_REAL_T>::quiet_NaN(); }

'''


# scalar prognostic arguments of noahmp_sflx that live in state planes, in the order refrun.py passes their plane numbers
NATIVE_SCALARS = ["albold", "sneqvo", "tah", "eah", "fwet", "canliq", "canice", "tv", "tg", "qsfc", "qsnow", "snowh", "sneqv", "zwt",
                  "wa", "wt", "wslake", "lfmass", "rtmass", "stmass", "wood", "stblcp", "fastcp", "lai", "sai", "cm", "ch", "tauss"]
NATIVE_FORCING = ["sfctmp", "sfcprs", "uu", "vv", "q2", "soldn", "lwdn", "prcp", "cosz"]
NATIVE_STATIC = ["lat", "shdfac", "tbot", "shdmax"]
NATIVE_ARRAYS = [("stc", 7), ("zsnso", 7), ("soilwat", 4), ("smc", 4), ("snice", 3), ("snliq", 3)]


def native_driver(tr, sub):
    """A column loop in C++ around the translated noahmp_sflx (same duties as RefModel.run, OpenMP over columns): the
    declarations and the call come from the parsed signature, the glue refers to the reference's argument names."""
    e = tr.emit
    e("struct RefRunParams { double julian0; float dt; float zsoil[4]; float zref, co2_frac, o2_frac, foln; int yearlen; };")
    e('extern "C" long long ref_run_columns(long long ncol, int nsteps, float* state, const float* forcing, const float* static_f,')
    e("    const int* static_i, float* diag, unsigned char* fatal, const RefRunParams* rp, const int* sp, const int* ap, int isnow_plane,")
    e("    const int* fp, const int* cp) {")
    e("  long long nfatal = 0;")
    e("#pragma omp parallel for schedule(dynamic, 16) reduction(+ : nfatal)")
    e("  for (long long i = 0; i < ncol; ++i) {")
    for a in sub.args:
        v = sub.vars[a]
        if v.rank == 0:
            e(f"    {v.ctype} {cname(a)}{{}};")
    for nm, n in NATIVE_ARRAYS + [("zsoil", 4), ("ficeold", 3)]:
        e(f"    {REAL} {nm}[{n}];")
    e("    iloc = 1; jloc = 1; nsoil = 4; nsnow = 3; dt = rp->dt; foln = rp->foln; yearlen = rp->yearlen; zlvl = rp->zref;")
    e("    for (int k = 0; k < 4; ++k) zsoil[k] = rp->zsoil[k];")
    for k, nm in enumerate(NATIVE_STATIC):
        e(f"    {nm} = static_f[(size_t)cp[{k}] * ncol + i];")
    e("    { const int ty = static_i[i]; lutyp = ty & 0xff; sltyp = (ty >> 8) & 0xff; slptyp = (ty >> 16) & 0xff; isc = (ty >> 24) & 0xf;")
    e("      ist = ((ty >> 28) & 1) + 1; ice = (ty >> 29) & 1; }")
    for k, nm in enumerate(NATIVE_SCALARS):
        e(f"    {cname(nm)} = state[(size_t)sp[{k}] * ncol + i];")
    e("    memcpy(&isnow, &state[(size_t)isnow_plane * ncol + i], 4);")
    e("    // (fp64 twin: the planes stay fp32 at the boundary, everything between load and store is double)")
    for k, (nm, n) in enumerate(NATIVE_ARRAYS):
        e(f"    for (int k = 0; k < {n}; ++k) {nm}[k] = state[(size_t)(ap[{k}] + k) * ncol + i];")
    e("    bool dead = false;")
    e("    for (int s = 0; s < nsteps && !dead; ++s) {")
    for k, nm in enumerate(NATIVE_FORCING):
        e(f"      {nm} = forcing[((size_t)s * {len(NATIVE_FORCING)} + fp[{k}]) * ncol + i];")
    e("      psfc = sfcprs; co2air = rp->co2_frac * sfcprs; o2air = rp->o2_frac * sfcprs;")
    e("      { const double jd = rp->julian0 + (double)s * (double)rp->dt / 86400.0, yl = (double)rp->yearlen;")
    e("        julian = (float)(jd - yl * floor(jd / yl)); }")
    e(f"      for (int k = -2; k <= 0; ++k) ficeold[k + 2] = (k >= isnow + 1) ? snice[k + 2] / (snice[k + 2] + snliq[k + 2]) : ({REAL})0;")
    e("      try {")
    e(f"        noahmp_sflx({', '.join(cname(a) for a in sub.args)});")
    e("      } catch (const FFatal&) { dead = true; }")
    e("    }")
    e("    if (fatal) fatal[i] = dead;")
    e("    if (dead) { nfatal += 1; continue; }")
    for k, nm in enumerate(NATIVE_SCALARS):
        e(f"    state[(size_t)sp[{k}] * ncol + i] = (float){cname(nm)};")
    e("    memcpy(&state[(size_t)isnow_plane * ncol + i], &isnow, 4);")
    for k, (nm, n) in enumerate(NATIVE_ARRAYS):
        e(f"    for (int k = 0; k < {n}; ++k) state[(size_t)(ap[{k}] + k) * ncol + i] = (float){nm}[k];")
    outs = [a for a in sub.args if sub.vars[a].intent == "out"]
    e("    if (diag) {")
    for k, nm in enumerate(outs):
        e(f"      diag[(size_t){k} * ncol + i] = (float){cname(nm)};")
    e("    }")
    e("  }")
    e("  return nfatal;")
    e("}")


def translate(ref_root: str, out_cpp: str, out_json: str, real_type: str = "float"):
    global REAL
    REAL = real_type
    core = os.path.join(ref_root, "core")
    mods = {}
    for fn, want in (("module_noahmp_const.f90", False), ("module_noahmp_global.f90", True),
                     ("module_noahmp_gen_param.f90", False), ("module_noahmp_soil_param.f90", False),
                     ("module_noahmp_veg_param.f90", False), ("module_noahmp_func.f90", True)):
        m = parse_file(os.path.join(core, fn), want_subs=want)
        mods[m.name] = m
    tr = Translator(mods)
    tr.out.append(f"#define F_REAL_T {real_type}\n" + RUNTIME)
    tr.emit("namespace ref {")
    for m in mods.values():
        tr.emit(f"// ---- module {m.name}")
        tr.emit_module_vars(m)
    tr.emit()
    # forward declarations
    for s in tr.subs.values():
        if s.name not in SKIP_SUBS:
            tr.emit(f"void {cname(s.name)}({tr.signature(s)});")
    tr.emit()
    for s in tr.subs.values():
        tr.emit_sub(s)
    desc = tr.registry()
    sigs = {}
    for s in tr.subs.values():
        if s.name in SKIP_SUBS:
            continue
        tr.c_entry(s)
        sigs[s.name] = tr.sig_json(s)
    native_driver(tr, tr.subs["noahmp_sflx"])
    tr.emit("}  // namespace ref")
    tr.emit(r'''
extern "C" int ref_set(const char* name, const void* data, int count) {
  for (ref::RegEntry* e = ref::registry; e->name; ++e)
    if (!strcmp(e->name, name)) {
      if (count != e->count) return -2;
      if (e->type == 0 && sizeof(F_REAL_T) == 8) {          // fp64 twin: `real` module variables are doubles, the caller sends floats
        for (int k = 0; k < count; ++k) ((double*)e->ptr)[k] = (double)((const float*)data)[k];
        return 0;
      }
      memcpy(e->ptr, data, (size_t)count * (e->type == 2 ? 8 : 4));
      return 0;
    }
  return -1;
}
extern "C" const char* ref_last_fatal() { return last_fatal.c_str(); }
''')
    os.makedirs(os.path.dirname(out_cpp), exist_ok=True)
    with open(out_cpp, "w") as fh:
        fh.write("\n".join(tr.out) + "\n")
    params = {}
    for m in mods.values():
        for n, v in m.vars.items():
            if v.param and v.rank == 0 and v.ftype == "integer" and v.init and v.init.isdigit():
                params[n] = int(v.init)
    with open(out_json, "w") as fh:
        json.dump({"module_vars": desc, "subs": sigs, "int_params": params}, fh, indent=1)
    return tr


if __name__ == "__main__":
    ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    here = os.path.dirname(os.path.abspath(__file__))
    out = os.path.join(here, "_ref")
    t = translate(ref, os.path.join(out, "noahmp_ref.cpp"), os.path.join(out, "noahmp_ref.json"))
    print(f"translated {len(t.subs)} subroutines -> {out}/noahmp_ref.cpp ({len(t.out)} lines)")
