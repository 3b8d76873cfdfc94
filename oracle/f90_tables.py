"""TEST INFRASTRUCTURE ONLY -- runs the reference's three `*_readptable` routines FROM THEIR SOURCE TEXT.

The column physics reads ~90 module arrays (`LK_*`, `KK_*`) that the reference fills in
`noahmp_{gen,soil,veg}_param_readptable` (core/module_noahmp_gen_param.f90:51-89, soil_param.f90:31-72,
veg_param.f90:77-161).  Which table column lands in which array, the section names, the size limits, the derived fields
(`LK_KDT`, `LK_FRZX` with its `0.412 / 0468`) and the text of every error message are facts of that source; the product's
parser (noahmp_b200/tables.py) restates them by hand.  This module brings them inside the pin: it parses the reference's
Fortran text (declarations, `call noahmp_ptable_read_var / read_tablestr / assert`, `do` loops around
`read(tblstr(ii), *, iostat=ierr) <item list with implied-do>`, the two array assignments) and executes it statement by
statement, with no knowledge of any table.  tests/test_tables.py compares every array it produces, for all four tag
combinations, with the product parser, and feeds both the same malformed files.

Two layers, and what each is:
* interpreted from /root/reference at run time: the three readptable routines and the declarations of the modules;
* restated by hand here, citing lines: the I/O helpers of core/module_noahmp_utils.f90:56-262 (they are Fortran I/O, not
  arithmetic) and list-directed READ itself (Fortran 2008 10.10.3: separators blank / comma / slash, `r*c` repeats, null
  values, quoted strings, D exponents) - there is no Fortran runtime in this image.
"""
from __future__ import annotations

import os
import re

import numpy as np

f32 = np.float32


class FortranStop(RuntimeError):
    """`call assert(.false., msg)` -> write(*,*) msg; stop  (utils:43-53)."""


# ---------------------------------------------------------------------------------------------------------------------
# list-directed input (F2008 10.10.3)
# ---------------------------------------------------------------------------------------------------------------------
class EndOfRecord(Exception):
    pass


def ld_values(text: str):
    """The value sequence of list-directed records `text`: yields str tokens, None for a null value; a slash ends input."""
    i, n = 0, len(text)
    out = []
    expect_value = True            # after a comma a missing value is a null
    while i < n:
        ch = text[i]
        if ch in " \t\r\n":
            i += 1
            continue
        if ch == "/":
            out.append("/")
            break
        if ch == ",":
            if expect_value:
                out.append(None)
            expect_value = True
            i += 1
            continue
        # repeat count r* or r*c
        m = re.match(r"(\d+)\*", text[i:])
        rep = 1
        if m:
            rep = int(m.group(1))
            i += m.end()
            if i >= n or text[i] in " \t\r\n,/":
                out.extend([None] * rep)
                expect_value = False
                continue
            ch = text[i]
        if ch in "'\"":
            q, j, buf = ch, i + 1, []
            while j < n:
                if text[j] == q:
                    if j + 1 < n and text[j + 1] == q:
                        buf.append(q)
                        j += 2
                        continue
                    break
                buf.append(text[j])
                j += 1
            tok = "".join(buf)
            i = j + 1
            out.extend([("str", tok)] * rep)
        else:
            j = i
            while j < n and text[j] not in " \t\r\n,/":
                j += 1
            out.extend([text[i:j]] * rep)
            i = j
        expect_value = False
    return out


_INT = re.compile(r"[+-]?\d+$")
_REAL = re.compile(r"[+-]?(\d+\.?\d*|\.\d+)([EeDd][+-]?\d+|[+-]\d+)?$")


def ld_convert(tok, kind):
    """One list item: kind 'integer' | 'real' | 'character'.  Raises ValueError where the runtime would set iostat."""
    if isinstance(tok, tuple):
        if kind != "character":
            raise ValueError("character constant for a numeric item")
        return tok[1]
    if kind == "character":
        return tok
    if kind == "integer":
        if not _INT.match(tok):
            raise ValueError(f"bad integer {tok!r}")
        return int(tok)
    m = _REAL.match(tok)
    if not m:
        raise ValueError(f"bad real {tok!r}")
    mant, ex = m.group(1), m.group(2)
    sign = -1.0 if tok.startswith("-") else 1.0
    e = 0
    if ex:
        e = int(ex[1:]) if ex[0] in "EeDd" else int(ex)
    return f32(sign * float(mant) * 10.0 ** e) if e == 0 else f32(float(f"{'-' if sign < 0 else ''}{mant}e{e}"))


def ld_read(records, kinds):
    """`read(unit, *, iostat=ierr) items`: consume as many records as the item list needs.  `records` is an iterator of lines.
    Returns (values, ierr); a null value leaves the item undefined (returned as None)."""
    vals = []
    pending = []
    try:
        while len(vals) < len(kinds):
            if not pending:
                rec = next(records, None)
                if rec is None:
                    return vals, -1            # end of file
                pending = ld_values(rec)
                continue
            tok = pending.pop(0)
            if tok == "/":
                vals.extend([None] * (len(kinds) - len(vals)))
                break
            vals.append(None if tok is None else ld_convert(tok, kinds[len(vals)]))
    except ValueError:
        return vals, 5010                      # a positive iostat, as libgfortran's LIBERROR_READ_VALUE
    return vals, 0


# ---------------------------------------------------------------------------------------------------------------------
# core/module_noahmp_utils.f90, restated
# ---------------------------------------------------------------------------------------------------------------------
def _assert(good, msg):                                                        # utils:43-53
    if not good:
        raise FortranStop(msg)


def nametag(name, tag):                                                        # utils:240-254
    if len(name.strip()) == 0:
        return " "
    if len(tag.strip()) == 0:
        return name.strip()
    return name.strip() + "#" + tag.strip()


def find_name_tag(lines, name, tag):                                           # utils:200-237
    """-> index of the record after the header, or None.  `read(iounit, *) strbuf` takes the first list item of the
    next non-empty record as a character value (blank records are skipped by list-directed input)."""
    if len(name.strip()) == 0:
        return None
    i = 0
    while i < len(lines):
        vals = ld_values(lines[i])
        i += 1
        if not vals:
            continue
        first = vals[0]
        if first is None or first == "/":
            continue
        strbuf = first[1] if isinstance(first, tuple) else first
        if strbuf[:1] == "&":
            tagloc = strbuf.find("#") + 1                                      # scan(): 1-based, 0 = absent
            found = False
            if tagloc == 0 and len(tag.strip()) == 0:
                found = strbuf[1:].strip() == name.strip()
            elif tagloc != 0 and len(tag.strip()) != 0:
                found = strbuf[1:tagloc - 1].strip() == name.strip() and strbuf[tagloc:].strip() == tag.strip()
            if found:
                return i
    return None


def _open(directory, filename):                                                # utils:78-79
    path = os.path.join(directory, filename.strip())
    try:
        with open(path) as fh:
            return fh.read().splitlines()
    except OSError:
        _assert(False, "Unable to open file parameter file " + filename.strip())


def read_tablestr(directory, filename, name, tag, size):                       # utils:56-97
    lines = _open(directory, filename)
    nt = nametag(name, tag)
    at = find_name_tag(lines, name, tag)
    _assert(at is not None, "Unable to find " + nt.strip() + " in file " + filename.strip())
    it = iter(lines[at:])
    consumed = [0]

    def counting():
        for ln in it:
            consumed[0] += 1
            yield ln
    vals, ierr = ld_read(counting(), ["integer"])
    _assert(ierr == 0 and vals and vals[0] is not None,
            "Unable to read the number of rows for table " + nt.strip() + " in file " + filename.strip())
    nrow = vals[0]
    _assert(nrow <= size, "Table " + nt.strip() + " in file " + filename.strip() + " is too big to be read")
    rows = []
    at += consumed[0]
    for _ in range(nrow):                                                      # read(iounit, '(A)') : next physical record
        _assert(at < len(lines), "Unable to read table " + nt.strip() + " from file " + filename.strip())
        rows.append(lines[at][:512])
        at += 1
    return nrow, rows


def read_var(directory, filename, name, tag, kind, count, what_table):         # utils:100-197
    lines = _open(directory, filename)
    nt = nametag(name, tag)
    at = find_name_tag(lines, name, tag)
    _assert(at is not None, "Unable to find " + nt.strip() + " in file " + filename.strip())
    vals, ierr = ld_read(iter(lines[at:]), [kind] * count)
    # int0d says "Unable to read table ..." (utils:117), the other three "Unable to read ..." (utils:141,165,189)
    _assert(ierr == 0, ("Unable to read table " if what_table else "Unable to read ") + nt.strip() + " from file " + filename.strip())
    return vals


# ---------------------------------------------------------------------------------------------------------------------
# the interpreter of the readptable routines
# ---------------------------------------------------------------------------------------------------------------------
def _logical_lines(src: str):
    """Fortran free-form statements: comments stripped (outside strings), `&` continuations joined."""
    out, cur = [], ""
    for raw in src.splitlines():
        line, q, buf = raw, None, []
        for ch in line:
            if q:
                buf.append(ch)
                if ch == q:
                    q = None
                continue
            if ch in "'\"":
                q = ch
                buf.append(ch)
                continue
            if ch == "!":
                break
            buf.append(ch)
        s = "".join(buf).strip()
        if not s:
            continue
        if s.startswith("&"):
            s = s[1:].lstrip()
        if s.endswith("&"):
            cur += s[:-1].rstrip() + " "
            continue
        out.append(cur + s)
        cur = ""
    return out


def _split_args(s: str):
    """Top-level comma split (parentheses and quotes respected)."""
    out, depth, q, cur = [], 0, None, []
    for ch in s:
        if q:
            cur.append(ch)
            if ch == q:
                q = None
            continue
        if ch in "'\"":
            q = ch
            cur.append(ch)
        elif ch == "(":
            depth += 1
            cur.append(ch)
        elif ch == ")":
            depth -= 1
            cur.append(ch)
        elif ch == "," and depth == 0:
            out.append("".join(cur).strip())
            cur = []
        else:
            cur.append(ch)
    if "".join(cur).strip():
        out.append("".join(cur).strip())
    return out


class Module:
    """Declarations of one parameter module: integer parameters, and arrays / scalars with type and declared extents."""

    def __init__(self, src: str, params: dict):
        self.params = dict(params)
        self.vars = {}      # lower name -> (kind, dims tuple)
        body = src.split("contains")[0]
        for st in _logical_lines(body):
            m = re.match(r"(integer|real\s*\(\s*r4\s*\)|real)\s*(,\s*parameter)?\s*(?:,\s*\w+\s*)*::\s*(.*)$", st, re.I)
            if not m:
                continue
            kind = "integer" if m.group(1).lower().startswith("integer") else "real"
            for decl in _split_args(m.group(3)):
                mm = re.match(r"(\w+)\s*(\(([^)]*)\))?\s*(=\s*(.*))?$", decl)
                if not mm:
                    continue
                name, dims, init = mm.group(1).lower(), mm.group(3), mm.group(5)
                if m.group(2):
                    try:
                        self.params[name] = int(eval(init, {}, {k: v for k, v in self.params.items()}))  # noqa: S307 - reference source
                    except Exception:
                        pass
                    continue
                shape = tuple(self._dim(d) for d in _split_args(dims)) if dims else ()
                self.vars[name] = (kind, shape)

    def _dim(self, d: str) -> int:
        d = d.strip().lower()
        return int(d) if d.isdigit() else int(self.params[d])


class Reader:
    """Executes `subroutine noahmp_*_param_readptable` of one module against a directory of .TBL files."""

    def __init__(self, ref_root: str = "/root/reference"):
        core = os.path.join(ref_root, "core")
        rd = lambda f: open(os.path.join(core, f)).read()  # noqa: E731
        glob = Module(rd("module_noahmp_global.f90"), {})
        self.src = {k: rd(f"module_noahmp_{k}_param.f90") for k in ("gen", "soil", "veg")}
        self.mods = {k: Module(self.src[k], glob.params) for k in self.src}
        self.values = {}    # lower name -> numpy array / scalar, as the module variables hold them afterwards
        for k, m in self.mods.items():
            for name, (kind, shape) in m.vars.items():
                if kind == "real":
                    self.values[name] = np.full(shape, np.nan, f32) if shape else f32(np.nan)
                else:
                    self.values[name] = np.zeros(shape, np.int32) if shape else 0

    def kind_of(self, name):
        for m in self.mods.values():
            if name in m.vars:
                return m.vars[name]
        raise KeyError(name)

    # -- string expressions of the assert messages: 'text'//trim(x)//... ------------------------------------------------
    def _str(self, expr: str, env: dict) -> str:
        parts = []
        for piece in re.split(r"//", expr):
            piece = piece.strip()
            m = re.match(r"trim\((\w+)\)$", piece, re.I)
            if m:
                parts.append(str(env[m.group(1).lower()]).rstrip())
            elif piece[:1] in "'\"":
                parts.append(piece[1:-1])
            else:
                parts.append(str(env[piece.lower()]))
        return "".join(parts)

    def _expand_items(self, items, env):
        """READ item list -> [(array name or None for a scalar local, index tuple)], implied-do loops unrolled."""
        out = []
        for it in items:
            it = it.strip()
            m = re.match(r"\(\s*(.+?)\s*,\s*(\w+)\s*=\s*(\w+)\s*,\s*(\w+)\s*\)$", it)
            if m:                                                              # (LK_X(jj,ii), jj=1,nband)
                var, lo, hi = m.group(2).lower(), self._int(m.group(3), env), self._int(m.group(4), env)
                for v in range(lo, hi + 1):
                    out.extend(self._expand_items(_split_args(m.group(1)), {**env, var: v}))
                continue
            m = re.match(r"(\w+)\s*\(([^)]*)\)$", it)
            if m:
                idx = tuple(self._int(a, env) - 1 for a in _split_args(m.group(2)))
                out.append((m.group(1).lower(), idx))
            else:
                out.append((it.lower(), None))
        return out

    def _int(self, s, env):
        s = s.strip().lower()
        if re.fullmatch(r"\d+", s):
            return int(s)
        if s in env:
            return int(env[s])
        for m in self.mods.values():
            if s in m.params:
                return int(m.params[s])
        return int(self.values[s])

    def _array_expr(self, expr: str):
        """Right-hand side of the two whole-array assignments (soil_param:58-61), evaluated in default real (fp32): real
        literals are fp32, an integer literal in a real expression is converted to real (`0468` -> 468.0)."""
        e = re.sub(r"\(\s*:\s*\)", "", expr)
        e = re.sub(r"(?<![\w.])(\d+\.\d*(?:[eE][+-]?\d+)?|\.\d+(?:[eE][+-]?\d+)?)", r"f32(\1)", e)
        e = re.sub(r"(?<![\w.(])0*(\d+)(?![\w.])", r"f32(\1)", e)
        env = {"f32": f32}
        for name in set(re.findall(r"[A-Za-z_]\w*", e)) - {"f32"}:
            env[name] = self.values[name.lower()]
        with np.errstate(divide="ignore", invalid="ignore"):
            return eval(e, {"__builtins__": {}}, env)  # noqa: S307 - expression text of the reference source

    def run(self, which: str, directory: str, tag: str = ""):
        src = self.src[which]
        m = re.search(r"subroutine\s+noahmp_%s_param_readptable\s*\((.*?)\)(.*?)end\s+subroutine" % which, src, re.S | re.I)
        body = m.group(2)
        env = {"tag": tag}
        stmts = _logical_lines(body)
        # character parameters (the file names)
        for st in stmts:
            mm = re.match(r"character\s*\([^)]*\)\s*,\s*parameter\s*::\s*(\w+)\s*=\s*'([^']*)'", st, re.I)
            if mm:
                env[mm.group(1).lower()] = mm.group(2)
        tbl = {"rows": []}
        i = 0
        while i < len(stmts):
            st = stmts[i]
            low = st.lower()
            if low.startswith("call noahmp_ptable_read_tablestr"):
                a = _split_args(st[st.index("(") + 1:st.rindex(")")])
                fname, name, tg, nvar = env[a[0].lower()], a[1][1:-1], (env["tag"] if a[2].lower() == "tag" else a[2][1:-1]), a[3].lower()
                size = max(self._int(x, env) for x in re.findall(r"\w+", re.search(r"tblstr\s*\(([^)]*(?:\([^)]*\))?[^)]*)\)", body, re.I).group(1)) if not x.lower() == "max")
                nrow, rows = read_tablestr(directory, fname, name, tg, size)
                self.values[nvar] = nrow
                tbl["rows"] = rows
            elif low.startswith("call noahmp_ptable_read_var"):
                a = _split_args(st[st.index("(") + 1:st.rindex(")")])
                fname, name, tg, var = env[a[0].lower()], a[1][1:-1], (env["tag"] if a[2].lower() == "tag" else a[2][1:-1]), a[3].lower()
                kind, shape = self.kind_of(var)
                count = int(np.prod(shape)) if shape else 1
                vals = read_var(directory, fname, name, tg, kind, count, what_table=(kind == "integer" and not shape))
                if shape:
                    arr = self.values[var]
                    for k, v in enumerate(vals):
                        if v is not None:
                            arr.reshape(-1, order="F")[k] = v
                elif vals and vals[0] is not None:
                    self.values[var] = vals[0]
            elif low.startswith("call assert"):
                a = _split_args(st[st.index("(") + 1:st.rindex(")")])
                cond = a[0].lower()
                mm = re.match(r"(\w+)\s*(<=|==)\s*(\w+)$", cond)
                lhs, op, rhs = mm.group(1), mm.group(2), mm.group(3)
                lv = env["ierr"] if lhs == "ierr" else self._int(lhs, env)
                rv = self._int(rhs, env)
                _assert(lv <= rv if op == "<=" else lv == rv, self._str(a[1], env))
            elif low.startswith("do "):
                mm = re.match(r"do\s+(\w+)\s*=\s*(\w+)\s*,\s*(\w+)$", low)
                var, lo, hi = mm.group(1), self._int(mm.group(2), env), self._int(mm.group(3), env)
                j = i + 1
                while not stmts[j].lower().replace(" ", "").startswith("enddo"):
                    j += 1
                inner = stmts[i + 1:j]
                for v in range(lo, hi + 1):
                    env[var] = v
                    for ist in inner:
                        il = ist.lower()
                        if il.startswith("read"):
                            mm2 = re.match(r"read\s*\(\s*tblstr\s*\(\s*(\w+)\s*\)\s*,\s*\*\s*,\s*iostat\s*=\s*ierr\s*\)\s*(.*)$", ist, re.I)
                            row = tbl["rows"][self._int(mm2.group(1), env) - 1]
                            items = self._expand_items(_split_args(mm2.group(2)), env)
                            kinds = ["integer" if (nm == "index" or (idx is not None and self.kind_of(nm)[0] == "integer")) else "real"
                                     for nm, idx in items]
                            vals, ierr = ld_read(iter([row]), kinds)
                            env["ierr"] = ierr
                            for (nm, idx), val in zip(items, vals):
                                if val is None or idx is None:
                                    continue
                                self.values[nm][idx] = val
                        elif il.startswith("call assert"):
                            a = _split_args(ist[ist.index("(") + 1:ist.rindex(")")])
                            _assert(env["ierr"] == 0, self._str(a[1], env))
                        else:
                            raise NotImplementedError(ist)
                i = j
            elif re.match(r"\w+\s*\(\s*:\s*\)\s*=", st):                       # LK_KDT(:) = ...
                name, rhs = st.split("=", 1)
                name = re.match(r"(\w+)", name).group(1).lower()
                self.values[name][...] = self._array_expr(rhs)
            elif low.startswith("where"):
                cond = re.match(r"where\s*\((.*)\)$", st, re.I).group(1)
                mask = self._array_expr(cond)
                asg = stmts[i + 1]
                name, rhs = asg.split("=", 1)
                val = self._array_expr(rhs)
                tgt = self.values[name.strip().lower()]
                tgt[mask] = np.asarray(val, f32)[mask]
                i += 2                                                         # the assignment and `end where`
            elif re.match(r"(use|implicit|integer|character|real|logical)\b", low):
                pass
            else:
                raise NotImplementedError(f"statement outside the interpreted subset: {st}")
            i += 1
        return self


def read_all(directory: str, soil_tag: str, veg_tag: str, ref_root: str = "/root/reference") -> dict:
    """Module variables after gen -> soil -> veg readptable, as {UPPER name: array / scalar}."""
    r = Reader(ref_root)
    r.run("gen", directory)
    r.run("soil", directory, soil_tag)
    r.run("veg", directory, veg_tag)
    return {k.upper(): v for k, v in r.values.items()}
