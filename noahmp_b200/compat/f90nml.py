"""Minimal stand-in for the third-party ``f90nml`` package (absent from this image; the reference imports
it at ``offline/noahmp_config.py:5`` and uses exactly ``f90nml.read(path)[GROUP][key]`` and ``key in group``).

Supports what a Noah-MP offline namelist needs: ``&group ... /`` blocks, ``key = value`` pairs separated by
newlines and/or commas, ``!`` comments, integers, reals (incl. ``d`` exponents), quoted strings, logicals and
comma-separated value lists.  Group and key lookup is case-insensitive, like Fortran namelists (and f90nml).
"""
from __future__ import annotations

import re


class _CaseInsensitiveDict(dict):
    def __getitem__(self, key):
        return super().__getitem__(key.lower())

    def __setitem__(self, key, value):
        super().__setitem__(key.lower(), value)

    def __contains__(self, key):
        return super().__contains__(key.lower())

    def get(self, key, default=None):
        return super().get(key.lower(), default)


class Namelist(_CaseInsensitiveDict):
    pass


_TOK = re.compile(r"""\s*(?:'((?:[^']|'')*)'|"((?:[^"]|"")*)"|([^\s,'"]+))\s*,?""")


def _convert(tok: str, quoted: bool):
    if quoted:
        return tok
    low = tok.lower()
    if low in (".true.", "t", ".t."):
        return True
    if low in (".false.", "f", ".f."):
        return False
    try:
        return int(tok)
    except ValueError:
        pass
    try:
        return float(low.replace("d", "e"))
    except ValueError:
        return tok


def _strip_comment(line: str) -> str:
    out, q = [], None
    for ch in line:
        if q:
            if ch == q:
                q = None
        elif ch in "'\"":
            q = ch
        elif ch == "!":
            break
        out.append(ch)
    return "".join(out)


def _values(text: str):
    vals, pos = [], 0
    while pos < len(text):
        m = _TOK.match(text, pos)
        if not m or m.end() == pos:
            break
        if m.group(1) is not None:
            vals.append(_convert(m.group(1).replace("''", "'"), True))
        elif m.group(2) is not None:
            vals.append(_convert(m.group(2).replace('""', '"'), True))
        else:
            vals.append(_convert(m.group(3), False))
        pos = m.end()
    return vals


def reads(text: str) -> Namelist:
    nml = Namelist()
    body = "\n".join(_strip_comment(ln) for ln in text.splitlines())
    for gm in re.finditer(r"&\s*(\w+)(.*?)(?:^\s*/|\s/\s*$|&end)", body, re.S | re.M | re.I):
        group = _CaseInsensitiveDict()
        content = gm.group(2)
        keys = list(re.finditer(r"(\w+(?:\s*\([^)]*\))?)\s*=", content))
        for i, km in enumerate(keys):
            end = keys[i + 1].start() if i + 1 < len(keys) else len(content)
            vals = _values(content[km.end():end].replace("\n", " "))
            group[km.group(1).strip()] = vals[0] if len(vals) == 1 else (vals if vals else None)
        nml[gm.group(1)] = group
    return nml


def read(path) -> Namelist:
    with open(path, "r") as fh:
        return reads(fh.read())
