"""GENPARMMP / SOILPARMMP / VEGPARMMP ``.TBL`` reader and packer.

Host-side replacement for the reference's Fortran table readers:

* section lookup rule ........ ``core/module_noahmp_utils.f90:200-237`` (``noahmp_ptable_find_name_tag``)
* table / scalar reads ....... ``core/module_noahmp_utils.f90:56-97, 100-197``
* GENPARM field list ......... ``core/module_noahmp_gen_param.f90:63-88``
* SOILPARM rows + derived .... ``core/module_noahmp_soil_param.f90:47-71`` (incl. ``0.412 / 0468`` at :60)
* VEGPARM sub-tables ......... ``core/module_noahmp_veg_param.f90:89-160``

The parsed values are packed into flat fp32 / int32 arrays whose slot order is the C ABI contract of
``include/noahmp_b200.h`` (``noahmp_set_tables``).  Errors raise :class:`TableError` with the same
messages the Fortran ``assert`` would print before ``stop`` (``utils:43-53``).
"""
from __future__ import annotations

import json
import os
import re
from dataclasses import dataclass, field

import numpy as np

# limits of the reference's static arrays
MSLOPETYP = 30  # gen_param:8
MSLTYP = 30     # soil_param:7
MSLCOL = 20     # soil_param:8
MLUTYP = 27     # veg_param:8

GEN_FILE = "GENPARMMP.TBL"    # gen_param:58
SOIL_FILE = "SOILPARMMP.TBL"  # soil_param:40
VEG_FILE = "VEGPARMMP.TBL"    # veg_param:83

# ---- packed slot order (must match include/noahmp_b200.h) -------------------------------------
GEN_SCALARS = ["CSOIL", "ZBOT", "CZIL", "DKREF", "KDTREF", "FRZK", "TIMEAN", "FSATMAX", "MLTFCT",
               "Z0SNO", "SSI", "SWEMAX"]
GEN_SLOTS = GEN_SCALARS + ["ALBICE_VIS", "ALBICE_NIR", "ALBLAKE_VIS", "ALBLAKE_NIR", "OMEGAS_VIS",
                           "OMEGAS_NIR", "BETADS", "BETAIS", "EMSSOIL", "EMSLAKE", "NSLOPE"]
NGEN = len(GEN_SLOTS) + MSLOPETYP  # 53
SOIL_SLOTS = ["BEXP", "SMCMAX", "SMCREF", "SMCWLT", "PSISAT", "DKSAT", "DWSAT", "QUARTZ", "KDT", "FRZX"]
COLOR_SLOTS = ["ALBSAT_VIS", "ALBSAT_NIR", "ALBDRY_VIS", "ALBDRY_NIR"]
VEG_RAD = ["XL", "RHOL_VIS", "RHOL_NIR", "RHOS_VIS", "RHOS_NIR", "TAUL_VIS", "TAUL_NIR", "TAUS_VIS", "TAUS_NIR"]
VEG_PHYS = ["CANWMXP", "DLEAF", "Z0MVT", "HVT", "HVB", "DEN", "RCROWN", "CWPVT"]
VEG_DVEG = ["SLA", "DILEFC", "DILEFW", "FRAGR", "LTOVRC", "WRRAT", "WDPOOL", "TDLEF"]
# order of the PHOTO read list (veg_param:145-150), minus the leading integer C3C4
VEG_PHOTO = ["RGL", "HS", "KC25", "AKC", "KO25", "AKO", "VCMX25", "AVCMX", "BP", "RSMAX", "RSMIN", "MP",
             "QE25", "AQE", "RMF25", "RMS25", "RMR25", "FOLNMX", "TOPT", "TMIN", "ARM", "MRP"]
VEG_F_SLOTS = (VEG_RAD + VEG_PHYS + [f"LAI{m:02d}" for m in range(1, 13)] +
               [f"SAI{m:02d}" for m in range(1, 13)] + VEG_DVEG + VEG_PHOTO)
NVEGF = len(VEG_F_SLOTS)  # 71
VEG_I_SLOTS = ["NROOT", "C3C4"]
VEG_FLAGS = ["ISURBAN", "ISWATER", "ISBARREN", "ISICE", "ISEGBLF"]


class TableError(RuntimeError):
    """Raised where the reference would ``assert`` and ``stop`` (utils:43-53)."""


# ------------------------------------------------------------------------------------------------
# Fortran list-directed tokenising (just enough of it)
# ------------------------------------------------------------------------------------------------
_TOKEN = re.compile(r"""\s*(?:'((?:[^']|'')*)'|"((?:[^"]|"")*)"|([^\s,]+))\s*,?""")


def _tokens(line: str) -> list[str]:
    """Split one record the way a list-directed READ separates items (blanks and/or one comma)."""
    out, pos = [], 0
    line = line.rstrip("\r\n")
    while pos < len(line):
        m = _TOKEN.match(line, pos)
        if not m or m.end() == pos:
            break
        out.append(next(g for g in m.groups() if g is not None))
        pos = m.end()
    return out


def _find_section(lines: list[str], name: str, tag: str) -> int:
    """Index of the line AFTER the ``&NAME[#TAG]`` header (utils:200-237); -1 if absent."""
    name = name.strip()
    tag = tag.strip()
    if not name:
        return -1
    for i, raw in enumerate(lines):
        toks = _tokens(raw)
        if not toks:  # list-directed read skips blank records
            continue
        first = toks[0]
        if not first.startswith("&"):
            continue
        tagloc = first.find("#")
        if tagloc < 0 and not tag:
            if first[1:].strip() == name:
                return i + 1
        elif tagloc >= 0 and tag:
            if first[1:tagloc].strip() == name and first[tagloc + 1:].strip() == tag:
                return i + 1
    return -1


def _nametag(name: str, tag: str) -> str:
    return name.strip() if not tag.strip() else f"{name.strip()}#{tag.strip()}"


def _read_lines(path: str) -> list[str]:
    try:
        with open(path, "r") as fh:
            return fh.readlines()
    except OSError:
        raise TableError(f"Unable to open file parameter file {path}") from None  # utils:79


def _next_record(lines: list[str], start: int) -> tuple[list[str], int]:
    i = start
    while i < len(lines):
        toks = _tokens(lines[i])
        i += 1
        if toks:
            return toks, i
    return [], i


def read_var(path: str, name: str, tag: str = "", n: int = 1, kind=float):
    """Scalar / short vector after a section header (utils:100-197)."""
    lines = _read_lines(path)
    at = _find_section(lines, name, tag)
    if at < 0:
        raise TableError(f"Unable to find {_nametag(name, tag)} in file {path}")
    toks, _ = _next_record(lines, at)
    try:
        vals = [kind(_fortran_number(t)) if kind is float else int(t) for t in toks[:n]]
        if len(vals) < n:
            raise ValueError
    except ValueError:
        # the integer-scalar reader words it differently from the other three (utils:117 vs 141, 165, 189)
        what = "Unable to read table" if (kind is int and n == 1) else "Unable to read"
        raise TableError(f"{what} {_nametag(name, tag)} from file {path}") from None
    return vals[0] if n == 1 else vals


def _fortran_number(tok: str) -> float:
    return float(tok.replace("D", "E").replace("d", "e"))


def read_table(path: str, name: str, tag: str, maxrow: int) -> list[list[str]]:
    """Row strings of a counted table (utils:56-97): first token of the next record is the row count,
    the rest of that record is a header and ignored; each row is returned tokenised."""
    lines = _read_lines(path)
    at = _find_section(lines, name, tag)
    nt = _nametag(name, tag)
    if at < 0:
        raise TableError(f"Unable to find {nt} in file {path}")
    toks, at = _next_record(lines, at)
    try:
        nrow = int(toks[0])
    except (IndexError, ValueError):
        raise TableError(f"Unable to read the number of rows for table {nt} in file {path}") from None
    if nrow > maxrow:
        raise TableError(f"Table {nt} in file {path} is too big to be read")
    rows = []
    for _ in range(nrow):  # read(iounit,'(A)') takes the next physical record, blank or not
        if at >= len(lines):
            raise TableError(f"Unable to read table {nt} from file {path}")
        rows.append(_tokens(lines[at]))
        at += 1
    return rows


def _row_values(row: list[str], nfloat: int, what: str, path: str, int_cols=()) -> list:
    """``read(tblstr,*) index, v1, v2, ...`` -- trailing text is ignored."""
    try:
        int(row[0])
        out = []
        for j in range(nfloat):
            tok = row[1 + j]
            out.append(int(tok) if j in int_cols else _fortran_number(tok))
        return out
    except (IndexError, ValueError):
        raise TableError(f"Unable to read {what} from file {path}") from None


# ------------------------------------------------------------------------------------------------
@dataclass
class ParamTables:
    """All parameter values the column physics reads, in reference units."""
    gen: dict = field(default_factory=dict)          # scalars + 2-vectors + SLOPE list
    soil: dict = field(default_factory=dict)         # name -> list per soil type (1-based order)
    color: dict = field(default_factory=dict)        # name -> list per soil colour
    veg_f: dict = field(default_factory=dict)        # name -> list per land-use type
    veg_i: dict = field(default_factory=dict)
    flags: dict = field(default_factory=dict)
    soil_tag: str = "STAS"
    veg_tag: str = "USGS"

    # ---- readers --------------------------------------------------------------------------------
    @classmethod
    def from_tbl_dir(cls, tbl_dir: str, soil_tag: str = "STAS", veg_tag: str = "USGS") -> "ParamTables":
        """Order matters as in the reference: gen, then soil (needs KDTREF/DKREF/FRZK), then veg."""
        t = cls(soil_tag=soil_tag, veg_tag=veg_tag)
        t._read_gen(os.path.join(tbl_dir, GEN_FILE))
        t._read_soil(os.path.join(tbl_dir, SOIL_FILE), soil_tag)
        t._read_veg(os.path.join(tbl_dir, VEG_FILE), veg_tag)
        return t

    def _read_gen(self, path: str) -> None:  # gen_param:63-88
        rows = read_table(path, "SLOPE", "", MSLOPETYP)
        self.gen["SLOPE"] = [_row_values(r, 1, "SLOPE", path)[0] for r in rows]
        for nm in ["CSOIL", "DKREF", "KDTREF", "FRZK", "ZBOT", "CZIL", "TIMEAN", "FSATMAX", "MLTFCT",
                   "Z0SNO", "SSI", "SWEMAX"]:
            self.gen[nm] = read_var(path, nm)
        for nm in ["ALBICE", "ALBLAKE", "OMEGAS"]:
            self.gen[nm] = read_var(path, nm, n=2)
        for nm in ["BETADS", "BETAIS", "EMSSOIL", "EMSLAKE"]:
            self.gen[nm] = read_var(path, nm)

    def _read_soil(self, path: str, tag: str) -> None:  # soil_param:47-71
        rows = read_table(path, "PARM", tag, MSLTYP)
        names = ["BEXP", "SMCMAX", "SMCREF", "SMCWLT", "PSISAT", "DKSAT", "DWSAT", "QUARTZ"]
        cols = [_row_values(r, 8, f"PARM.{tag}", path) for r in rows]
        for j, nm in enumerate(names):
            self.soil[nm] = [c[j] for c in cols]
        f32 = np.float32
        dksat = np.asarray(self.soil["DKSAT"], f32)
        smcmax = np.asarray(self.soil["SMCMAX"], f32)
        smcref = np.asarray(self.soil["SMCREF"], f32)
        # LK_KDT(:) = KK_KDTREF * LK_DKSAT(:) / KK_DKREF                        soil_param:58
        kdt = f32(self.gen["KDTREF"]) * dksat / f32(self.gen["DKREF"])
        # where (LK_SMCREF > 0) LK_FRZX = KK_FRZK*(LK_SMCMAX/LK_SMCREF)*(0.412/0468)   soil_param:59-61
        # `0468` is an INTEGER literal (468): the reference's quirk is reproduced, not "fixed".
        with np.errstate(divide="ignore", invalid="ignore"):
            frzx = f32(self.gen["FRZK"]) * (smcmax / smcref) * (f32(0.412) / f32(468))
        frzx = np.where(smcref > 0, frzx, f32(np.nan)).astype(f32)  # untouched entries stay nan4
        self.soil["KDT"] = [float(v) for v in kdt]
        self.soil["FRZX"] = [float(v) for v in frzx]
        rows = read_table(path, "COLOR", "", MSLCOL)
        cols = [_row_values(r, 4, "COLOR", path) for r in rows]
        for j, nm in enumerate(COLOR_SLOTS):
            self.color[nm] = [c[j] for c in cols]

    def _read_veg(self, path: str, tag: str) -> None:  # veg_param:89-160
        for nm in VEG_FLAGS:
            self.flags[nm] = read_var(path, nm, tag, kind=int)

        def table(name, names, int_cols=()):
            rows = read_table(path, name, tag, MLUTYP)
            cols = [_row_values(r, len(names), f"{name}#{tag}", path, int_cols) for r in rows]
            return {nm: [c[j] for c in cols] for j, nm in enumerate(names)}

        self.veg_f.update(table("RAD", VEG_RAD))
        self.veg_f.update(table("LAI12M", [f"LAI{m:02d}" for m in range(1, 13)]))
        self.veg_f.update(table("SAI12M", [f"SAI{m:02d}" for m in range(1, 13)]))
        self.veg_f.update(table("DVEG", VEG_DVEG))
        phys = table("PHYS", ["NROOT"] + VEG_PHYS, int_cols=(0,))
        self.veg_i["NROOT"] = phys.pop("NROOT")
        self.veg_f.update(phys)
        photo = table("PHOTO", ["C3C4"] + VEG_PHOTO, int_cols=(0,))
        self.veg_i["C3C4"] = photo.pop("C3C4")
        self.veg_f.update(photo)
        # VOC#tag (SLAREA, EPS) feeds only bvocflux, which the reference never calls (func:6749)

    # ---- JSON round trip (packed copy shipped in noahmp_b200/data/) ------------------------------
    def to_json(self) -> str:
        def clean(d):
            return {k: ([None if (isinstance(x, float) and np.isnan(x)) else x for x in v]
                        if isinstance(v, list) else v) for k, v in d.items()}
        return json.dumps({"soil_tag": self.soil_tag, "veg_tag": self.veg_tag, "gen": clean(self.gen),
                           "soil": clean(self.soil), "color": clean(self.color), "veg_f": clean(self.veg_f),
                           "veg_i": self.veg_i, "flags": self.flags}, indent=1)

    @classmethod
    def from_json(cls, path: str) -> "ParamTables":
        with open(path) as fh:
            d = json.load(fh)

        def back(dd):
            return {k: ([float("nan") if x is None else x for x in v] if isinstance(v, list) else v)
                    for k, v in dd.items()}
        return cls(gen=back(d["gen"]), soil=back(d["soil"]), color=back(d["color"]), veg_f=back(d["veg_f"]),
                   veg_i=d["veg_i"], flags=d["flags"], soil_tag=d["soil_tag"], veg_tag=d["veg_tag"])

    @classmethod
    def default(cls, soil_tag: str = "STAS", veg_tag: str = "USGS") -> "ParamTables":
        """Packed copy of the reference's tbl/ values (generated by tools/pack_reference_tables.py)."""
        here = os.path.dirname(os.path.abspath(__file__))
        return cls.from_json(os.path.join(here, "data", f"tables_{soil_tag}_{veg_tag}.json"))

    # ---- sizes ----------------------------------------------------------------------------------
    @property
    def nsoiltyp(self) -> int:
        return len(self.soil["BEXP"])

    @property
    def ncolor(self) -> int:
        return len(self.color["ALBSAT_VIS"])

    @property
    def nvegtyp(self) -> int:
        return len(self.veg_f["XL"])

    # ---- packing for the C ABI --------------------------------------------------------------------
    def pack(self) -> dict:
        f32, i32 = np.float32, np.int32
        gen = np.zeros(NGEN, f32)
        g = self.gen
        for i, nm in enumerate(GEN_SCALARS):
            gen[i] = g[nm]
        base = len(GEN_SCALARS)
        gen[base + 0:base + 2] = g["ALBICE"]
        gen[base + 2:base + 4] = g["ALBLAKE"]
        gen[base + 4:base + 6] = g["OMEGAS"]
        gen[base + 6] = g["BETADS"]
        gen[base + 7] = g["BETAIS"]
        gen[base + 8] = g["EMSSOIL"]
        gen[base + 9] = g["EMSLAKE"]
        gen[base + 10] = len(g["SLOPE"])
        slope = np.full(MSLOPETYP, np.nan, f32)
        slope[:len(g["SLOPE"])] = g["SLOPE"]
        gen[base + 11:] = slope
        soil = np.stack([np.asarray(self.soil[nm], f32) for nm in SOIL_SLOTS], axis=1)
        color = np.stack([np.asarray(self.color[nm], f32) for nm in COLOR_SLOTS], axis=1)
        veg_f = np.stack([np.asarray(self.veg_f[nm], f32) for nm in VEG_F_SLOTS], axis=1)
        veg_i = np.stack([np.asarray(self.veg_i[nm], i32) for nm in VEG_I_SLOTS], axis=1)
        flags = np.asarray([self.flags[nm] for nm in VEG_FLAGS], i32)
        return {"gen": np.ascontiguousarray(gen), "soil": np.ascontiguousarray(soil),
                "color": np.ascontiguousarray(color), "veg_f": np.ascontiguousarray(veg_f),
                "veg_i": np.ascontiguousarray(veg_i), "flags": flags}
