"""TBL reader/packer vs the reference's table semantics (utils:56-97,200-237; gen/soil/veg readers)."""
import os
import textwrap

import numpy as np
import pytest

from conftest import REFERENCE
from noahmp_b200 import tables as T


def _write(tmp_path, name, text):
    p = tmp_path / name
    p.write_text(textwrap.dedent(text))
    return str(p)


def test_section_lookup_rules(tmp_path):
    path = _write(tmp_path, "X.TBL", """\
        &A#T1
        1.5
        &A
        2.5
        &AB
        3.5
        &A#T2
        4.5, 5.5
        """)
    assert T.read_var(path, "A") == 2.5              # untagged header must have no '#'
    assert T.read_var(path, "A", "T1") == 1.5        # tagged header must match NAME and TAG exactly
    assert T.read_var(path, "A", "T2", n=2) == [4.5, 5.5]
    assert T.read_var(path, "AB") == 3.5             # '&A' does not match '&AB'
    with pytest.raises(T.TableError, match="Unable to find A#T3"):
        T.read_var(path, "A", "T3")
    with pytest.raises(T.TableError, match="Unable to open file"):
        T.read_var(str(tmp_path / "missing.TBL"), "A")


def test_counted_table_and_trailing_text(tmp_path):
    path = _write(tmp_path, "S.TBL", """\
        &PARM#STAS
        2      BB     MAXSMC  header text is ignored
        1,     2.79,   0.339,   0.236,   0.010,   0.069,  4.66E-5,  0.608E-6,  0.92, 'SAND'
        2,     4.26,   0.421,   0.383,   0.028,   0.036,  1.41D-5,  0.514E-5,  0.82, 'LOAMY SAND'
        """)
    rows = T.read_table(path, "PARM", "STAS", 30)
    assert len(rows) == 2 and rows[1][0] == "2"
    vals = T._row_values(rows[1], 8, "PARM.STAS", path)
    assert vals[0] == 4.26 and vals[5] == 1.41e-5   # Fortran D exponent
    with pytest.raises(T.TableError, match="too big"):
        T.read_table(path, "PARM", "STAS", 1)


def test_default_tables_shapes_and_values(tables):
    p = tables.pack()
    assert p["gen"].shape == (T.NGEN,) and p["gen"].dtype == np.float32
    assert p["soil"].shape == (19, 10) and p["color"].shape == (9, 4)
    assert p["veg_f"].shape == (27, T.NVEGF) and p["veg_i"].shape == (27, 2)
    assert list(p["flags"]) == [1, 16, 19, 24, 13]              # USGS ISURBAN ISWATER ISBARREN ISICE ISEGBLF
    assert p["gen"][T.GEN_SLOTS.index("ZBOT")] == np.float32(-8.0)
    assert p["gen"][T.GEN_SLOTS.index("NSLOPE")] == 9
    loam = p["soil"][5]
    assert loam[T.SOIL_SLOTS.index("SMCMAX")] == np.float32(0.439)
    grass = p["veg_f"][6]
    assert grass[T.VEG_F_SLOTS.index("LAI07")] == np.float32(3.5)
    assert p["veg_i"][6].tolist() == [3, 1]


def test_derived_soil_fields_reproduce_reference_arithmetic(tables):
    """soil_param:58-61 incl. the integer literal `0468` (=468, not 0.468) and the WHERE mask."""
    f32 = np.float32
    g, s = tables.gen, tables.soil
    kdt = f32(g["KDTREF"]) * np.asarray(s["DKSAT"], f32) / f32(g["DKREF"])
    assert np.array_equal(np.asarray(s["KDT"], f32), kdt)
    smcref = np.asarray(s["SMCREF"], f32)
    frzx = np.asarray(s["FRZX"], f32)
    expect = f32(g["FRZK"]) * (np.asarray(s["SMCMAX"], f32)[smcref > 0] / smcref[smcref > 0]) * (f32(0.412) / f32(468))
    assert np.array_equal(frzx[smcref > 0], expect)
    assert np.isnan(frzx[smcref <= 0]).all()                    # type 14 (water) keeps nan4
    assert frzx[5] < 1e-3                                       # /468: three orders below the "intended" /0.468


def test_json_round_trip(tables, tmp_path):
    p = tmp_path / "t.json"
    p.write_text(tables.to_json())
    again = T.ParamTables.from_json(str(p)).pack()
    for k, v in tables.pack().items():
        assert np.array_equal(v, again[k], equal_nan=True), k


@pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="reference tree only exists in the build container")
@pytest.mark.parametrize("soil_tag,veg_tag", [("STAS", "USGS"), ("STAS-RUC", "MODIFIED_IGBP_MODIS_NOAH")])
def test_shipped_json_matches_reference_tbl(soil_tag, veg_tag):
    ref = T.ParamTables.from_tbl_dir(os.path.join(REFERENCE, "tbl"), soil_tag, veg_tag).pack()
    got = T.ParamTables.default(soil_tag, veg_tag).pack()
    for k in ref:
        assert np.array_equal(ref[k], got[k], equal_nan=True), k


# ---- the readers inside the pin: the reference's readptable routines executed from their source text ---------------------
COMBOS = [("STAS", "USGS"), ("STAS", "MODIFIED_IGBP_MODIS_NOAH"), ("STAS-RUC", "USGS"), ("STAS-RUC", "MODIFIED_IGBP_MODIS_NOAH")]


def _module_arrays_of(t: T.ParamTables) -> dict:
    """The product parser's values under the names of the reference's module variables (Fortran index order)."""
    f32 = np.float32
    out = {}
    g = t.gen
    for nm in ("CSOIL", "ZBOT", "CZIL", "DKREF", "KDTREF", "FRZK", "TIMEAN", "FSATMAX", "MLTFCT", "Z0SNO", "SSI", "SWEMAX", "BETADS",
               "BETAIS", "EMSSOIL", "EMSLAKE"):
        out["KK_" + nm] = f32(g[nm])
    for nm in ("ALBICE", "ALBLAKE", "OMEGAS"):
        out["KK_" + nm] = np.asarray(g[nm], f32)
    out["LK_SLOPE"], out["NSLPTYP"] = np.asarray(g["SLOPE"], f32), len(g["SLOPE"])
    for nm in T.SOIL_SLOTS:
        out["LK_" + nm] = np.asarray(t.soil[nm], f32)
    out["NSLTYP"], out["NSOILCOL"], out["NLUTYP"] = t.nsoiltyp, t.ncolor, t.nvegtyp
    for nm in ("ALBSAT", "ALBDRY"):
        out["LK_" + nm] = np.stack([np.asarray(t.color[nm + "_VIS"], f32), np.asarray(t.color[nm + "_NIR"], f32)])
    v = t.veg_f
    for nm in ("RHOL", "RHOS", "TAUL", "TAUS"):
        out["LK_" + nm] = np.stack([np.asarray(v[nm + "_VIS"], f32), np.asarray(v[nm + "_NIR"], f32)])
    for nm in ("LAI", "SAI"):
        out[f"LK_{nm}12M"] = np.stack([np.asarray(v[f"{nm}{m:02d}"], f32) for m in range(1, 13)])
    for nm in ["XL"] + T.VEG_PHYS + T.VEG_DVEG + T.VEG_PHOTO:
        out["LK_" + nm] = np.asarray(v[nm], f32)
    for nm in ("NROOT", "C3C4"):
        out["LK_" + nm] = np.asarray(t.veg_i[nm], np.int32)
    for nm in T.VEG_FLAGS:
        out[nm] = int(t.flags[nm])
    return out


def _same_as_reference(ours: dict, ref: dict):
    """Every array the product parser holds equals the reference routine's, on the rows the table has; beyond them the
    reference's arrays keep their initial value (nan4 / 0)."""
    for nm, a in ours.items():
        r = np.asarray(ref[nm])
        a = np.asarray(a)
        if a.ndim == 0:
            assert np.array_equal(a, r, equal_nan=True), nm
            continue
        head = r[..., :a.shape[-1]]
        assert np.array_equal(a, head, equal_nan=True), nm
        rest = r[..., a.shape[-1]:]
        assert rest.size == 0 or (np.isnan(rest).all() if rest.dtype.kind == "f" else (rest == 0).all()), nm
    # nothing the physics reads is missing from the product side (VOC and description tables feed only dead code)
    unread = {k for k in ref if k.startswith(("LK_", "KK_")) and k not in ours} - {"LK_SLAREA", "LK_EPS"}
    assert not unread, unread


@pytest.mark.parametrize("soil_tag,veg_tag", COMBOS)
def test_all_module_arrays_equal_the_committed_reader_output(soil_tag, veg_tag):
    """tests/golden/tables_reference.npz (tools/make_table_fixture.py): what the reference's `*_readptable` routines,
    executed from their Fortran text, leave in every LK_/KK_ array - against the shipped packed tables."""
    from conftest import ROOT
    g = np.load(os.path.join(ROOT, "tests", "golden", "tables_reference.npz"))
    ref = {k.split("|", 2)[2]: g[k] for k in g.files if k.startswith(f"{soil_tag}|{veg_tag}|")}
    assert len(ref) >= 85
    _same_as_reference(_module_arrays_of(T.ParamTables.default(soil_tag, veg_tag)), ref)


@pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="reference tree only exists in the build container")
@pytest.mark.parametrize("soil_tag,veg_tag", COMBOS)
def test_product_parser_equals_the_interpreted_reference_readers(soil_tag, veg_tag):
    """Live: oracle/f90_tables.py runs gen -> soil -> veg readptable from /root/reference/core on /root/reference/tbl."""
    from oracle import f90_tables
    ref = f90_tables.read_all(os.path.join(REFERENCE, "tbl"), soil_tag, veg_tag)
    _same_as_reference(_module_arrays_of(T.ParamTables.from_tbl_dir(os.path.join(REFERENCE, "tbl"), soil_tag, veg_tag)), ref)


MALFORMED = {
    # name: (file to damage, regex substitution, expected message fragment of the reference's assert)
    "missing tag": ("SOILPARMMP.TBL", (r"&PARM#STAS\b(?!-)", "&PARM#OTHER"), "Unable to find PARM#STAS in file SOILPARMMP.TBL"),
    "row count above the array": ("SOILPARMMP.TBL", (r"(&PARM#STAS\s*\n\s*)19", r"\g<1>31"), "Table PARM#STAS in file SOILPARMMP.TBL is too big to be read"),
    "row count not a number": ("GENPARMMP.TBL", (r"(&SLOPE\s*\n\s*)9", r"\g<1>nine"), "Unable to read the number of rows for table SLOPE in file GENPARMMP.TBL"),
    "table cut short": ("VEGPARMMP.TBL", (r"(?s)(&VOC#MODIFIED_IGBP_MODIS_NOAH.*?\n(?:.*\n){5}).*", r"\g<1>"), "Unable to read table VOC#MODIFIED_IGBP_MODIS_NOAH from file VEGPARMMP.TBL"),
    "text where a number belongs": ("SOILPARMMP.TBL", (r"(&PARM#STAS\s*\n.*\n\s*1,\s*)2\.79", r"\g<1>abc"), "Unable to read PARM.STAS from file SOILPARMMP.TBL"),
    "scalar missing": ("GENPARMMP.TBL", (r"&CZIL\b", "&CZIL_GONE"), "Unable to find CZIL in file GENPARMMP.TBL"),
    "flag not an integer": ("VEGPARMMP.TBL", (r"(&ISURBAN#USGS\s*\n\s*)1", r"\g<1>1.5"), "Unable to read table ISURBAN#USGS from file VEGPARMMP.TBL"),
}


@pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="reference tree only exists in the build container")
@pytest.mark.parametrize("case", list(MALFORMED))
def test_malformed_tables_raise_the_reference_messages(tmp_path, case):
    """Both readers get the same damaged file; the product raises TableError where the reference's routine would
    `assert` and `stop` (utils:43-53), with the reference's message."""
    import re
    import shutil
    from oracle import f90_tables
    fname, (pat, rep), msg = MALFORMED[case]
    for f in (T.GEN_FILE, T.SOIL_FILE, T.VEG_FILE):
        shutil.copy(os.path.join(REFERENCE, "tbl", f), tmp_path / f)
    text = (tmp_path / fname).read_text()
    damaged, n = re.subn(pat, rep, text, count=1)
    assert n == 1, "the damage pattern must hit"
    (tmp_path / fname).write_text(damaged)
    veg_tag = "MODIFIED_IGBP_MODIS_NOAH" if "MODIS" in msg else "USGS"
    with pytest.raises(f90_tables.FortranStop) as ref:
        f90_tables.read_all(str(tmp_path), "STAS", veg_tag)
    assert msg in str(ref.value)
    if "VOC#" in msg:
        return      # the product does not read the VOC table (it feeds only bvocflux, dead code at func:6749)
    with pytest.raises(T.TableError) as got:
        T.ParamTables.from_tbl_dir(str(tmp_path), "STAS", veg_tag)
    ours = str(got.value).replace(str(tmp_path) + os.sep, "")
    assert ours == str(ref.value), (ours, str(ref.value))


@pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="reference tree only exists in the build container")
def test_list_directed_forms_both_readers_accept(tmp_path):
    """D exponents, quoted strings with commas, blanks instead of commas, a repeat count and trailing text."""
    import shutil
    from oracle import f90_tables
    for f in (T.GEN_FILE, T.SOIL_FILE, T.VEG_FILE):
        shutil.copy(os.path.join(REFERENCE, "tbl", f), tmp_path / f)
    p = tmp_path / T.SOIL_FILE
    text = p.read_text().splitlines()
    i = next(k for k, ln in enumerate(text) if ln.strip().startswith("&PARM#STAS") and "RUC" not in ln)
    row = text[i + 2]
    vals = [v.strip() for v in row.split(",")]
    text[i + 2] = f"{vals[0]}  {vals[1]} {vals[2]}, {vals[3]} ,{vals[4]}  0.069D0 4.66D-5 , 0.608E-6 {vals[8]} 'SAND, coarse' extra"
    p.write_text("\n".join(text) + "\n")
    ref = f90_tables.read_all(str(tmp_path), "STAS", "USGS")
    _same_as_reference(_module_arrays_of(T.ParamTables.from_tbl_dir(str(tmp_path), "STAS", "USGS")), ref)
    assert ref["LK_DKSAT"][0] == np.float32(4.66e-5) and ref["LK_PSISAT"][0] == np.float32(0.069)
