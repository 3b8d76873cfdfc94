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
