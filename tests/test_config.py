"""Config / namelist surface: same attributes and exit behaviour as offline/noahmp_config.py of the reference."""
import datetime
import os

import pytest

from conftest import REFERENCE, ROOT
from noahmp_b200.compat import f90nml


def test_namelist_reader_types():
    nml = f90nml.reads("""
        ! comment
        &grp
          a = 1, b = 2.5e0, c = 'it''s', d = .true.   ! trailing comment
          e = 1, 2, 3
          F = "x ! not a comment"
        /
        &Other  x = -3 /
    """)
    g = nml["GRP"]
    assert g["a"] == 1 and g["b"] == 2.5 and g["c"] == "it's" and g["d"] is True and g["e"] == [1, 2, 3]
    assert g["f"] == "x ! not a comment" and "A" in g and "zz" not in g
    assert nml["other"]["X"] == -3


def test_config_attributes_match_reference_contract():
    import noahmp_config
    cfg = noahmp_config.Config(os.path.join(ROOT, "run", "case.nml"))
    assert cfg.constfile == "geo_em.d01.nc" and cfg.initfile == "init.nc" and cfg.resfile == "restart.nc"
    assert (cfg.indir, cfg.infreq, cfg.outdir, cfg.outfreq) == ("ldasin", "1 hour", "ldasout", "3 hour")
    assert (cfg.resdir, cfg.resfreq) == ("restart", "1 month")
    assert cfg.timestep == datetime.timedelta(seconds=900)
    assert cfg.begdatetime == datetime.datetime(2000, 1, 1) and cfg.enddatetime == datetime.datetime(2000, 1, 2)
    assert [getattr(cfg, k) for k in ("opt_veg", "opt_run", "opt_btr", "opt_rad", "opt_tub", "opt_can", "opt_inf",
                                      "opt_snf", "opt_tbot")] == [1] * 9
    # namelist names -> noahmp_set_options order (global:77-80); absent opt_frz/alb/stc take 1/2/1
    assert cfg.physics_options() == [1, 1, 1, 1, 1, 1, 1, 1, 2, 1, 1, 1]


def test_missing_file_and_missing_field_exit_1(tmp_path, capsys):
    import noahmp_config
    with pytest.raises(SystemExit) as e:
        noahmp_config.Config(str(tmp_path / "nope.nml"))
    assert e.value.code == 1 and "ERR: Unable to find configuration file" in capsys.readouterr().out
    text = open(os.path.join(ROOT, "run", "case.nml")).read().replace("interval_seconds = 900,", "")
    bad = tmp_path / "bad.nml"
    bad.write_text(text)
    with pytest.raises(SystemExit) as e:
        noahmp_config.Config(str(bad))
    assert e.value.code == 1 and "Unable to find interval_seconds" in capsys.readouterr().out


@pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="reference tree only exists in the build container")
def test_reference_namelist_parses_identically():
    import noahmp_config
    ref = noahmp_config.Config(os.path.join(REFERENCE, "run", "case.nml"))
    ours = noahmp_config.Config(os.path.join(ROOT, "run", "case.nml"))
    for k in ("constfile", "initfile", "resfile", "indir", "infreq", "outdir", "outfreq", "resdir", "resfreq", "timestep",
              "begdatetime", "enddatetime", "opt_veg", "opt_run", "opt_btr", "opt_rad", "opt_tub", "opt_can", "opt_inf",
              "opt_snf", "opt_tbot"):
        assert getattr(ref, k) == getattr(ours, k), k


def test_main_entry_point_signature():
    import importlib.util
    spec = importlib.util.spec_from_file_location("noahmp_main", os.path.join(ROOT, "run", "main.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    assert mod.DEFAULT_NAMELIST_FILE == "case.nml" and callable(mod.main)
