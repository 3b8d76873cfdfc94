"""GPU tests of the widened rows: forcing preparation kernel parity, file-driven run, restart continuity."""
import datetime as dt
import os

import numpy as np
import pytest
import torch

from noahmp_b200 import forcing as F
from noahmp_b200 import layout as L, synth
from noahmp_b200.engine import NoahMP
from oracle.forcing_prep import prepare_forcing

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("ncol", [1, 3, 4, 1000, 100_003, 262_144])
def test_prepare_forcing_bit_exact(tables, ncol):
    rng = np.random.default_rng(ncol)
    a = rng.normal(280.0, 30.0, size=(8, ncol)).astype(np.float32)
    b = rng.normal(280.0, 30.0, size=(8, ncol)).astype(np.float32)
    geo = F.geo_planes(rng.uniform(-1.5, 1.5, ncol), rng.uniform(-3.1, 3.1, ncol))
    eng = NoahMP(tables)
    out = torch.empty((9, ncol), dtype=torch.float32, device=DEV)
    for w, when in ((0.0, dt.datetime(2000, 1, 1)), (0.375, dt.datetime(2000, 7, 4, 17, 45)), (1.0, dt.datetime(2001, 12, 31, 23))):
        s = F.solar_scalars(when)
        eng.prepare_forcing(torch.from_numpy(a).to(DEV), torch.from_numpy(b).to(DEV), w, torch.from_numpy(geo).to(DEV), s, out)
        torch.cuda.synchronize()
        ref = prepare_forcing(a, b, w, geo, s)
        assert np.array_equal(out.cpu().numpy().view(np.int32), ref.view(np.int32)), (ncol, w)
    eng.close()


def _make_case(tmp_path, ncol=96, hours=7, opt_veg=1):
    """Write an init file, hourly forcing records and a namelist; return the Config."""
    import noahmp_config
    w = synth.make_workload("global025", ncol=ncol, nsteps=1)
    rng = np.random.default_rng(3)
    lat = np.degrees(w.static_f[L.C["LAT"]].astype(np.float64))
    init = dict(lat=lat, lon=rng.uniform(-120, 40, ncol), lutyp=w.static_i & 0xff, sltyp=(w.static_i >> 8) & 0xff,
                slptyp=(w.static_i >> 16) & 0xff, isc=(w.static_i >> 24) & 0xf, shdfac=w.static_f[L.C["SHDFAC"]],
                tbot=w.static_f[L.C["TBOT"]], tsk=w.state[L.S["TG"]], stc=np.stack([w.state[L.stc(k)] for k in range(1, 5)]),
                smc=np.stack([w.state[L.smc(k)] for k in range(1, 5)]), snowh=np.where(lat > 55, 0.3, 0.0).astype(np.float32),
                sneqv=np.where(lat > 55, 75.0, 0.0).astype(np.float32))
    os.makedirs(tmp_path / "ldasin")
    np.savez(tmp_path / "init.npz", **init)
    t0 = dt.datetime(2000, 6, 30, 22)
    for h in range(hours + 1):
        t = t0 + dt.timedelta(hours=h)
        rec = synth.forcing_step(w, h)[:8].copy()
        np.save(tmp_path / "ldasin" / (F.stamp(t) + ".npy"), rec)
    nml = f"""&NOAHMP_OFFLINE
      static_parameter_file = 'none', initialization_file = '{tmp_path}/init.npz', restart_file = 'none.npz',
      input_directory = '{tmp_path}/ldasin', input_frequency = '1 hour', output_directory = '{tmp_path}/out',
      output_frequency = '3 hour', restart_directory = '{tmp_path}/restart', restart_frequency = '2 hour',
      start_year = 2000, start_month = 6, start_day = 30, start_hour = 22, start_minute = 0, start_second = 0,
      end_year = 2000, end_month = 7, end_day = 1, end_hour = {22 + hours - 24}, end_minute = 0, end_second = 0,
      interval_seconds = 900,
      opt_veg = {opt_veg}, opt_run = 1, opt_btr = 1, opt_rad = 1, opt_tub = 1, opt_can = 1, opt_inf = 1, opt_tbot = 1, opt_snf = 1
    /"""
    (tmp_path / "case.nml").write_text(nml)
    return noahmp_config.Config(str(tmp_path / "case.nml"))


def test_file_driven_run_outputs_and_restart_continuity(tmp_path, tables):
    from noahmp_b200.driver import read_restart, run_files
    cfg = _make_case(tmp_path)
    full = run_files(cfg, variant="parity")
    assert full["steps"] == 28 and full["flagged_column_launches"] == 0 and full["columns"] == 96
    names = sorted(os.path.basename(p) for p in full["files_written"])
    assert names == ["200007010100.LDASOUT.npz", "200007010400.LDASOUT.npz", "RESTART.200007010000.npz",
                     "RESTART.200007010200.npz", "RESTART.200007010400.npz"]
    end_state, when, step = read_restart(str(tmp_path / "restart" / "RESTART.200007010400.npz"))
    assert when == dt.datetime(2000, 7, 1, 4) and step == 24
    out = np.load(tmp_path / "out" / "200007010400.LDASOUT.npz")
    assert list(out["planes"]) == L.DIAG and np.isfinite(out["diag"][L.D["TRAD"]]).all()
    # resume from the 02:00 restart: the remaining steps must reproduce the uninterrupted run bit for bit
    cfg.resfile = str(tmp_path / "restart" / "RESTART.200007010200.npz")
    resumed = run_files(cfg, variant="parity")
    assert resumed["steps"] == 28
    again, _, _ = read_restart(str(tmp_path / "restart" / "RESTART.200007010400.npz"))
    assert np.array_equal(again.view(np.int32), end_state.view(np.int32))
    # the snowy columns really carry layers and the forcing really varies in time
    isn = end_state[L.S["ISNOW"]].view(np.int32)
    assert (isn < 0).any() and (isn == 0).any()


def test_netcdf_driven_run_equals_the_npz_driven_run(tmp_path, tables):
    """The same case with a WRF-style netCDF-3 initialisation file and LDASIN records: identical state, netCDF outputs."""
    from scipy.io import netcdf_file
    from test_driver import _write_wrf_style_init
    from noahmp_b200 import ncio
    from noahmp_b200.driver import read_restart, run_files
    (tmp_path / "a").mkdir()
    (tmp_path / "b").mkdir()
    cfg = _make_case(tmp_path / "a", hours=4)
    init = dict(np.load(tmp_path / "a" / "init.npz"))
    # WRF files carry the vegetation fraction as a percentage: use fractions k/128, which survive x100 /100 in fp32 exactly
    init["shdfac"] = (np.round(init["shdfac"] * 128.0) / 128.0).astype(np.float32)
    init["lat"], init["lon"] = init["lat"].astype(np.float32), init["lon"].astype(np.float32)   # XLAT / XLONG are float32
    np.savez(tmp_path / "a" / "init.npz", **init)
    ref = run_files(cfg, variant="parity")
    ref_state, _, _ = read_restart(str(tmp_path / "a" / "restart" / "RESTART.200007010200.npz"))
    # the netCDF twin: init on an 8 x 12 grid, records as <YYYYMMDDHH>.LDASIN_DOMAIN1
    _write_wrf_style_init(str(tmp_path / "b" / "init.nc"), init, (8, 12))
    os.makedirs(tmp_path / "b" / "ldasin")
    for f in sorted(os.listdir(tmp_path / "a" / "ldasin")):
        rec = np.load(tmp_path / "a" / "ldasin" / f)
        with netcdf_file(str(tmp_path / "b" / "ldasin" / (f[:10] + ".LDASIN_DOMAIN1")), "w", version=2) as nc:
            nc.createDimension("Time", 1)
            nc.createDimension("south_north", 8)
            nc.createDimension("west_east", 12)
            for i, name in enumerate(("T2D", "PSFC", "U2D", "V2D", "Q2D", "SWDOWN", "LWDOWN", "RAINRATE")):
                v = nc.createVariable(name, "f4", ("Time", "south_north", "west_east"))
                v[:] = rec[i].reshape(1, 8, 12)
    cfg.initfile = str(tmp_path / "b" / "init.nc")
    cfg.indir, cfg.outdir, cfg.resdir = (str(tmp_path / "b" / d) for d in ("ldasin", "out", "restart"))
    got = run_files(cfg, variant="parity")
    assert got["steps"] == ref["steps"] == 16 and got["flagged_column_launches"] == 0
    names = sorted(os.path.basename(p) for p in got["files_written"])
    assert names == ["200007010100.LDASOUT_DOMAIN1", "RESTART.200007010000_DOMAIN1", "RESTART.200007010200_DOMAIN1"]
    st, when, step = read_restart(str(tmp_path / "b" / "restart" / "RESTART.200007010200_DOMAIN1"))
    assert when == dt.datetime(2000, 7, 1, 2) and step == 16
    assert np.array_equal(st.view(np.int32), ref_state.view(np.int32))
    out, _, _ = ncio.read_planes(str(tmp_path / "b" / "out" / "200007010100.LDASOUT_DOMAIN1"), L.DIAG, int_planes=("ITER", "FLAGS"))
    assert np.isfinite(out[L.D["TRAD"]]).all()


def test_dynamic_vegetation_run_starts_with_a_canopy(tmp_path, tables):
    """opt_veg = 2 reads LAI/SAI and the carbon pools from the state: the driver initialises them from the tables."""
    from noahmp_b200.driver import columns_from_init, read_restart, run_files
    cfg = _make_case(tmp_path, hours=4, opt_veg=2)
    out = run_files(cfg, variant="parity")
    assert out["steps"] == 16 and out["flagged_column_launches"] == 0
    st, _, _ = read_restart(str(tmp_path / "restart" / "RESTART.200007010200.npz"))
    c0 = columns_from_init(cfg.initfile, tables, cfg.begdatetime)
    veg = c0.state[L.S["LFMASS"]] > 0
    assert veg.sum() > 10 and (~veg).sum() > 0                          # vegetated and barren/ice/water columns both present
    assert (st[L.S["LAI"]][veg] >= 0.05).all() and (st[L.S["LAI"]][~veg] == 0).all()
    assert (st[L.S["LFMASS"]][veg] != c0.state[L.S["LFMASS"]][veg]).mean() > 0.7     # the pools evolve; some sit at the minimum leaf mass (func:6642-7025)
    dg = np.load(tmp_path / "out" / "200007010100.LDASOUT.npz")["diag"]
    assert (dg[L.D["FVEG"]][veg] > 0.01).mean() > 0.8            # snow-buried short vegetation has no canopy tile
    assert np.isfinite(dg[L.D["NPP"]]).all()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_rank_file_driven_run(tmp_path, tables):
    """torchrun, two ranks: each GPU advances its block of columns; gathered outputs and restarts equal the one-GPU run."""
    import subprocess
    import sys
    from conftest import ROOT
    from noahmp_b200.driver import read_restart, run_files
    (tmp_path / "one").mkdir()
    cfg = _make_case(tmp_path / "one", ncol=101, hours=4)                # uneven shards: 51 + 50 columns
    ref = run_files(cfg, variant="parity")
    ref_state, _, _ = read_restart(str(tmp_path / "one" / "restart" / "RESTART.200007010200.npz"))
    ref_diag = np.load(tmp_path / "one" / "out" / "200007010100.LDASOUT.npz")["diag"]
    # same inputs, separate output directories
    nml = (tmp_path / "one" / "case.nml").read_text().replace(f"{tmp_path}/one/out", f"{tmp_path}/two/out") \
        .replace(f"{tmp_path}/one/restart", f"{tmp_path}/two/restart")
    (tmp_path / "two").mkdir()
    (tmp_path / "two" / "case.nml").write_text(nml)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tests", "_multi_gpu_driver.py"),
                        str(tmp_path / "two" / "case.nml")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    assert f"RANKS 2 STEPS {ref['steps']} FLAGGED 0 COLUMNS 101" in r.stdout
    st, when, step = read_restart(str(tmp_path / "two" / "restart" / "RESTART.200007010200.npz"))
    assert step == 16 and np.array_equal(st.view(np.int32), ref_state.view(np.int32))
    dg = np.load(tmp_path / "two" / "out" / "200007010100.LDASOUT.npz")["diag"]
    assert np.array_equal(dg.view(np.int32), ref_diag.view(np.int32))


def test_main_entry_point_runs_the_synthetic_single_point(tmp_path):
    import importlib.util
    from conftest import ROOT
    spec = importlib.util.spec_from_file_location("noahmp_main", os.path.join(ROOT, "run", "main.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    cwd = os.getcwd()
    os.chdir(tmp_path)           # so the relative 'init.nc' of run/case.nml does not exist -> synthetic mode
    try:
        s = mod.main(os.path.join(ROOT, "run", "case.nml"), verbose=False)
    finally:
        os.chdir(cwd)
    assert s["steps"] == 96 and s["columns"] == 1 and s["flagged_column_launches"] == 0
    assert 250.0 < s["TG"] < 300.0 and s["max_abs_water_error_mm"] < 0.05


# ---- records-driven launches: interpolation + COSZ fused into the step kernel's load --------------------------------------
def _records_case(ncol, nrec=5, steps_per_record=4):
    """nrec hourly records, steps_per_record model steps per record interval (the reference's own namelist: 900 s steps,
    hourly input; run/case.nml:27, offline/noahmp_config.py:81-82)."""
    w = synth.make_workload("global025", ncol=ncol, nsteps=1)
    w.dt = 3600.0 / steps_per_record
    rng = np.random.default_rng(11)
    recs = np.stack([synth.forcing_step(w, h)[:8] for h in range(nrec)]).astype(np.float32)
    lon = rng.uniform(-3.0, 3.0, ncol)
    geo = F.geo_planes(w.static_f[L.C["LAT"]].astype(np.float64), lon)
    t0 = dt.datetime(2000, 6, 30, 22)
    plan = []
    for s in range((nrec - 1) * steps_per_record):
        t = t0 + dt.timedelta(seconds=s * w.dt)
        r, k = divmod(s, steps_per_record)
        plan.append((r, r + 1 if k else r, k / steps_per_record, F.solar_scalars(t)))
    return w, recs, geo, plan, F.julian_of(t0), F.yearlen_of(t0)


@pytest.mark.parametrize("variant", ["strict", "parity", "fast"])
@pytest.mark.parametrize("ncol", [1, 777, 20_000])
def test_records_launch_equals_prepared_forcing(tables, variant, ncol):
    """noahmp_run_records == noahmp_prepare_forcing per step + noahmp_run, bit for bit (state, diagnostics, budget), in every
    build: the fused load performs the preparation kernel's round-to-nearest arithmetic."""
    w, recs, geo, plan, j0, ylen = _records_case(ncol)
    eng = NoahMP(tables, w.opts, device=0, fast=variant)
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(DEV)  # noqa: E731
    drecs, dgeo, sf, si = d(recs), d(geo), d(w.static_f), d(w.static_i)
    n = len(plan)
    forcing = torch.empty((n, L.NFORCING, ncol), dtype=torch.float32, device=DEV)
    for s, (r0, r1, wt, sc) in enumerate(plan):
        eng.prepare_forcing(drecs[r0], drecs[r1], wt, dgeo, sc, forcing[s])
    rp = eng.run_params(w.dt, synth.ZSOIL, j0, ylen)
    sa, da, ba = d(w.state), torch.zeros((L.NDIAG, ncol), dtype=torch.float32, device=DEV), torch.zeros(L.NBUDGET, dtype=torch.float64, device=DEV)
    sb, db, bb = d(w.state), torch.zeros((L.NDIAG, ncol), dtype=torch.float32, device=DEV), torch.zeros(L.NBUDGET, dtype=torch.float64, device=DEV)
    eng.run(rp, sa, forcing, sf, si, da, ba)
    steps = eng.step_table([p[0] for p in plan], [p[1] for p in plan], [p[2] for p in plan], [p[3] for p in plan])
    eng.run_records(rp, sb, drecs, steps, dgeo, sf, si, db, bb)
    torch.cuda.synchronize()
    assert torch.equal(sa.view(torch.int32), sb.view(torch.int32))
    assert torch.equal(da.view(torch.int32), db.view(torch.int32))
    assert torch.allclose(ba, bb, rtol=1e-12, atol=1e-9) and int(ba[L.B["NFLAGGED"]].item()) == 0   # fp64 atomics: addition order
    # the same pair without diagnostics / budget (the benchmark's kernel instantiation; in the fast build the instantiations are
    # separately optimised code and need not agree with each other bit for bit, each agrees with itself across the two modes)
    sc_, sd_ = d(w.state), d(w.state)
    eng.run(rp, sc_, forcing, sf, si)
    eng.run_records(rp, sd_, drecs, steps, dgeo, sf, si)
    torch.cuda.synchronize()
    assert torch.equal(sc_.view(torch.int32), sd_.view(torch.int32))
    if variant != "fast":
        assert torch.equal(sa.view(torch.int32), sc_.view(torch.int32))
    eng.close()


@pytest.mark.parametrize("spl", [1, 3, 16])
def test_streamed_records_equal_resident_records(tables, spl):
    """noahmp_advance_records_host (records from pinned host memory through the device ring, each copied once) ==
    noahmp_run_records on resident records; 40 steps exceed one launch's table (32), the ring wraps several times."""
    w, recs, geo, plan, j0, ylen = _records_case(5000, nrec=11)
    eng = NoahMP(tables, w.opts, device=0, fast="fast")
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(DEV)  # noqa: E731
    dgeo, sf, si = d(geo), d(w.static_f), d(w.static_i)
    rp = eng.run_params(w.dt, synth.ZSOIL, j0, ylen)
    steps = eng.step_table([p[0] for p in plan], [p[1] for p in plan], [p[2] for p in plan], [p[3] for p in plan])
    sa, ba = d(w.state), torch.zeros(L.NBUDGET, dtype=torch.float64, device=DEV)
    eng.run_records(rp, sa, d(recs), steps, dgeo, sf, si, None, ba)
    sb = d(w.state)
    hrec = torch.from_numpy(recs).pin_memory()
    hb = torch.zeros(L.NBUDGET, dtype=torch.float64).pin_memory()
    for _ in range(2):     # twice: the second call re-uses the ring of the first
        sb.copy_(d(w.state))
        nflag = eng.advance_records_host(rp, sb, hrec, steps, dgeo, sf, si, hb, spl)
        assert nflag == 0
        assert torch.equal(sa.view(torch.int32), sb.view(torch.int32))
    bsum, ref = hb.numpy(), ba.cpu().numpy()
    snap = [L.B["NCELL"], L.B["SNEQV"], L.B["SOILWATER"]]
    assert np.array_equal(bsum[snap], ref[snap])
    assert np.allclose(np.delete(bsum, snap), np.delete(ref, snap), rtol=1e-12, atol=1e-9)   # fp64 atomics: order of addition
    eng.close()


@pytest.mark.parametrize("variant", ["fast", "strict"])
@pytest.mark.parametrize("ncol,spl", [(33, 1), (5000, 3), (70_001, 1), (70_001, 16)])
def test_packed_records_end_in_the_same_state(tables, variant, ncol, spl):
    """noahmp_set_record_packing: zero-suppressed record planes (the default candidates SOLDN | PRCP, every plane as a
    candidate, packing off) give the bit-identical state and budget - the packing is lossless - and fewer bytes cross the
    host link.  The records include a night (SOLDN all zero), a dry record, a half-zero VV plane, -0 and a ragged last group;
    with four model steps per record the kernel interpolates between a packed and a dense plane."""
    w, recs, geo, plan, j0, ylen = _records_case(ncol, nrec=7)
    rng = np.random.default_rng(5)
    recs = recs.copy()
    recs[1, 5] = 0.0                                   # night
    recs[2, 5, : ncol // 2] = 0.0                      # terminator
    recs[3, 7] = 0.0                                   # dry everywhere
    recs[4, 7, rng.random(ncol) < 0.5] = 0.0
    recs[:, 3, rng.random(ncol) < 0.5] = 0.0           # calm meridional wind on half of the columns (packed only with mask 0xFF)
    recs[5, 7, 0] = -0.0
    eng = NoahMP(tables, w.opts, device=0, fast=variant)
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(DEV)  # noqa: E731
    dgeo, sf, si = d(geo), d(w.static_f), d(w.static_i)
    rp = eng.run_params(w.dt, synth.ZSOIL, j0, ylen)
    steps = eng.step_table([p[0] for p in plan], [p[1] for p in plan], [p[2] for p in plan], [p[3] for p in plan])
    hrec = torch.from_numpy(recs).pin_memory()
    out = {}
    for name, mask, threads in (("off", 0, 0), ("default", 0xA0, 3), ("all", 0xFF, 2)):
        eng.set_record_packing(mask, threads)
        st = d(w.state)
        hb = torch.zeros(L.NBUDGET, dtype=torch.float64).pin_memory()
        before = eng.h2d_bytes()
        for _ in range(2):      # the second call re-uses ring and staging areas
            st.copy_(d(w.state))
            nflag = eng.advance_records_host(rp, st, hrec, steps, dgeo, sf, si, hb, spl)
        out[name] = (st.cpu().numpy().view(np.int32), hb.numpy().copy(), nflag, (eng.h2d_bytes() - before) // 2)
    for name in ("default", "all"):
        assert np.array_equal(out[name][0], out["off"][0]), name
        assert out[name][2] == out["off"][2]
        snap = [L.B["NCELL"], L.B["SNEQV"], L.B["SOILWATER"]]
        assert np.array_equal(out[name][1][snap], out["off"][1][snap])
        assert np.allclose(out[name][1], out["off"][1], rtol=1e-12, atol=1e-9)
    assert out["off"][3] == 7 * 8 * 4 * ncol                        # every record once, eight dense planes
    assert out["all"][3] < out["default"][3] < out["off"][3]
    with pytest.raises(Exception):
        eng.set_record_packing(0x100, 0)
    eng.close()
