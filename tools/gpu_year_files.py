#!/usr/bin/env python3
"""A file-driven (NetCDF-3) year on the GPU box: BASELINE configs[2] (global 0.25 deg, ~280 000 land columns, hourly steps)
through the namelist driver (`noahmp_b200.driver.run_files`), on 1..8 GPUs under torchrun, with the wall time split into
ingest / kernel / gather / write and the restart written at 1 July resumed to the same end state bit for bit.

TEST / MEASUREMENT INFRASTRUCTURE.  Nothing here is on the product path.

    python tools/gpu_year_files.py generate --dir /tmp/nmp_year [--grid 400x700] [--days 365] [--input-hours 3] [--procs 16]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29577 \
        tools/gpu_year_files.py run --dir /tmp/nmp_year --tag full
    python -m torch.distributed.run ... tools/gpu_year_files.py run --dir /tmp/nmp_year --tag resumed --resume
    python tools/gpu_year_files.py compare --dir /tmp/nmp_year --out gpurun_out/r02_year_files.json

generate: WRF-style `init.nc` on a south_north x west_east grid (columns of `synth.make_workload("global025")`), one
`<YYYYMMDDHH>.LDASIN_DOMAIN1` record every `--input-hours` hours from the synthetic generator (written by a process pool),
and the namelist (monthly LDASOUT, restarts every 6 months, hourly steps, 2001-01-01 .. 2002-01-01).
run: `run_files` with NOAHMP_TIMING=1 (a synchronise after every launch, so that `launch_s` is device time); rank 0 writes
`<dir>/<tag>.json`.  --resume starts from `RESTART.200107010000_DOMAIN1` of the `full` run and writes into `<dir>/resumed/`.
compare: final restarts of the two runs bit for bit, the twelve outputs present, the summary JSON.
"""
from __future__ import annotations

import argparse
import datetime as dt
import json
import os
import shutil
import sys
import time

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "offline"))

T0 = dt.datetime(2001, 1, 1)
VARS = ("T2D", "PSFC", "U2D", "V2D", "Q2D", "SWDOWN", "LWDOWN", "RAINRATE")
_W = None


def _workload(ncol):
    global _W
    if _W is None or _W.ncol != ncol:
        from noahmp_b200 import synth
        _W = synth.make_workload("global025", ncol=ncol, nsteps=1)
    return _W


def _write_record(job):
    """One LDASIN record (process-pool worker): the synthetic generator's step `hour` of the hourly workload."""
    from scipy.io import netcdf_file
    from noahmp_b200 import synth
    path, hour, ny, nx = job
    w = _workload(ny * nx)
    rec = synth.forcing_step(w, hour)[:8]
    with netcdf_file(path, "w", version=2) as nc:
        nc.createDimension("Time", 1)
        nc.createDimension("south_north", ny)
        nc.createDimension("west_east", nx)
        for i, name in enumerate(VARS):
            v = nc.createVariable(name, "f4", ("Time", "south_north", "west_east"))
            v[:] = rec[i].reshape(1, ny, nx)
    return os.path.getsize(path)


def generate(a):
    from multiprocessing import Pool
    from scipy.io import netcdf_file
    from noahmp_b200 import layout as L
    ny, nx = (int(x) for x in a.grid.split("x"))
    ncol = ny * nx
    os.makedirs(os.path.join(a.dir, "ldasin"), exist_ok=True)
    free = shutil.disk_usage(a.dir).free
    nrec = a.days * 24 // a.input_hours + 1
    need = nrec * ncol * 8 * 4
    print(f"generate: {ncol} columns, {nrec} records every {a.input_hours} h, {need / 1e9:.1f} GB (free {free / 1e9:.0f} GB)", flush=True)
    if need > 0.8 * free:
        sys.exit("not enough disk for the forcing records: raise --input-hours")
    w = _workload(ncol)
    lat = np.degrees(w.static_f[L.C["LAT"]].astype(np.float64)).astype(np.float32)
    lon = np.tile(-180.0 + 360.0 * (np.arange(nx) + 0.5) / nx, ny).astype(np.float32)
    two = ("Time", "south_north", "west_east")
    si = w.static_i
    fields = {"XLAT": lat, "XLONG": lon, "IVGTYP": si & 0xff, "ISLTYP": (si >> 8) & 0xff, "SLOPETYP": (si >> 16) & 0xff,
              "ISC": (si >> 24) & 0xf, "TMN": w.static_f[L.C["TBOT"]], "TSK": w.state[L.S["TG"]],
              "SNOWH": w.state[L.S["SNOWH"]], "SNOW": w.state[L.S["SNEQV"]],
              "VEGFRA": w.static_f[L.C["SHDFAC"]] * np.float32(100.0)}
    with netcdf_file(os.path.join(a.dir, "init.nc"), "w", version=2) as nc:
        nc.createDimension("Time", 1)
        nc.createDimension("soil_layers_stag", 4)
        nc.createDimension("south_north", ny)
        nc.createDimension("west_east", nx)
        for name, val in fields.items():
            typ = "i4" if name in ("IVGTYP", "ISLTYP", "SLOPETYP", "ISC") else "f4"
            v = nc.createVariable(name, typ, two)
            v[:] = np.asarray(val).astype(typ).reshape(1, ny, nx)
        for name, fn in (("TSLB", L.stc), ("SMOIS", L.smc), ("SH2O", L.soilwat)):
            v = nc.createVariable(name, "f4", ("Time", "soil_layers_stag", "south_north", "west_east"))
            v[:] = np.stack([w.state[fn(k)] for k in range(1, 5)]).reshape(1, 4, ny, nx)
    jobs = []
    for r in range(nrec):
        t = T0 + dt.timedelta(hours=r * a.input_hours)
        jobs.append((os.path.join(a.dir, "ldasin", t.strftime("%Y%m%d%H") + ".LDASIN_DOMAIN1"), r * a.input_hours, ny, nx))
    tic = time.perf_counter()
    with Pool(a.procs) as pool:
        total = sum(pool.imap_unordered(_write_record, jobs, chunksize=8))
    end = T0 + dt.timedelta(days=a.days)
    nml = f"""&NOAHMP_OFFLINE
      static_parameter_file = 'none', initialization_file = '{a.dir}/init.nc', restart_file = 'none',
      input_directory = '{a.dir}/ldasin', input_frequency = '{a.input_hours} hour', output_directory = '{a.dir}/out',
      output_frequency = '1 month', restart_directory = '{a.dir}/restart', restart_frequency = '6 month',
      start_year = {T0.year}, start_month = 1, start_day = 1, start_hour = 0, start_minute = 0, start_second = 0,
      end_year = {end.year}, end_month = {end.month}, end_day = {end.day}, end_hour = 0, end_minute = 0, end_second = 0,
      interval_seconds = 3600,
      opt_veg = 1, opt_run = 1, opt_btr = 1, opt_rad = 1, opt_tub = 1, opt_can = 1, opt_inf = 1, opt_tbot = 1, opt_snf = 1
    /"""
    open(os.path.join(a.dir, "case.nml"), "w").write(nml)
    meta = {"columns": ncol, "grid": [ny, nx], "records": nrec, "input_hours": a.input_hours, "days": a.days,
            "forcing_bytes": total, "generate_s": time.perf_counter() - tic, "procs": a.procs}
    json.dump(meta, open(os.path.join(a.dir, "generate.json"), "w"))
    print("generated", meta, flush=True)


def run(a):
    import noahmp_config
    from noahmp_b200.driver import run_files
    os.environ["NOAHMP_TIMING"] = "1"
    cfg = noahmp_config.Config(os.path.join(a.dir, "case.nml"))
    if a.resume:
        cfg.resfile = os.path.join(a.dir, "restart", "RESTART.200107010000_DOMAIN1")
        cfg.outdir, cfg.resdir = os.path.join(a.dir, "resumed", "out"), os.path.join(a.dir, "resumed", "restart")
    tic = time.perf_counter()
    out = run_files(cfg, variant=a.variant)
    out["wall_s_with_setup"] = time.perf_counter() - tic
    out["variant"] = a.variant
    if int(os.environ.get("RANK", "0")) == 0:
        out["files_written"] = [os.path.basename(p) for p in out["files_written"]]
        json.dump(out, open(os.path.join(a.dir, a.tag + ".json"), "w"), indent=1)
        print(a.tag, json.dumps({k: out[k] for k in ("ranks", "columns", "steps", "launches", "flagged_column_launches", "timing")}), flush=True)


def compare(a):
    from noahmp_b200 import ncio
    full, res = (json.load(open(os.path.join(a.dir, t + ".json"))) for t in ("full", "resumed"))
    gen = json.load(open(os.path.join(a.dir, "generate.json")))
    end = (T0 + dt.timedelta(days=gen["days"])).strftime("%Y%m%d%H%M")
    sa, ta, na = ncio.read_restart(os.path.join(a.dir, "restart", f"RESTART.{end}_DOMAIN1"))
    sb, tb, nb = ncio.read_restart(os.path.join(a.dir, "resumed", "restart", f"RESTART.{end}_DOMAIN1"))
    same = bool(np.array_equal(sa.view(np.int32), sb.view(np.int32)))
    outs = sorted(f for f in os.listdir(os.path.join(a.dir, "out")) if f.endswith("LDASOUT_DOMAIN1"))
    cs = full["columns"] * full["steps"]
    summary = {"what": "file-driven NetCDF-3 year through driver.run_files (tools/gpu_year_files.py)", "generate": gen,
               "full": full, "resumed": res, "end_restart_time": str(ta), "end_restart_step": [na, nb],
               "resumed_end_state_bit_identical": same, "monthly_outputs": len(outs),
               "cell_steps": cs, "cell_steps_per_s_wall": cs / full["timing"]["total_s"],
               "cell_steps_per_s_kernel": cs / max(full["timing"]["launch_s"], 1e-9)}
    os.makedirs(os.path.dirname(os.path.abspath(a.out)), exist_ok=True)
    json.dump(summary, open(a.out, "w"), indent=1)
    print(json.dumps({k: summary[k] for k in ("resumed_end_state_bit_identical", "monthly_outputs", "cell_steps_per_s_wall",
                                              "cell_steps_per_s_kernel")}), "timing", full["timing"], flush=True)
    if not same or na != nb or len(outs) < min(12, gen["days"] // 31):
        sys.exit("FAILED: resumed run differs from the uninterrupted run or outputs are missing")


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("cmd", choices=["generate", "run", "compare"])
    ap.add_argument("--dir", default="/tmp/nmp_year")
    ap.add_argument("--grid", default="400x700")
    ap.add_argument("--days", type=int, default=365)
    ap.add_argument("--input-hours", type=int, default=3)
    ap.add_argument("--procs", type=int, default=16)
    ap.add_argument("--tag", default="full")
    ap.add_argument("--variant", default="fast")
    ap.add_argument("--resume", action="store_true")
    ap.add_argument("--out", default="gpurun_out/r02_year_files.json")
    args = ap.parse_args()
    {"generate": generate, "run": run, "compare": compare}[args.cmd](args)
