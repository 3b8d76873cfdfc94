"""Time loop around the engine for the namelist-driven single-point case (BASELINE.json configs[0]).

What the reference leaves to a driver it never wrote: the step count from the namelist
period, JULIAN / YEARLEN, cold-start state and forcing.  Forcing here is the synthetic generator of
``synth.py``; a real ingest of ``input_directory`` files is a "next" item.
"""
from __future__ import annotations

import calendar

import numpy as np

from . import layout as L, synth
from .engine import NoahMP
from .tables import ParamTables


def steps_of(cfg) -> int:
    total = (cfg.enddatetime - cfg.begdatetime).total_seconds()
    return int(total // cfg.timestep.total_seconds())


def run_case(cfg, ncol: int = 1, chunk: int = 96, device: int = 0, variant="parity") -> dict:
    import torch
    tables = (ParamTables.from_tbl_dir(cfg.table_directory, cfg.soil_table_tag, cfg.veg_table_tag)
              if getattr(cfg, "table_directory", None) else ParamTables.default(cfg.soil_table_tag, cfg.veg_table_tag))
    opts = cfg.physics_options()
    nsteps = steps_of(cfg)
    w = synth.make_workload("single", ncol=ncol, nsteps=nsteps, tables=tables, opts=opts)
    w.dt = float(cfg.timestep.total_seconds())
    w.yearlen = 366 if calendar.isleap(cfg.begdatetime.year) else 365
    w.julian0 = float(cfg.begdatetime.timetuple().tm_yday - 1) + (cfg.begdatetime.hour * 3600 + cfg.begdatetime.minute * 60
                                                                + cfg.begdatetime.second) / 86400.0
    w.hour0 = cfg.begdatetime.hour + cfg.begdatetime.minute / 60.0
    eng = NoahMP(tables, opts, device=device, fast=variant)
    dev = torch.device("cuda", device)
    state = torch.from_numpy(w.state).to(dev)
    static_f = torch.from_numpy(w.static_f).to(dev)
    static_i = torch.from_numpy(w.static_i).to(dev)
    diag = torch.zeros((L.NDIAG, ncol), dtype=torch.float32, device=dev)
    budget = torch.zeros(L.NBUDGET, dtype=torch.float64, device=dev)
    done = 0
    while done < nsteps:  # forcing is produced chunk by chunk, state stays on the device
        n = min(chunk, nsteps - done)
        forcing = np.stack([synth.forcing_step(w, done + s) for s in range(n)])
        rp = eng.run_params(w.dt, synth.ZSOIL, w.julian0 + done * w.dt / 86400.0, w.yearlen, w.zref, w.co2_frac, w.o2_frac, w.foln)
        eng.run(rp, state, torch.from_numpy(forcing).to(dev), static_f, static_i, diag, budget)
        done += n
    torch.cuda.synchronize()
    st, dg, bd = state.cpu().numpy(), diag.cpu().numpy(), budget.cpu().numpy()
    out = {"columns": ncol, "steps": nsteps, "dt_s": w.dt, "options": opts,
           "flagged_columns": int(bd[L.B["NFLAGGED"]]), "max_abs_water_error_mm": float(bd[L.B["ERRWAT_MAXABS"]])}
    for name in ("TV", "TG", "EAH", "SNEQV", "ZWT"):
        out[name] = float(st[L.S[name]].mean())
    for k in range(1, 5):
        out[f"STC{k}"] = float(st[L.stc(k)].mean())
        out[f"SMC{k}"] = float(st[L.smc(k)].mean())
    for name in ("FSA", "FIRA", "FSH", "SSOIL", "TRAD"):
        out[name] = float(dg[L.D[name]].mean())
    eng.close()
    return out
