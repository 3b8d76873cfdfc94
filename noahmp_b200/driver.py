"""Offline driver around the engine: the time loop, forcing ingest, output and restart cadence the reference's
namelist describes (offline/noahmp_config.py:77-102, run/case.nml) but nothing in the reference implements
(SURVEY 8f ranks 1-3).  Host logic only; the column physics and the forcing preparation run on the GPU.

Two modes:
* file driven  -- ``initialization_file`` is an ``.npz`` of columns or a NetCDF-3 file (see :func:`columns_from_init`,
  ``ncio.py``) and ``input_directory`` holds forcing records (``forcing.py``); outputs and restarts are written in the
  format of the initialisation file (``.npz``, or ``*.LDASOUT_DOMAIN1`` / ``RESTART.*_DOMAIN1`` netCDF on its grid).
* synthetic    -- no initialization file: the single-point workload of ``synth.py`` over the namelist period
  (BASELINE.json configs[0]).
"""
from __future__ import annotations

import datetime as dt
import os
from dataclasses import dataclass

import numpy as np

from . import forcing as F
from . import layout as L, ncio, synth
from .engine import NoahMP
from .tables import ParamTables

f32 = np.float32


@dataclass
class Columns:
    state: np.ndarray      # [NSTATE][ncol]
    static_f: np.ndarray   # [NSTATIC_F][ncol]
    static_i: np.ndarray   # [ncol]
    lat: np.ndarray        # radians
    lon: np.ndarray        # radians
    grid: tuple = ()       # (south_north, west_east) of a netCDF initialisation file; () for flat column lists

    @property
    def ncol(self) -> int:
        return self.state.shape[1]


# ------------------------------------------------------------------------------------------------
def snow_layers_from_depth(snowh, sneqv, tsnow, zsoil=synth.ZSOIL):
    """Initial snow layering from a bulk depth, following common upstream Noah-MP practice (the reference has no
    initialisation code): < 0.025 m no layer; <= 0.05 one; <= 0.10 two halves; <= 0.25 0.05 + rest;
    <= 0.45 0.05 + half + half of the rest; deeper 0.05 + 0.20 + rest.  Ice mass is spread by thickness.
    Returns ISNOW [ncol] and DZ, SNICE [3][ncol] (layers -2..0), ZSNSO [7][ncol]."""
    snowh = np.asarray(snowh, f32)
    sneqv = np.asarray(sneqv, f32)
    n = snowh.size
    dz = np.zeros((3, n), f32)
    isnow = np.zeros(n, np.int32)
    a = (snowh >= 0.025) & (snowh <= 0.05)
    isnow[a] = -1
    dz[2, a] = snowh[a]
    b = (snowh > 0.05) & (snowh <= 0.10)
    isnow[b] = -2
    dz[1, b] = snowh[b] / 2
    dz[2, b] = snowh[b] / 2
    c = (snowh > 0.10) & (snowh <= 0.25)
    isnow[c] = -2
    dz[1, c] = 0.05
    dz[2, c] = snowh[c] - 0.05
    d = (snowh > 0.25) & (snowh <= 0.45)
    isnow[d] = -3
    dz[0, d] = 0.05
    dz[1, d] = 0.5 * (snowh[d] - 0.05)
    dz[2, d] = 0.5 * (snowh[d] - 0.05)
    e = snowh > 0.45
    isnow[e] = -3
    dz[0, e] = 0.05
    dz[1, e] = 0.20
    dz[2, e] = snowh[e] - 0.25
    with np.errstate(divide="ignore", invalid="ignore"):
        snice = np.where(dz > 0, dz * (sneqv / np.maximum(snowh, 1e-6)), 0.0).astype(f32)
    zs = np.zeros((7, n), f32)
    top = -np.cumsum(dz, axis=0)
    for j in range(3):
        zs[j] = np.where(dz[j] > 0, top[j], 0.0)
    base = top[2]
    for k in range(4):
        zs[3 + k] = base + zsoil[k]
    return isnow, dz, snice, zs


def initial_vegetation(tables: ParamTables, lutyp, lat_rad, when: dt.datetime | None):
    """Initial LAI, SAI and carbon pools, following common upstream Noah-MP practice (WRF's NOAHMP_INIT; the reference has
    no initialisation code): LAI / SAI from the monthly tables at the start date (southern hemisphere shifted by six
    months, as phenology does, func:580-597), at least 0.05 / max(0.1 LAI, 0.05); leaf mass LAI x 1000 / max(SLA, 1),
    stem mass SAI x 1000 / 3, root mass and wood 500, stable and fast soil carbon 1000 g/m2; all zero on urban, water,
    barren and ice.  Only the dynamic-vegetation options (opt_veg 2 / 5, func:6642-7025) read the pools."""
    lutyp = np.asarray(lutyp, np.int64)
    month = (when.month if when is not None else 7) - 1
    m = np.where(np.asarray(lat_rad) >= 0.0, month, (month + 6) % 12)
    lai12 = np.stack([np.asarray(tables.veg_f[f"LAI{k:02d}"], f32) for k in range(1, 13)])     # [12][nvegtyp]
    sai12 = np.stack([np.asarray(tables.veg_f[f"SAI{k:02d}"], f32) for k in range(1, 13)])
    lai = np.maximum(lai12[m, lutyp - 1], f32(0.05))
    sai = np.maximum(np.maximum(sai12[m, lutyp - 1], f32(0.1) * lai), f32(0.05))
    sla = np.maximum(np.asarray(tables.veg_f["SLA"], f32)[lutyp - 1], f32(1.0))
    pools = {"LAI": lai, "SAI": sai, "LFMASS": lai * (f32(1000.0) / sla), "STMASS": sai * f32(1000.0 / 3.0),
             "RTMASS": np.full(lai.shape, 500.0, f32), "WOOD": np.full(lai.shape, 500.0, f32),
             "STBLCP": np.full(lai.shape, 1000.0, f32), "FASTCP": np.full(lai.shape, 1000.0, f32)}
    bare = np.isin(lutyp, [tables.flags[k] for k in ("ISURBAN", "ISWATER", "ISBARREN", "ISICE")])
    return {k: np.where(bare, f32(0.0), v).astype(f32) for k, v in pools.items()}


def equilibrium_water_table(tables: ParamTables, sltyp, smc4, sh2o4, timean=None, zsoil=synth.ZSOIL, zwt_max: float = 25.0):
    """Initial water-table depth ZWT [m] and aquifer storage WA = WT [mm] of the SIMGM groundwater scheme (opt_run = 1) at
    which recharge equals discharge for the initial moisture of the bottom soil layer - the state ``groundwater``
    (func:6458-6639) neither fills nor drains.  The reference has no initialisation code (upstream WRF's
    NOAHMP_INIT sets the constants WA = 4900 mm, ZWT = 2.5 m); this is the specification used here, in the reference's own
    terms, for a water table below the soil column (IWT = NSOIL, func:6545-6551):

      QDIS(z) = (1 - FCRMAX) RSBMX exp(-TIMEAN) exp(-FFF (z - 2))                        func:6557 (RSBMX 5, FFF 6)
      QIN(z)  = KA (1 + SMPFZ / (1000 (z - ZNODE4)))                                      func:6566-6570 rearranged
      KA = 1000 DKSAT max(.01, SMC4/SMCMAX)^(2b+3) (1 - FCR4)                             func:6413-6415 (wdfcnd1)
      SMPFZ = max(-120000, CMIC (-PSISAT 1000 S^-b)), S = clip(SMC4/SMCMAX, .01, 1)       func:6560-6563 (CMIC 0.2)
      FCR = max(0, exp(-A (1 - FICE)) - exp(-A)) / (1 - exp(-A)), A = 4, FICE = min(1, ice/SMCMAX)     func:5900-5904

    QIN - QDIS increases with z, so the root on [2, zwt_max] is unique; it is found by 60 bisections in fp64, and taken at
    the interval end when there is no sign change (2.0 m is where WA reaches its cap of 5000 mm, func:6584).  WA follows
    from the scheme's own relation ZWT = (-ZSOIL(4) + 25) - WA / 1000 / ROUS (func:6579), ROUS = 0.2, so that the first call of
    ``groundwater`` does not jump: 1.5 m <= ZWT (func:6609) and WA <= 5000 mm hold by construction."""
    sl = np.asarray(sltyp, np.int64) - 1
    g = lambda k: np.asarray(tables.soil[k], np.float64)[sl]  # noqa: E731
    smcmax, bexp, psisat, dksat = g("SMCMAX"), g("BEXP"), g("PSISAT"), g("DKSAT")
    smc4, sh2o4 = np.asarray(smc4, np.float64), np.asarray(sh2o4, np.float64)
    timean = float(tables.gen["TIMEAN"]) if timean is None else float(timean)
    a = 4.0
    fice = np.minimum(1.0, np.maximum(smc4 - sh2o4, 0.0) / smcmax)
    fcr = np.maximum(0.0, np.exp(-a * (1.0 - fice)) - np.exp(-a)) / (1.0 - np.exp(-a))
    sat = np.clip(smc4 / smcmax, 0.01, 1.0)
    ka = 1.0e3 * dksat * sat ** (2.0 * bexp + 3.0) * (1.0 - fcr)
    smpfz = np.maximum(-120000.0, 0.2 * (-psisat * 1000.0 * sat ** (-bexp)))
    znode4 = -zsoil[2] + 0.5 * (zsoil[2] - zsoil[3])
    qdis0 = (1.0 - fcr) * 5.0 * np.exp(-timean)

    def excess(z):
        return ka * (1.0 + smpfz / (1000.0 * (z - znode4))) - qdis0 * np.exp(-6.0 * (z - 2.0))

    lo = np.full(smc4.shape, -float(zsoil[3]), np.float64)
    hi = np.full(smc4.shape, float(zwt_max), np.float64)
    flo, fhi = excess(lo), excess(hi)
    for _ in range(60):
        mid = 0.5 * (lo + hi)
        up = excess(mid) < 0.0      # discharge still larger: the balance point is deeper
        lo = np.where(up, mid, lo)
        hi = np.where(up, hi, mid)
    zwt = 0.5 * (lo + hi)
    zwt = np.where(flo >= 0.0, -float(zsoil[3]), np.where(fhi <= 0.0, float(zwt_max), zwt))
    wa = ((-float(zsoil[3]) + 25.0) - zwt) * 1000.0 * 0.2
    return zwt.astype(f32), np.minimum(wa, 5000.0).astype(f32)


def columns_from_init(path: str, tables: ParamTables, when: dt.datetime | None = None, static_path: str | None = None,
                      water_table: str = "equilibrium") -> Columns:
    """``.npz`` keys: lat, lon [deg]; lutyp, sltyp, slptyp, isc (ints); shdfac, tbot; tsk (skin T), stc[4], smc[4],
    optional sh2o[4], snowh [m], sneqv [mm], canliq, zwt, wa (both absent: :func:`equilibrium_water_table`).  A NetCDF-3 file with the HRLDAS / WRF variable names
    (``ncio.INIT_ALIASES``) is accepted as well; its grid is flattened row-major."""
    z = ncio.read_init(path, static_path) if ncio.is_netcdf(path) else np.load(path)
    n = z["lat"].size
    st = synth._blank_state(n)
    sl = z["sltyp"].astype(np.int64)
    smcmax = np.asarray(tables.soil["SMCMAX"], f32)[sl - 1]
    S = L.S
    get = lambda k, default: (z[k].astype(f32) if k in z.files else np.full(n, default, f32))  # noqa: E731
    snowh, sneqv = get("snowh", 0.0), get("sneqv", 0.0)
    isnow, dz, snice, zs = snow_layers_from_depth(snowh, sneqv, None)
    st[S["ISNOW"]] = isnow.view(f32)
    multi = isnow < 0
    st[S["SNOWH"]] = snowh
    st[S["SNEQV"]] = sneqv
    st[S["SNEQVO"]] = sneqv
    tsk = z["tsk"].astype(f32)
    for j in range(3):
        st[L.snice(j - 2)] = snice[j]
        st[L.stc(j - 2)] = np.where(dz[j] > 0, np.minimum(tsk, 273.15), 0.0)
        st[L.zsnso(j - 2)] = zs[j]
    for k in range(1, 5):
        st[L.zsnso(k)] = zs[2 + k]
        st[L.stc(k)] = z["stc"][k - 1]
        st[L.smc(k)] = np.minimum(z["smc"][k - 1].astype(f32), smcmax)
        st[L.soilwat(k)] = np.minimum(z["sh2o"][k - 1].astype(f32), st[L.smc(k)]) if "sh2o" in z.files else st[L.smc(k)]
    del multi
    st[S["TV"]] = tsk
    st[S["TG"]] = tsk
    st[S["TAH"]] = tsk
    st[S["EAH"]] = get("eah", 2000.0)
    st[S["ALBOLD"]] = 0.65
    st[S["CANLIQ"]] = get("canliq", 0.0)
    # water table / aquifer: from the file when it carries them, else the equilibrium of the SIMGM scheme for the initial
    # bottom-layer moisture (``water_table="constant"`` keeps upstream's 2.5 m / 4900 mm)
    if "zwt" in z.files and "wa" in z.files:
        st[S["ZWT"]], st[S["WA"]] = z["zwt"].astype(f32), z["wa"].astype(f32)
    elif water_table == "equilibrium":
        st[S["ZWT"]], st[S["WA"]] = equilibrium_water_table(tables, sl, st[L.smc(4)], st[L.soilwat(4)])
    else:
        st[S["ZWT"]] = get("zwt", 2.5)
        st[S["WA"]] = get("wa", 4900.0)
    st[S["WT"]] = st[S["WA"]]
    st[S["CM"]] = 0.1
    st[S["CH"]] = 0.1
    st[S["QSFC"]] = 0.005
    lat, lon = np.deg2rad(z["lat"].astype(np.float64)), np.deg2rad(z["lon"].astype(np.float64))
    for name, val in initial_vegetation(tables, z["lutyp"], lat, when).items():
        st[S[name]] = val
    sf = np.zeros((L.NSTATIC_F, n), f32)
    sf[L.C["LAT"]] = lat
    sf[L.C["SHDFAC"]] = get("shdfac", 0.7)
    sf[L.C["TBOT"]] = get("tbot", 285.0)
    sf[L.C["SHDMAX"]] = get("shdmax", 0.8)
    ones = np.ones(n, np.int64)
    lut = z["lutyp"].astype(np.int64)
    ist = np.where(lut == tables.flags["ISWATER"], 2, 1)          # IST: 1 soil, 2 lake (func:117)
    ice = np.where(lut == tables.flags["ISICE"], 1, 0)            # ICE: 1 land ice (func:116)
    si = L.pack_types(lut, sl, z["slptyp"].astype(np.int64) if "slptyp" in z.files else ones,
                      z["isc"].astype(np.int64) if "isc" in z.files else 4 * ones, ist, ice).astype(np.int32)
    return Columns(st, sf, si, lat, lon, tuple(getattr(z, "grid", ())))


def write_restart(path: str, state: np.ndarray, when: dt.datetime, step: int) -> None:
    os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
    np.savez(path, state=state, time=F.stamp(when), step=step, planes=np.asarray(L.STATE))


def read_restart(path: str):
    if ncio.is_netcdf(path):
        return ncio.read_restart(path)
    z = np.load(path)
    if list(z["planes"]) != L.STATE:
        raise ValueError(f"{path}: state plane list does not match this build")
    return z["state"].astype(f32), dt.datetime.strptime(str(z["time"]), "%Y%m%d%H%M"), int(z["step"])


class _AsyncWriter:
    """Output and restart files are written by a background thread of the writing rank: the planes go to a page-locked host
    buffer with an asynchronous copy, and the time loop goes on while the thread waits for the copy and writes the file."""

    def __init__(self, depth: int = 2):
        import queue
        import threading
        self.q: "queue.Queue" = queue.Queue(maxsize=depth)
        self.busy_s, self.error = 0.0, None
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def _run(self):
        import time as _t
        while True:
            job = self.q.get()
            if job is None:
                return
            tic = _t.perf_counter()
            try:
                job()
            except Exception as exc:      # reported by close(): a failed write must not pass silently
                self.error = exc
            self.busy_s += _t.perf_counter() - tic

    def submit(self, planes_dev, write) -> None:
        """``write(host_array)`` runs on the thread once ``planes_dev`` has arrived in host memory."""
        import torch
        host = torch.empty(planes_dev.shape, dtype=planes_dev.dtype, pin_memory=planes_dev.is_cuda)
        host.copy_(planes_dev, non_blocking=True)
        ev = None
        if planes_dev.is_cuda:
            ev = torch.cuda.Event()
            ev.record()

        def job(host=host, ev=ev, keep=planes_dev):
            if ev is not None:
                ev.synchronize()
            write(host.numpy())
        self.q.put(job)

    def close(self) -> None:
        self.q.put(None)
        self.thread.join()
        if self.error is not None:
            raise self.error


# ------------------------------------------------------------------------------------------------
def run_files(cfg, device: int | None = None, variant="fast", max_steps_per_launch: int = 24) -> dict:
    """The file-driven time loop.  Forcing valid at the start of each step; launches are cut at output and restart
    boundaries and otherwise fuse up to ``max_steps_per_launch`` steps.

    Under ``torchrun`` (WORLD_SIZE > 1) every rank advances its block of columns (``parallel.shard``) on its own GPU with no
    exchange on the step path; at output / restart boundaries the planes are all-gathered and rank 0 writes the files, and
    the budget sums are all-reduced at the end."""
    import torch
    import torch.distributed as dist
    from .parallel import allreduce_budget, device_for_rank, gather_columns, shard
    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    if device is None:
        device = device_for_rank(int(os.environ.get("LOCAL_RANK", "0")), world)
    dev = torch.device("cuda", device)
    own_group = False
    if world > 1 and not dist.is_initialized():
        torch.cuda.set_device(dev)
        dist.init_process_group("nccl", device_id=dev)
        own_group = True
    tables = (ParamTables.from_tbl_dir(cfg.table_directory, cfg.soil_table_tag, cfg.veg_table_tag)
              if getattr(cfg, "table_directory", None) else ParamTables.default(cfg.soil_table_tag, cfg.veg_table_tag))
    cols = columns_from_init(cfg.initfile, tables, cfg.begdatetime, getattr(cfg, "constfile", None))
    as_nc = ncio.is_netcdf(cfg.initfile)    # outputs and restarts follow the format of the initialisation file
    t, step0 = cfg.begdatetime, 0
    res_path = cfg.resfile if os.path.isabs(str(cfg.resfile)) else os.path.join(cfg.resdir, str(cfg.resfile))
    if os.path.isfile(res_path):
        rst, t, step0 = read_restart(res_path)
        if rst.shape != cols.state.shape:
            raise ValueError(f"ERR: restart {res_path} holds {rst.shape[1]} columns x {rst.shape[0]} planes, the initialization "
                             f"file {cfg.initfile} describes {cols.ncol} columns x {cols.state.shape[0]} planes")
        cols.state = rst
    dtm = cfg.timestep
    out_freq, res_freq = F.parse_frequency(cfg.outfreq), F.parse_frequency(cfg.resfreq)
    eng = NoahMP(tables, cfg.physics_options(), device=device, fast=variant)
    ntot = cols.ncol
    c0, ncol = shard(ntot, rank, world)
    mine = slice(c0, c0 + ncol)
    state = torch.from_numpy(np.ascontiguousarray(cols.state[:, mine])).to(dev)
    sf = torch.from_numpy(np.ascontiguousarray(cols.static_f[:, mine])).to(dev)
    si = torch.from_numpy(np.ascontiguousarray(cols.static_i[mine])).to(dev)
    geo = torch.from_numpy(np.ascontiguousarray(F.geo_planes(cols.lat, cols.lon)[:, mine])).to(dev)
    diag = torch.zeros((L.NDIAG, ncol), dtype=torch.float32, device=dev)
    budget = torch.zeros(L.NBUDGET, dtype=torch.float64, device=dev)
    end_slots = torch.tensor([L.B["NCELL"], L.B["SNEQV"], L.B["SOILWATER"]], device=dev)
    # records-driven launches (noahmp_run_records): the forcing records go to the device once each and the step kernel
    # interpolates them and computes COSZ in its own load - no prepared forcing planes, no preparation pass
    max_steps_per_launch = min(max_steps_per_launch, 32)
    reader = F.RecordReader(cfg.indir, cfg.infreq, cfg.begdatetime, ntot, dev)
    # file reads run ahead of the time loop on a thread (page-locked buffers, asynchronous copies); NOAHMP_PREFETCH=0: inline
    prefetch = None
    if os.environ.get("NOAHMP_PREFETCH", "1") != "0":
        prefetch = F.RecordPrefetcher(reader, mine, reader.bracket(t)[0], cfg.enddatetime)
    ring = F.RecordRing(reader, max_steps_per_launch + 2, mine, dev, prefetch)
    writer = _AsyncWriter() if rank == 0 else None
    if world > 1:       # communicator set-up belongs to initialisation, not to the first output of the time loop
        gather_columns(torch.zeros((1, ncol), dtype=torch.float32, device=dev), ntot)
        torch.cuda.synchronize()
    nsteps, written = step0, []
    timing = {"ingest_s": 0.0, "launch_s": 0.0, "gather_s": 0.0, "write_s": 0.0}
    import time as _time
    # wall-time split (ingest = record files -> device ring; kernel = step launches; gather = all-gather + D2H at output /
    # restart cadence; write = files).  NOAHMP_TIMING=1 synchronises after every launch so that `kernel_s` is the device
    # time and not the enqueue time; without it the launches overlap the next ingest and only the total is meaningful.
    detail = os.environ.get("NOAHMP_TIMING", "0") == "1"
    t_wall0 = _time.perf_counter()
    while t < cfg.enddatetime:
        t_launch, plan = t, []
        while len(plan) < max_steps_per_launch and t < cfg.enddatetime:
            t0, t1, wgt = reader.bracket(t)
            plan.append((t0, t1, wgt, F.solar_scalars(t)))
            t = t + dtm
            if F.is_boundary(t, cfg.begdatetime, out_freq) or F.is_boundary(t, cfg.begdatetime, res_freq):
                break
        g = len(plan)
        tic = _time.perf_counter()
        slot = ring.slots_for([p[0] for p in plan] + [p[1] for p in plan])
        timing["ingest_s"] += _time.perf_counter() - tic
        steps = eng.step_table([slot[p[0]] for p in plan], [slot[p[1]] for p in plan], [p[2] for p in plan], [p[3] for p in plan])
        rp = eng.run_params(dtm.total_seconds(), synth.ZSOIL, F.julian_of(t_launch), F.yearlen_of(t_launch))
        budget[end_slots] = 0.0   # NCELL / SNEQV / SOILWATER are end-of-launch snapshots: keep the last launch's only
        tic = _time.perf_counter()
        eng.run_records(rp, state, ring.ring, steps, geo, sf, si, diag, budget)
        if detail:
            torch.cuda.synchronize()
        timing["launch_s"] += _time.perf_counter() - tic
        nsteps += g
        if F.is_boundary(t, cfg.begdatetime, out_freq):
            tic = _time.perf_counter()
            full = gather_columns(diag, ntot)         # only the writing rank brings the planes to the host
            timing["gather_s"] += _time.perf_counter() - tic
            tic = _time.perf_counter()
            if as_nc:
                p = os.path.join(cfg.outdir, F.stamp(t) + ".LDASOUT_DOMAIN1")
            else:
                p = os.path.join(cfg.outdir, F.stamp(t) + ".LDASOUT.npz")
            if rank == 0:
                os.makedirs(cfg.outdir, exist_ok=True)
                if as_nc:
                    writer.submit(full, lambda a, p=p, t=t: ncio.write_output(p, a, t, cols.grid))
                else:
                    writer.submit(full, lambda a, p=p, t=t: np.savez(p, diag=a, planes=np.asarray(L.DIAG), time=F.stamp(t)))
            written.append(p)
            timing["write_s"] += _time.perf_counter() - tic
            diag[L.ACC_DIAG_PLANES[0]:] = 0.0   # ACCPRCP ... UGDRNOFF accumulate from one output to the next
        if F.is_boundary(t, cfg.begdatetime, res_freq):
            tic = _time.perf_counter()
            full = gather_columns(state, ntot)
            if world == 1:
                full = full.clone()      # the one-rank gather is the state tensor itself, which the next launch advances
            timing["gather_s"] += _time.perf_counter() - tic
            tic = _time.perf_counter()
            if as_nc:
                p = os.path.join(cfg.resdir, "RESTART." + F.stamp(t) + "_DOMAIN1")
            else:
                p = os.path.join(cfg.resdir, "RESTART." + F.stamp(t) + ".npz")
            if rank == 0:
                if as_nc:
                    writer.submit(full, lambda a, p=p, t=t, k=nsteps: ncio.write_restart(p, a, t, k, cols.grid))
                else:
                    writer.submit(full, lambda a, p=p, t=t, k=nsteps: write_restart(p, a, t, k))
            written.append(p)
            timing["write_s"] += _time.perf_counter() - tic
    torch.cuda.synchronize()
    tic = _time.perf_counter()
    if writer is not None:
        writer.close()                       # every file is on disk before the run reports
        timing["writer_thread_busy_s"] = writer.busy_s
    timing["write_s"] += _time.perf_counter() - tic
    if prefetch is not None:
        prefetch.close()
        timing["reader_thread_busy_s"] = prefetch.busy_s
    timing["total_s"] = _time.perf_counter() - t_wall0
    timing["synchronised_per_launch"] = detail
    timing["records_loaded"] = ring.loads
    timing["records_prefetched"] = ring.prefetched
    allreduce_budget(budget)
    out = _summary(gather_columns(state, ntot).cpu().numpy(), gather_columns(diag, ntot).cpu().numpy(), budget.cpu().numpy(), ntot,
                   nsteps, dtm.total_seconds(), cfg.physics_options())
    out["files_written"] = written
    out["timing"] = timing
    out["launches"] = eng.launch_count()
    out["ranks"] = world
    eng.close()
    if own_group:
        dist.barrier()
        dist.destroy_process_group()
    return out


def _summary(st, dg, bd, ncol, nsteps, dt_s, opts) -> dict:
    out = {"columns": ncol, "steps": nsteps, "dt_s": dt_s, "options": opts,
           "flagged_column_launches": int(bd[L.B["NFLAGGED"]]), "max_abs_water_error_mm": float(bd[L.B["ERRWAT_MAXABS"]])}
    for name in ("TV", "TG", "EAH", "SNEQV", "ZWT"):
        out[name] = float(st[L.S[name]].mean())
    for k in range(1, 5):
        out[f"STC{k}"] = float(st[L.stc(k)].mean())
        out[f"SMC{k}"] = float(st[L.smc(k)].mean())
    for name in ("FSA", "FIRA", "FSH", "SSOIL", "TRAD"):
        out[name] = float(dg[L.D[name]].mean())
    return out


def steps_of(cfg) -> int:
    return int((cfg.enddatetime - cfg.begdatetime).total_seconds() // cfg.timestep.total_seconds())


def run_case(cfg, ncol: int = 1, chunk: int = 96, device: int = 0, variant="fast") -> dict:
    """Entry used by run/main.py: file driven when the namelist's initialization file exists, else synthetic."""
    if os.path.isfile(str(cfg.initfile)):
        return run_files(cfg, device=device if "LOCAL_RANK" not in os.environ else None, variant=variant)
    import torch
    tables = (ParamTables.from_tbl_dir(cfg.table_directory, cfg.soil_table_tag, cfg.veg_table_tag)
              if getattr(cfg, "table_directory", None) else ParamTables.default(cfg.soil_table_tag, cfg.veg_table_tag))
    opts = cfg.physics_options()
    nsteps = steps_of(cfg)
    w = synth.make_workload("single", ncol=ncol, nsteps=nsteps, tables=tables, opts=opts)
    w.dt = float(cfg.timestep.total_seconds())
    w.yearlen = F.yearlen_of(cfg.begdatetime)
    w.julian0 = F.julian_of(cfg.begdatetime)
    w.hour0 = cfg.begdatetime.hour + cfg.begdatetime.minute / 60.0
    eng = NoahMP(tables, opts, device=device, fast=variant)
    dev = torch.device("cuda", device)
    state = torch.from_numpy(w.state).to(dev)
    static_f = torch.from_numpy(w.static_f).to(dev)
    static_i = torch.from_numpy(w.static_i).to(dev)
    diag = torch.zeros((L.NDIAG, ncol), dtype=torch.float32, device=dev)
    budget = torch.zeros(L.NBUDGET, dtype=torch.float64, device=dev)
    end_slots = torch.tensor([L.B["NCELL"], L.B["SNEQV"], L.B["SOILWATER"]], device=dev)
    done = 0
    while done < nsteps:  # forcing is produced chunk by chunk, state stays on the device
        n = min(chunk, nsteps - done)
        fo = np.stack([synth.forcing_step(w, done + s) for s in range(n)])
        t_chunk = cfg.begdatetime + dt.timedelta(seconds=done * w.dt)   # calendar date of the chunk: JULIAN / YEARLEN follow it
        rp = eng.run_params(w.dt, synth.ZSOIL, F.julian_of(t_chunk), F.yearlen_of(t_chunk), w.zref, w.co2_frac, w.o2_frac, w.foln)
        budget[end_slots] = 0.0
        eng.run(rp, state, torch.from_numpy(fo).to(dev), static_f, static_i, diag, budget)
        done += n
    torch.cuda.synchronize()
    out = _summary(state.cpu().numpy(), diag.cpu().numpy(), budget.cpu().numpy(), ncol, nsteps, w.dt, opts)
    eng.close()
    return out
