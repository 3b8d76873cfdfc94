"""Synthetic columns and forcing for the BASELINE.json configurations.

The reference ships no forcing, no ``init.nc`` and no driver (``run/main.py:12-14`` only prints the
config), so the workloads are generated.  The per-step forcing generator is a counter-based integer
hash followed by +,-,*,/ only, with every transcendental evaluated once per step on the host
(:func:`step_scalars`); the same arithmetic is implemented on the device (``synth_forcing_kernel`` in ``csrc/noahmp_abi.cu``),
so host and device forcing are bit-identical and the CPU oracle and the CUDA kernel can be fed the
same numbers at any size.

Initial states follow SURVEY App. G (cold-start practice) and stay inside the validity envelope listed
there (no soil type 14, no water/ice land-use rows, 0 < SMC <= SMCMAX, consistent snow layering).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field, asdict

import numpy as np

from . import layout as L
from .tables import ParamTables

f32 = np.float32
ZSOIL = np.array([-0.1, -0.4, -1.0, -2.0], f32)   # WRF standard 4-layer grid (not in the reference repo)

TFRZ, HFUS, GRAV = 273.15, 0.3336e6, 9.80616


@dataclass
class SynthParams:
    """Coefficients of the forcing generator (all fp32 on the device)."""
    seed: int = 1234
    t_mean: float = 288.0       # K
    t_lat: float = 0.0          # K * sin(lat)^2
    t_season: float = 0.0       # K, multiplies cos(2pi(doy-200)/yearlen) * sin(lat)^2  (per-step scalar)
    t_diurnal: float = 8.0      # K, multiplies sin(2pi(h-9)/24)                        (per-step scalar)
    t_noise: float = 1.0        # K * approx-N(0,1)
    sw_max: float = 900.0       # W/m2 * max(cosz,0)
    lw_mean: float = 330.0      # W/m2
    lw_tcoef: float = 4.5       # W/m2/K * (T - t_mean)
    lw_noise: float = 10.0
    p_mean: float = 95000.0     # Pa
    p_noise: float = 2000.0     # uniform +-
    wind_min: float = 1.0
    wind_max: float = 8.0
    q_mean: float = 0.007
    q_noise: float = 0.2857     # relative, uniform +-
    q_tcoef: float = 0.0        # 1/K: q follows q_mean*(1+q_tcoef*(T-t_mean)/16)^16 ~ Clausius-Clapeyron
    prcp_prob: float = 0.05
    prcp_scale: float = 3.0e-4  # kg m-2 s-1 * u^2   (mean = scale/3)


@dataclass
class Workload:
    name: str
    ncol: int
    nsteps: int
    dt: float
    julian0: float
    yearlen: int
    hour0: float
    opts: list
    synth: SynthParams
    state: np.ndarray = field(repr=False, default=None)      # [NSTATE][ncol] fp32
    static_f: np.ndarray = field(repr=False, default=None)   # [NSTATIC_F][ncol] fp32
    static_i: np.ndarray = field(repr=False, default=None)   # [ncol] int32 packed types
    sincos: np.ndarray = field(repr=False, default=None)     # [2][ncol] fp32 sin(lat), cos(lat)
    zref: float = 10.0
    co2_frac: float = 355.0e-6
    o2_frac: float = 0.209
    foln: float = 1.0


# the reference's shipped option set: run/case.nml:29-37 + defaults of global.f90:24,41,45,60,73
#            veg crs btr run sfc frz inf rad alb snf tbot stc
DEFAULT_OPTS = [1, 1, 1, 1, 1, 1, 1, 1, 2, 1, 1, 1]


# ------------------------------------------------------------------------------------------------
# counter-based hash -> uniform [0,1) with 24 random bits (exact in fp32)
# ------------------------------------------------------------------------------------------------
def _mix(h: np.ndarray) -> np.ndarray:
    h = h.astype(np.uint32)
    h ^= h >> np.uint32(16)
    h = (h * np.uint32(0x7feb352d)).astype(np.uint32)
    h ^= h >> np.uint32(15)
    h = (h * np.uint32(0x846ca68b)).astype(np.uint32)
    h ^= h >> np.uint32(16)
    return h


def uniform(seed: int, step: int, col: np.ndarray, stream: int) -> np.ndarray:
    with np.errstate(over="ignore"):
        key = np.uint32((seed * 0x9E3779B1 + step * 0x85EBCA77 + stream * 0xC2B2AE3D) & 0xFFFFFFFF)
        h = _mix(col.astype(np.uint32) ^ key)
        h = _mix(h + np.uint32(0x27D4EB2F))
    return (h >> np.uint32(8)).astype(f32) * f32(1.0 / 16777216.0)


def step_scalars(w: Workload, step: int) -> dict:
    """Host-evaluated transcendental factors of one step (passed to the device generator as fp32)."""
    t_sec = step * w.dt
    hour = (w.hour0 + t_sec / 3600.0) % 24.0
    doy = w.julian0 + t_sec / 86400.0
    decl = 23.44 * math.pi / 180.0 * math.cos(2.0 * math.pi * (doy - 172.0) / w.yearlen)
    hang = 2.0 * math.pi * (hour - 12.0) / 24.0
    return {"sd": f32(math.sin(decl)), "cdch": f32(math.cos(decl) * math.cos(hang)),
            "t_diur": f32(w.synth.t_diurnal * math.sin(2.0 * math.pi * (hour - 9.0) / 24.0)),
            "t_seas": f32(w.synth.t_season * math.cos(2.0 * math.pi * (doy - 200.0) / w.yearlen))}


def forcing_step(w: Workload, step: int, col0: int = 0, out: np.ndarray | None = None, cols=None) -> np.ndarray:
    """[NFORCING][ncol] fp32 forcing of one step.  ``col0`` = global index of this shard's first column
    (multi-GPU shards hash GLOBAL column ids so the forcing does not depend on the rank count); ``cols`` =
    explicit global column ids (for a strided sample of a larger workload)."""
    p = w.synth
    n = w.sincos.shape[1]
    col = (np.arange(col0, col0 + n, dtype=np.uint64) if cols is None else np.asarray(cols, np.uint64)).astype(np.uint32)
    sc = step_scalars(w, step)
    sinlat, coslat = w.sincos[0], w.sincos[1]
    u = [uniform(p.seed, step, col, s) for s in range(12)]
    two, sqrt3 = f32(2.0), f32(1.7320508)
    g1 = (((u[0] + u[1]) + (u[2] + u[3])) - two) * sqrt3          # approx N(0,1) (Irwin-Hall, n=4)
    g2 = (((u[4] + u[5]) + (u[6] + u[7])) - two) * sqrt3
    s2 = sinlat * sinlat
    tcore = (f32(p.t_mean) + (f32(p.t_lat) + sc["t_seas"]) * s2) + sc["t_diur"]
    T = tcore + f32(p.t_noise) * g1
    cosz = sinlat * sc["sd"] + coslat * sc["cdch"]
    cosz = np.minimum(np.maximum(cosz, f32(-1.0)), f32(1.0))   # fp32 rounding can give 1.0000001: acos() in twostream needs <= 1
    sw = f32(p.sw_max) * np.maximum(cosz, f32(0.0))
    lw = (f32(p.lw_mean) + f32(p.lw_tcoef) * (T - f32(p.t_mean))) + f32(p.lw_noise) * g2
    ps = f32(p.p_mean) + f32(p.p_noise) * (two * u[8] - f32(1.0))
    wind = f32(p.wind_min) + f32(p.wind_max - p.wind_min) * u[9]
    uu = wind * f32(0.8)
    vv = wind * f32(0.6)
    e = np.maximum(f32(1.0) + (f32(p.q_tcoef) * (tcore - f32(p.t_mean))) * f32(0.0625), f32(0.05))
    e = e * e
    e = e * e
    e = e * e
    e = e * e                                                      # ~exp(q_tcoef*(T-t_mean)), no libm call
    q = (f32(p.q_mean) * e) * (f32(1.0) + f32(p.q_noise) * (two * u[10] - f32(1.0)))
    q = np.maximum(q, f32(1.0e-4))
    wet = u[11] < f32(p.prcp_prob)
    ur = u[11] * f32(1.0 / max(p.prcp_prob, 1e-9))                # re-use the draw: uniform on wet cells
    pr = np.where(wet, f32(p.prcp_scale) * (ur * ur), f32(0.0)).astype(f32)
    if out is None:
        out = np.empty((L.NFORCING, n), f32)
    out[L.F["SFCTMP"]] = T
    out[L.F["SFCPRS"]] = ps
    out[L.F["UU"]] = uu
    out[L.F["VV"]] = vv
    out[L.F["Q2"]] = q
    out[L.F["SOLDN"]] = sw
    out[L.F["LWDN"]] = lw
    out[L.F["PRCP"]] = pr
    out[L.F["COSZ"]] = cosz
    return out


def forcing_all(w: Workload, col0: int = 0, cols=None) -> np.ndarray:
    out = np.empty((w.nsteps, L.NFORCING, w.sincos.shape[1]), f32)
    for s in range(w.nsteps):
        forcing_step(w, s, col0, out[s], cols=cols)
    return out


# ------------------------------------------------------------------------------------------------
# initial columns
# ------------------------------------------------------------------------------------------------
def _blank_state(ncol: int) -> np.ndarray:
    st = np.zeros((L.NSTATE, ncol), f32)
    st[L.S["ISNOW"]] = np.zeros(ncol, np.int32).view(f32)
    return st


def _cold_start(st, tables: ParamTables, sltyp, tsoil, smc_frac, tv=None):
    """App. G cold start: no snow, canopy dry, TV=TG=STC(1), EAH 2000 Pa, ZWT 2.5 m, WA=WT=4900 mm."""
    ncol = st.shape[1]
    smcmax = np.asarray(tables.soil["SMCMAX"], f32)[sltyp - 1]
    S = L.S
    st[S["ALBOLD"]] = 0.65
    st[S["EAH"]] = 1500.0
    for k in range(1, 5):
        st[L.stc(k)] = tsoil[k - 1]
        st[L.smc(k)] = smc_frac[k - 1] * smcmax
        st[L.soilwat(k)] = st[L.smc(k)]
        st[L.zsnso(k)] = ZSOIL[k - 1]
    st[S["TV"]] = st[L.stc(1)] if tv is None else tv
    st[S["TG"]] = st[L.stc(1)]
    st[S["ZWT"]] = 2.5
    st[S["WA"]] = 4900.0
    st[S["WT"]] = 4900.0
    st[S["TAH"]] = st[S["TV"]]
    st[S["CM"]] = 0.1
    st[S["CH"]] = 0.1
    st[S["QSFC"]] = 0.005
    st[S["LFMASS"]] = 50.0
    st[S["RTMASS"]] = 500.0
    st[S["STMASS"]] = 10.0
    st[S["WOOD"]] = 500.0
    st[S["STBLCP"]] = 1000.0
    st[S["FASTCP"]] = 1000.0
    st[S["LAI"]] = 2.0
    st[S["SAI"]] = 0.5
    return ncol


def _supercool(tables: ParamTables, sltyp, T):
    """Liquid water that stays unfrozen at T<TFRZ (func:4376-4378 relation), volumetric."""
    smcmax = np.asarray(tables.soil["SMCMAX"], np.float64)[sltyp - 1]
    psisat = np.asarray(tables.soil["PSISAT"], np.float64)[sltyp - 1]
    bexp = np.asarray(tables.soil["BEXP"], np.float64)[sltyp - 1]
    T = np.asarray(T, np.float64)
    smp = HFUS * np.maximum(TFRZ - T, 1e-6) / (GRAV * T)
    return (smcmax * (smp / psisat) ** (-1.0 / bexp)).astype(f32)


def _finish(w: Workload, lat_deg, shdfac, tbot, lutyp, sltyp, slptyp, isc):
    ncol = w.ncol
    lat = np.deg2rad(np.asarray(lat_deg, np.float64))
    w.static_f = np.zeros((L.NSTATIC_F, ncol), f32)
    w.static_f[L.C["LAT"]] = lat.astype(f32)
    w.static_f[L.C["SHDFAC"]] = shdfac
    w.static_f[L.C["TBOT"]] = tbot
    w.static_f[L.C["SHDMAX"]] = np.minimum(np.asarray(shdfac, f32) + f32(0.1), f32(1.0))
    ist = np.ones(ncol, np.int64)
    ice = np.zeros(ncol, np.int64)
    w.static_i = L.pack_types(np.asarray(lutyp, np.int64), np.asarray(sltyp, np.int64),
                              np.asarray(slptyp, np.int64), np.asarray(isc, np.int64), ist, ice).astype(np.int32)
    w.sincos = np.stack([np.sin(lat), np.cos(lat)]).astype(f32)
    return w


def make_workload(name: str, ncol: int | None = None, nsteps: int | None = None,
                  tables: ParamTables | None = None, opts: list | None = None) -> Workload:
    """Build one of the BASELINE.json configs: 'single', 'uniform100k', 'global025',
    'conus1km', 'snow1m'.  ``ncol``/``nsteps`` override the nominal size (parity tests subsample)."""
    tables = tables or ParamTables.default()
    opts = list(opts or DEFAULT_OPTS)
    if name == "single":       # config 1: run/case.nml as shipped -- 1 column, DT=900 s, 96 steps, 2000-01-01
        n = ncol or 1
        w = Workload(name, n, nsteps or 96, 900.0, 0.0, 366, 0.0, opts,
                     SynthParams(seed=1, t_mean=275.0, t_diurnal=5.0, t_noise=0.5, sw_max=700.0, lw_mean=280.0,
                                 q_mean=0.003, q_noise=0.15, prcp_prob=0.05))
        st = _blank_state(n)
        sl = np.full(n, 6)
        _cold_start(st, tables, sl, [276.0, 277.0, 278.0, 279.0], [0.6, 0.6, 0.6, 0.6])
        w.state = st
        return _finish(w, np.full(n, 40.0), 0.7, 285.0, np.full(n, 7), sl, np.full(n, 1), np.full(n, 4))
    if name == "uniform100k":  # config 2
        n = ncol or 100_000
        w = Workload(name, n, nsteps or 48, 1800.0, 180.0, 365, 0.0, opts, SynthParams(seed=1234))
        rng = np.random.default_rng(1234)
        st = _blank_state(n)
        sl = np.full(n, 6)
        smcmax6 = np.asarray(tables.soil["SMCMAX"], f32)[5]
        fr = f32(0.25) / smcmax6
        _cold_start(st, tables, sl, [285.0, 284.0, 283.0, 282.0], [fr, fr, fr, fr])
        st[L.S["TV"]] = 285.0
        st[L.S["TG"]] = 285.0
        w.state = st
        return _finish(w, rng.uniform(25.0, 50.0, n), 0.7, 283.0, np.full(n, 7), sl, np.full(n, 1), np.full(n, 4))
    if name in ("global025", "conus1km"):  # configs 3 and 4: mixed types, latitude-dependent climate
        n = ncol or (280_000 if name == "global025" else 17_000_000)
        seed = 2024 if name == "global025" else 4242
        rng = np.random.default_rng(seed)
        if name == "global025":
            lat = np.degrees(np.arcsin(rng.uniform(np.sin(np.radians(-55.0)), np.sin(np.radians(72.0)), n)))
            w = Workload(name, n, nsteps or 8760, 3600.0, 0.0, 365, 0.0, opts,
                         SynthParams(seed=seed, t_mean=300.0, t_lat=-38.0, t_season=-16.0, t_diurnal=6.0,
                                     t_noise=1.5, lw_mean=400.0, q_mean=0.0155, q_noise=0.15, q_tcoef=0.065,
                                     prcp_prob=0.06))
        else:
            lat = rng.uniform(25.0, 50.0, n)
            w = Workload(name, n, nsteps or 24, 3600.0, 180.0, 365, 0.0, opts,
                         SynthParams(seed=seed, t_mean=298.0, t_lat=-20.0, t_diurnal=8.0, t_noise=1.5,
                                     lw_mean=380.0, q_mean=0.012, q_noise=0.15, q_tcoef=0.06))
        a = np.abs(lat)
        # land-use by latitude band from USGS {2,5,7,8,10,11,14,15,19,21}
        bands = [(0, 12, [10, 11, 15, 2]), (12, 25, [10, 8, 7, 19]), (25, 38, [7, 8, 2, 5, 19]),
                 (38, 50, [2, 5, 7, 11, 15]), (50, 62, [14, 15, 7, 11]), (62, 91, [21, 14, 19, 7])]
        lutyp = np.zeros(n, np.int64)
        pick = rng.integers(0, 1 << 30, n)
        for lo, hi, types in bands:
            m = (a >= lo) & (a < hi)
            lutyp[m] = np.asarray(types)[pick[m] % len(types)]
        sltyp = rng.integers(1, 13, n)
        sinl2 = np.sin(np.radians(lat)) ** 2
        doy = w.julian0
        tclim = (w.synth.t_mean + (w.synth.t_lat + w.synth.t_season * math.cos(2 * math.pi * (doy - 200.0) / 365.0)) * sinl2)
        st = _blank_state(n)
        tsoil = [np.maximum(tclim - 0.5 * k, 250.0).astype(f32) for k in range(4)]
        fr = rng.uniform(0.45, 0.8, n).astype(f32)
        _cold_start(st, tables, sltyp, tsoil, [fr, fr, fr, fr])
        # frozen soil starts with the equilibrium supercooled liquid fraction
        for k in range(1, 5):
            T = st[L.stc(k)]
            frozen = T < TFRZ
            liq = np.minimum(st[L.smc(k)], _supercool(tables, sltyp, np.minimum(T, TFRZ - 0.01)))
            st[L.soilwat(k)] = np.where(frozen, liq, st[L.soilwat(k)])
        st[L.S["EAH"]] = np.clip(610.0 * np.exp(17.27 * (tclim - 273.15) / (tclim - 35.85)) * 0.6, 50.0, 3500.0)
        w.state = st
        shd = rng.uniform(0.3, 0.95, n).astype(f32)
        return _finish(w, lat, shd, np.maximum(tclim - 2.0, 255.0).astype(f32), lutyp, sltyp,
                       rng.integers(1, 8, n), rng.integers(1, 9, n))
    if name == "snow1m":  # config 5: high-latitude ensemble, active 3-layer pack, frozen soil
        n = ncol or 1_000_000
        rng = np.random.default_rng(55)
        w = Workload(name, n, nsteps or 720, 3600.0, 60.0, 365, 0.0, opts,
                     SynthParams(seed=55, t_mean=272.0, t_lat=0.0, t_diurnal=6.0, t_noise=1.5, sw_max=600.0,
                                 lw_mean=270.0, lw_tcoef=4.0, q_mean=0.0028, q_noise=0.15, q_tcoef=0.075,
                                 prcp_prob=0.08, prcp_scale=2.0e-4))
        lat = rng.uniform(55.0, 75.0, n)
        lutyp = np.asarray([14, 15, 21, 22])[rng.integers(0, 4, n)]
        sltyp = rng.integers(1, 13, n)
        st = _blank_state(n)
        tsoil = [rng.uniform(268.0, 272.0, n).astype(f32) for _ in range(4)]
        smcmax = np.asarray(tables.soil["SMCMAX"], f32)[sltyp - 1]
        fr = np.minimum(f32(0.35) / smcmax, f32(0.95)).astype(f32)   # SMC 0.35, but never above porosity (sand)
        _cold_start(st, tables, sltyp, tsoil, [fr, fr, fr, fr])
        for k in range(1, 5):
            st[L.soilwat(k)] = np.minimum(st[L.smc(k)], _supercool(tables, sltyp, st[L.stc(k)]))
        # 3-layer pack: 0.05 / 0.15 / 0.40 m at 250 kg m-3, snow temperatures 263..270 K
        dz = np.array([0.05, 0.15, 0.40], f32)
        zs = -np.cumsum(dz)
        tsn = [rng.uniform(263.0, 266.0, n), rng.uniform(265.0, 268.0, n), rng.uniform(267.0, 270.0, n)]
        for j in range(3):  # Fortran layers -2, -1, 0
            st[L.stc(j - 2)] = tsn[j].astype(f32)
            st[L.zsnso(j - 2)] = zs[j]
            st[L.snice(j - 2)] = f32(250.0) * dz[j]
            st[L.snliq(j - 2)] = 0.0
        for k in range(1, 5):
            st[L.zsnso(k)] = zs[2] + ZSOIL[k - 1]
        st[L.S["ISNOW"]] = np.full(n, -3, np.int32).view(f32)
        st[L.S["SNOWH"]] = f32(dz.sum())
        st[L.S["SNEQV"]] = f32(250.0) * f32(dz.sum())
        st[L.S["SNEQVO"]] = st[L.S["SNEQV"]]
        st[L.S["TV"]] = tsn[0].astype(f32)
        st[L.S["TG"]] = tsn[0].astype(f32)
        st[L.S["TAH"]] = tsn[0].astype(f32)
        st[L.S["EAH"]] = 300.0
        st[L.S["ALBOLD"]] = 0.75
        w.state = st
        return _finish(w, lat, rng.uniform(0.3, 0.9, n).astype(f32), 271.0, lutyp, sltyp,
                       rng.integers(1, 8, n), rng.integers(1, 9, n))
    raise ValueError(f"unknown workload {name!r}")


def synth_params_array(p: SynthParams) -> np.ndarray:
    """fp32 vector handed to the device generator (order = struct noahmp_synth_params)."""
    return np.asarray([p.t_mean, p.t_lat, p.t_noise, p.sw_max, p.lw_mean, p.lw_tcoef, p.lw_noise, p.p_mean,
                       p.p_noise, p.wind_min, p.wind_max - p.wind_min, p.q_mean, p.q_tcoef, p.q_noise,
                       p.prcp_prob, 1.0 / max(p.prcp_prob, 1e-9), p.prcp_scale], f32)


def describe(w: Workload) -> dict:
    d = {k: v for k, v in asdict(w).items() if k not in ("state", "static_f", "static_i", "sincos")}
    return d
