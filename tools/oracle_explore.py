#!/usr/bin/env python3
"""Run the CPU oracle on a (subsampled) workload and print state / flux / closure summaries.
Usage: python tools/oracle_explore.py NAME [NCOL] [NSTEPS] [opt=val ...]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from noahmp_b200 import layout as L, synth  # noqa: E402
from noahmp_b200.tables import ParamTables  # noqa: E402
from oracle.oracle import Oracle  # noqa: E402

OPT_NAMES = ["veg", "crs", "btr", "run", "sfc", "frz", "inf", "rad", "alb", "snf", "tbot", "stc"]


def main():
    name = sys.argv[1]
    ncol = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
    nsteps = int(sys.argv[3]) if len(sys.argv) > 3 else None
    opts = list(synth.DEFAULT_OPTS)
    for a in sys.argv[4:]:
        k, v = a.split("=")
        opts[OPT_NAMES.index(k)] = int(v)
    t = ParamTables.default()
    w = synth.make_workload(name, ncol=ncol, nsteps=nsteps, opts=opts)
    f = synth.forcing_all(w)
    o = Oracle(t, w.opts)
    diag = np.zeros((L.NDIAG, w.ncol), np.float32)
    worst = {k: 0.0 for k in ("ERRWAT", "ERRSW", "ERRENG")}
    flagged = np.zeros(w.ncol, np.int32)
    t0 = time.time()
    # step one at a time so closure residuals of EVERY step are seen
    for s in range(w.nsteps):
        o.run(w.state, f[s:s + 1], w.static_f, w.static_i, dt=w.dt, zsoil=synth.ZSOIL,
              julian0=w.julian0 + s * w.dt / 86400.0, yearlen=w.yearlen, zref=w.zref, diag=diag)
        for k in worst:
            worst[k] = max(worst[k], float(np.nanmax(np.abs(diag[L.D[k]]))))
        flagged |= diag[L.D["FLAGS"]].view(np.int32)
    dt_ = time.time() - t0
    print(f"{name}: {w.ncol} cols x {w.nsteps} steps in {dt_:.2f}s = {w.ncol * w.nsteps / dt_:.3g} cell-steps/s "
          f"({o.num_threads()} threads)")
    print("worst |residual| over all steps:", worst)
    print("flag histogram:", {k: int(((flagged & b) != 0).sum()) for k, b in L.FLAG_BITS.items()})
    st = w.state
    isn = st[L.S["ISNOW"]].view(np.int32)
    print("ISNOW histogram (0,-1,-2,-3):", [int((isn == -k).sum()) for k in range(4)])
    for n in ["TV", "TG", "EAH", "SNOWH", "SNEQV", "ZWT", "WA", "CANLIQ", "CANICE", "FWET", "ALBOLD", "TAUSS"]:
        a = st[L.S[n]]
        print(f"  {n:7s} min {np.nanmin(a):12.5g} mean {np.nanmean(a):12.5g} max {np.nanmax(a):12.5g} nan {int(np.isnan(a).sum())}")
    for k in range(1, 5):
        print(f"  soil {k}: STC {st[L.stc(k)].mean():9.4f}  SMC {st[L.smc(k)].mean():.5f}  LIQ {st[L.soilwat(k)].mean():.5f}")
    for n in ["FSA", "FSR", "FIRA", "FSH", "SSOIL", "FCEV", "FGEV", "FCTR", "TRAD", "T2MV", "T2MB", "RSSUN", "PSN",
              "RUNSRF", "RUNSUB", "FSNO", "ALBEDO"]:
        a = diag[L.D[n]]
        print(f"  {n:7s} min {np.nanmin(a):12.5g} mean {np.nanmean(a):12.5g} max {np.nanmax(a):12.5g}")
    it = diag[L.D["ITER"]].view(np.int32)
    print("  loop1 iters:", np.bincount(it & 0xff).tolist())
    print("  stomata sun iters:", np.bincount((it >> 8) & 0xff).tolist())
    print("  soil substeps:", np.bincount((it >> 24) & 0xff).tolist())


if __name__ == "__main__":
    main()
