#!/usr/bin/env python3
"""Live comparison of the hand-written oracle with the mechanically translated reference (needs /root/reference):
any workload, any size, any options - beyond the committed fixtures.  Prints the number of columns in which ANY of the
57 state planes or 58 outputs differs in ANY bit after the run.

    python tools/compare_oracle_reference.py [--ncol N] [--steps S] [workload ...]       # default: all five
    python tools/compare_oracle_reference.py --ncol 24 --steps 8760 global025              # a whole year
    python tools/compare_oracle_reference.py --options --ncol 2048                          # every value of every option
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import make_golden  # noqa: E402
from noahmp_b200 import layout as L, synth  # noqa: E402
from noahmp_b200.tables import ParamTables  # noqa: E402
from oracle.oracle import Oracle  # noqa: E402
from oracle.refrun import RefModel  # noqa: E402


def compare(workload, ncol, nsteps, over, tables):
    opts = make_golden.options(over)
    w = synth.make_workload(workload, ncol=ncol, nsteps=nsteps, opts=opts)
    f = synth.forcing_all(w)
    sr, dr = w.state.copy(), np.zeros((L.NDIAG, w.ncol), np.float32)
    t0 = time.time()
    fatal = RefModel(tables, opts).run_native(sr, f, w.static_f, w.static_i, w.dt, synth.ZSOIL, w.julian0, w.yearlen, w.zref,
                                              w.co2_frac, w.o2_frac, w.foln, diag=dr)     # generated C++ column loop, OpenMP
    so, do = w.state.copy(), np.zeros((L.NDIAG, w.ncol), np.float32)
    nf = Oracle(tables, opts).run_workload(w, f, diag=do, state=so)
    ok = ~fatal
    same = (sr.view(np.int32) == so.view(np.int32)) | (np.isnan(sr) & np.isnan(so))
    dsame = (dr[:58].view(np.int32) == do[:58].view(np.int32)) | (np.isnan(dr[:58]) & np.isnan(do[:58]))
    bad = (~same[:, ok]).any(axis=0) | (~dsame[:, ok]).any(axis=0)
    print(f"{workload:12s} {w.ncol:5d} x {nsteps:4d} {str(over):40s} reference aborted {int(fatal.sum()):3d}, oracle flagged {nf:3d}, "
          f"columns differing {int(bad.sum()):4d}   ({time.time() - t0:.0f} s)", flush=True)
    if bad.any():
        print("      state:", [L.STATE[p] for p in range(L.NSTATE) if (~same[p, ok]).any()][:10],
              "outputs:", [L.DIAG[p] for p in range(58) if (~dsame[p, ok]).any()][:10])
    return int(bad.sum())


def main():
    argv = sys.argv[1:]
    ncol, steps = 1024, None
    for flag in ("--ncol", "--steps"):
        if flag in argv:
            i = argv.index(flag)
            v = int(argv[i + 1])
            del argv[i:i + 2]
            ncol, steps = (v, steps) if flag == "--ncol" else (ncol, v)
    t = ParamTables.default()
    total = 0
    if "--options" in argv:
        argv.remove("--options")
        for k, vals in (("veg", (2, 3, 4, 5)), ("crs", (2,)), ("btr", (2, 3)), ("run", (2, 3, 4)), ("sfc", (2,)), ("frz", (2,)),
                        ("inf", (2,)), ("rad", (2, 3)), ("alb", (1,)), ("snf", (2, 3)), ("tbot", (2,)), ("stc", (2,))):
            for v in vals:
                for wl in ("global025", "snow1m"):
                    total += compare(wl, ncol, steps or 48, {k: v}, t)
        print("TOTAL differing columns", total)
        return 1 if total else 0
    for wl in argv or ["uniform100k", "global025", "snow1m", "conus1km", "single"]:
        total += compare(wl, 1 if wl == "single" else ncol, steps or {"uniform100k": 48, "conus1km": 24}.get(wl, 96), {}, t)
    print("TOTAL differing columns", total)
    return 1 if total else 0


if __name__ == "__main__":
    sys.exit(main())
