#!/usr/bin/env python3
"""Generate tests/golden/ref_*.npz: outputs of the REFERENCE's own source, executed here through the mechanical
Fortran -> C++ translation of oracle/f90_to_cpp.py (there is no Fortran compiler in the image).

Needs /root/reference (read-only).  For every case the synthetic inputs are rebuilt from their seeds
(noahmp_b200/synth.py), every column is stepped through `noahmp_sflx` of the translated module, and the end state
(57 planes) plus the 58 outputs of the last step are stored.  tests/test_reference_golden.py then demands that the
hand-written oracle reproduces them BIT FOR BIT (default build, same glibc) - this is what pins the oracle, and through
it the CUDA kernels, to the reference.

    python tools/make_reference_golden.py            # all cases, ~2 minutes
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import make_golden  # noqa: E402  (options() helper)
from noahmp_b200 import layout as L, synth  # noqa: E402
from noahmp_b200.tables import ParamTables  # noqa: E402

# name, workload, columns, steps, option overrides -- every workload with the shipped options, then every non-default
# value of every option on the workload that exercises it (the matrix of tests/test_parity_gpu.py STRICT_CASES)
CASES = [("uniform", "uniform100k", 128, 48, {}), ("global", "global025", 128, 96, {}), ("snow", "snow1m", 128, 96, {}),
         ("conus", "conus1km", 128, 24, {}), ("single", "single", 1, 96, {})]
for _k, _vals in (("veg", (2, 3, 4, 5)), ("crs", (2,)), ("btr", (2, 3)), ("run", (2, 3, 4)), ("sfc", (2,)), ("frz", (2,)),
                  ("inf", (2,)), ("rad", (2, 3)), ("alb", (1,)), ("snf", (2, 3)), ("tbot", (2,)), ("stc", (2,))):
    for _v in _vals:
        CASES.append((f"{_k}{_v}", "snow1m" if _k in ("frz", "snf", "alb", "stc", "tbot") else "global025", 64, 36, {_k: _v}))
CASES.append(("noah_like", "global025", 64, 36, {"sfc": 2, "run": 3, "frz": 2, "inf": 2, "alb": 1, "tbot": 2, "stc": 2, "snf": 3}))
# surface types the generators never draw: urban, evergreen broadleaf, lake (IST=2), glacier (ICE=1), barren, tundra, ...
CASES += [("special_global", "global025", 120, 48, {}), ("special_snow", "snow1m", 120, 48, {})]
SPECIAL_TYPES = [(1, 1, 0), (13, 1, 0), (16, 2, 0), (24, 1, 1), (19, 1, 0), (23, 1, 0), (20, 1, 0), (27, 1, 0)]   # lutyp, IST, ICE


def case_workload(name):
    """The synthetic inputs of a case (rebuilt from the generators' seeds) and its option vector."""
    _, workload, ncol, nsteps, over = next(c for c in CASES if c[0] == name)
    opts = make_golden.options(over)
    w = synth.make_workload(workload, ncol=ncol, nsteps=nsteps, opts=opts)
    if name.startswith("special"):
        si = w.static_i.copy()
        for i in range(w.ncol):
            lu, ist, ice = SPECIAL_TYPES[i % len(SPECIAL_TYPES)]
            sl = 16 if ice else (14 if ist == 2 else (int(si[i]) >> 8) & 0xff)           # STAS: 16 = land ice, 14 = water
            si[i] = L.pack_types(lu, sl, (int(si[i]) >> 16) & 0xff, (int(si[i]) >> 24) & 0xf, ist, ice)
        w.static_i = si
    return w, opts


def run_reference(name, tables=None):
    from oracle.refrun import RefModel
    t = tables or ParamTables.default()
    w, opts = case_workload(name)
    forcing = synth.forcing_all(w)
    state = w.state.copy()
    diag = np.zeros((L.NDIAG, w.ncol), np.float32)
    fatal = RefModel(t, opts).run(state, forcing, w.static_f, w.static_i, w.dt, synth.ZSOIL, w.julian0, w.yearlen, w.zref,
                                  w.co2_frac, w.o2_frac, w.foln, diag=diag)
    return w, state, diag[:58], fatal


FP64_CASES = ["uniform", "global", "snow", "conus", "special_global", "special_snow"]


def run_reference_fp64(name, tables=None):
    """The same statements with every `real` a double (constants keep their fp32 values): the yardstick for how much of a
    difference is fp32 rounding noise (SURVEY 8c iv).  End state rounded to fp32 planes."""
    from oracle.refrun import RefModel
    t = tables or ParamTables.default()
    w, opts = case_workload(name)
    state = w.state.copy()
    fatal = RefModel(t, opts, fp64=True).run_native(state, synth.forcing_all(w), w.static_f, w.static_i, w.dt, synth.ZSOIL, w.julian0,
                                                    w.yearlen, w.zref, w.co2_frac, w.o2_frac, w.foln)
    return w, state, fatal


def main():
    out = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out, exist_ok=True)
    for name in FP64_CASES:
        w, state, fatal = run_reference_fp64(name)
        np.savez_compressed(os.path.join(out, f"ref64_{name}.npz"), state=state, fatal=fatal)
        print(f"{name:12s} fp64 twin: aborted columns {int(fatal.sum())}  TG mean {float(state[L.S['TG']].mean()):.3f}", flush=True)
    for name, workload, ncol, nsteps, over in CASES:
        t0 = time.time()
        w, state, diag, fatal = run_reference(name)
        np.savez_compressed(os.path.join(out, f"ref_{name}.npz"), state=state, diag=diag, fatal=fatal, opts=np.asarray(w.opts),
                            workload=workload, ncol=ncol, nsteps=nsteps)
        print(f"{name:12s} {workload:12s} {ncol:4d} x {nsteps:3d}  aborted columns {int(fatal.sum())}  TG mean "
              f"{float(state[L.S['TG']].mean()):.3f}  ({time.time() - t0:.1f} s)", flush=True)


if __name__ == "__main__":
    main()
