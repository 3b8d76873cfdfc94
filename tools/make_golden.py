#!/usr/bin/env python3
"""Generate tests/golden/*.npz: end states + last-step diagnostics of the CPU oracle (strict-libm build, whose
results do not depend on the host's libm version) for small slices of every workload and several option sets.

The reference ships no golden vectors and cannot be run, so these are REGRESSION pins of
this repo's oracle, not outputs of the Fortran.  Re-generate only when an oracle fix is intended:
    python tools/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from noahmp_b200 import layout as L, synth  # noqa: E402
from noahmp_b200.tables import ParamTables  # noqa: E402
from oracle.oracle import Oracle  # noqa: E402

CASES = [  # name, workload, ncol, nsteps, option overrides
    ("single_default", "single", 1, 96, {}),
    ("uniform_default", "uniform100k", 64, 48, {}),
    ("global_default", "global025", 64, 72, {}),
    ("snow_default", "snow1m", 64, 72, {}),
    ("global_noah_options", "global025", 64, 48, {"sfc": 2, "run": 3, "frz": 2, "inf": 2, "alb": 1, "tbot": 2, "stc": 2, "snf": 3}),
    ("snow_simtop_jarvis", "snow1m", 64, 48, {"run": 2, "crs": 2, "btr": 2, "rad": 3, "snf": 2}),
    ("global_dynveg_bats", "global025", 64, 48, {"veg": 2, "run": 4, "btr": 3, "rad": 2}),
]
OPT_NAMES = ["veg", "crs", "btr", "run", "sfc", "frz", "inf", "rad", "alb", "snf", "tbot", "stc"]


def options(over):
    o = list(synth.DEFAULT_OPTS)
    for k, v in over.items():
        o[OPT_NAMES.index(k)] = v
    return o


def run_case(workload, ncol, nsteps, over, variant="strict"):
    opts = options(over)
    w = synth.make_workload(workload, ncol=ncol, nsteps=nsteps, opts=opts)
    forcing = synth.forcing_all(w)
    diag = np.zeros((L.NDIAG, w.ncol), np.float32)
    state = w.state.copy()
    nflag = Oracle(ParamTables.default(), opts, fast=variant).run_workload(w, forcing, diag, state=state)
    return w, state, diag, nflag


def main():
    out = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out, exist_ok=True)
    for name, workload, ncol, nsteps, over in CASES:
        w, state, diag, nflag = run_case(workload, ncol, nsteps, over)
        np.savez_compressed(os.path.join(out, f"oracle_{name}.npz"), state=state, diag=diag, opts=np.asarray(w.opts),
                            nflag=nflag, workload=workload, ncol=ncol, nsteps=nsteps)
        print(name, "flagged", nflag, "TV", float(state[L.S["TV"]].mean()), "nan", int(np.isnan(state).sum()))


if __name__ == "__main__":
    main()
