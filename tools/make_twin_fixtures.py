#!/usr/bin/env python3
"""Generate tests/golden/twin_*.npz: FULL-LENGTH end states of the reference's own statements, in fp32 and in fp64.

Needs /root/reference (read-only) - run here, in the build container; the fixtures travel.  For every case the synthetic
inputs are rebuilt from their seeds at the config's stated size (noahmp_b200/synth.py), a strided sample of columns is taken
(columns are independent, func:66 is a one-column routine) and stepped through the mechanical Fortran -> C++ translation
of core/module_noahmp_func.f90 (oracle/f90_to_cpp.py) twice: once as written (`real` = fp32: what the reference computes)
and once with every `real` a double (the fp64 twin: what the reference's statements mean).  The distance of the fp32
reference from its twin is the reference's OWN rounding noise after that many steps; tests/test_accuracy_gpu.py demands
that the GPU builds are no further from the twin than a stated multiple of it, on all 44 live state planes.

    python tools/make_twin_fixtures.py [case ...]         # all cases: about two minutes on 8 cores
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from noahmp_b200 import layout as L, synth  # noqa: E402
from noahmp_b200.tables import ParamTables  # noqa: E402

# name: (workload, columns of the config, model steps, sampled columns kept in the fixture)
#   uniform100k    BASELINE configs[1] at its stated 100 000 x 48
#   global025_year BASELINE configs[2]: one year of hourly steps on a parity subsample (the config's wording)
#   snow1m_month   BASELINE configs[4]: 720 melt / refreeze steps of the snow ensemble
#   conus1km_day   BASELINE configs[3]: the benchmark workload, 24 steps (bench.py carries its figures in `parity`)
CASES = {
    "uniform100k": ("uniform100k", 100_000, 48, 8192),
    "global025_year": ("global025", 280_000, 8760, 4096),
    "snow1m_month": ("snow1m", 1_000_000, 720, 8192),
    "conus1km_day": ("conus1km", 17_000_000, 24, 8192),
}


def sample_ids(ncol: int, keep: int) -> np.ndarray:
    """Generator column ids of the fixture's sample: every (ncol/keep)-th column."""
    return (np.arange(keep, dtype=np.int64) * ncol) // keep


def case_inputs(name: str):
    """(sampled workload, its generator column ids, forcing [nsteps][9][keep]) of a case."""
    import copy
    workload, ncol, nsteps, keep = CASES[name]
    w = synth.make_workload(workload, ncol=ncol, nsteps=nsteps)
    idx = sample_ids(ncol, keep)
    s = copy.copy(w)
    s.ncol = keep
    s.state = np.ascontiguousarray(w.state[:, idx])
    s.static_f = np.ascontiguousarray(w.static_f[:, idx])
    s.static_i = np.ascontiguousarray(w.static_i[idx])
    s.sincos = np.ascontiguousarray(w.sincos[:, idx])
    return s, idx, synth.forcing_all(s, cols=idx)


def run_case(name: str, nthreads: int = 0):
    from oracle.refrun import RefModel
    t = ParamTables.default()
    w, idx, forcing = case_inputs(name)
    out = {}
    for tag, fp64 in (("ref32", False), ("ref64", True)):
        st = w.state.copy()
        t0 = time.time()
        fatal = RefModel(t, w.opts, fp64=fp64).run_native(st, forcing, w.static_f, w.static_i, w.dt, synth.ZSOIL, w.julian0, w.yearlen,
                                                          w.zref, w.co2_frac, w.o2_frac, w.foln, nthreads=nthreads)
        out[tag] = st[:L.NSTATE_CORE].copy()
        out[tag + "_fatal"] = fatal
        print(f"{name:16s} {tag}: {w.ncol} columns x {w.nsteps} steps in {time.time() - t0:.1f} s, aborted columns {int(fatal.sum())}",
              flush=True)
    return w, idx, out


def main():
    names = sys.argv[1:] or list(CASES)
    dst = os.path.join(ROOT, "tests", "golden")
    for name in names:
        w, idx, out = run_case(name)
        workload, ncol, nsteps, keep = CASES[name]
        np.savez_compressed(os.path.join(dst, f"twin_{name}.npz"), workload=workload, ncol=ncol, nsteps=nsteps, columns=idx,
                            opts=np.asarray(w.opts), **out)


if __name__ == "__main__":
    main()
