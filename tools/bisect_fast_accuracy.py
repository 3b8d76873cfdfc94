#!/usr/bin/env python3
"""Which compiler flag or algebra site costs the fast build its accuracy on a given plane?  (TEST INFRASTRUCTURE.)

  python tools/bisect_fast_accuracy.py [--workload snow1m --ncol 8192 --nsteps 96] name=/path/lib.so ...

Every library (built by tools/tune_builds.sh from the same source with other flags) advances the same columns; the final
state is compared with the plain CPU oracle's plane by plane: fraction of columns inside rel 1e-4 / abs 1e-6 (the north
star's tolerance), beside the same figure for the oracle rebuilt with FMA contraction (a second valid CPU evaluation)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from noahmp_b200 import _lib, layout as L, synth  # noqa: E402
from noahmp_b200.engine import NoahMP  # noqa: E402
from noahmp_b200.tables import ParamTables  # noqa: E402
from oracle.oracle import Oracle  # noqa: E402


def frac_within(a, b, rel=1e-4, floor=1e-6):
    d = np.abs(a.astype(np.float64) - b.astype(np.float64))
    return float(((d <= rel * np.maximum(np.abs(a), np.abs(b))) | (d <= floor)).mean())


def main():
    argv = sys.argv[1:]
    opt = {"--workload": "snow1m", "--ncol": "8192", "--nsteps": "96"}
    for k in list(opt):
        if k in argv:
            i = argv.index(k)
            opt[k] = argv[i + 1]
            del argv[i:i + 2]
    libs = [("shipped-fast", None)] + [tuple(a.split("=", 1)) for a in argv]
    t = ParamTables.default()
    w = synth.make_workload(opt["--workload"], ncol=int(opt["--ncol"]), nsteps=int(opt["--nsteps"]))
    forcing = synth.forcing_all(w)
    ref = w.state.copy()
    Oracle(t, w.opts, fast=False).run_workload(w, forcing, state=ref)
    fma = w.state.copy()
    Oracle(t, w.opts, fast="fast").run_workload(w, forcing, state=fma)
    planes = [L.STATE[p] for p in range(L.NSTATE_CORE) if p != L.S["ISNOW"]]
    isn = lambda s: np.ascontiguousarray(s[L.S["ISNOW"]]).view(np.int32)  # noqa: E731
    rows = {"oracle-FMA": fma}
    dev = torch.device("cuda", 0)
    for name, path in libs:
        if path:
            os.environ["NOAHMP_B200_LIB"] = os.path.abspath(path)
        else:
            os.environ.pop("NOAHMP_B200_LIB", None)
        _lib._cache.clear()
        eng = NoahMP(t, w.opts, fast="fast")
        st = torch.from_numpy(w.state.copy()).to(dev)
        eng.run(eng.run_params(w.dt, synth.ZSOIL, w.julian0, w.yearlen, w.zref, w.co2_frac, w.o2_frac, w.foln), st,
                torch.from_numpy(forcing).to(dev), torch.from_numpy(w.static_f).to(dev), torch.from_numpy(w.static_i).to(dev))
        torch.cuda.synchronize()
        rows[name] = st.cpu().numpy()
        eng.close()
    show = [p for p in planes if min(frac_within(ref[L.S[p]], s[L.S[p]]) for s in rows.values()) < 0.995]
    print(f"{opt['--workload']} {w.ncol} columns x {w.nsteps} steps; fraction of columns inside the tolerance, planes where any build is below 0.995")
    print(f"{'build':<22}{'ISNOW==':>8}" + "".join(f"{p:>9}" for p in show) + f"{'min(all)':>10}")
    for name, s in rows.items():
        same = isn(s) == isn(ref)
        fr = {p: frac_within(ref[L.S[p]][same], s[L.S[p]][same]) for p in planes}
        print(f"{name:<22}{same.mean():>8.4f}" + "".join(f"{fr[p]:>9.4f}" for p in show) + f"{min(fr.values()):>10.4f}", flush=True)


if __name__ == "__main__":
    main()
