#!/usr/bin/env python3
"""Bit-compare the end state of two builds of the library on the same columns (TEST INFRASTRUCTURE for kernel experiments).

  python tools/compare_libs.py VARIANT /path/a.so /path/b.so        # VARIANT = the engine variant both stand in for"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from noahmp_b200 import _lib, layout as L, synth  # noqa: E402
from noahmp_b200.engine import NoahMP  # noqa: E402
from noahmp_b200.tables import ParamTables  # noqa: E402


def main():
    variant, liba, libb = sys.argv[1:4]
    t = ParamTables.default()
    for wl, ncol, nsteps in (("conus1km", 300_000, 24), ("global025", 100_000, 48), ("snow1m", 100_000, 48)):
        w = synth.make_workload(wl, ncol=ncol, nsteps=nsteps)
        f = torch.from_numpy(synth.forcing_all(w)).cuda()
        out = []
        for lib in (liba, libb):
            os.environ["NOAHMP_B200_LIB"] = os.path.abspath(lib)
            _lib._cache.clear()
            eng = NoahMP(t, w.opts, fast=variant)
            st = torch.from_numpy(w.state.copy()).cuda()
            eng.run(eng.run_params(w.dt, synth.ZSOIL, w.julian0, w.yearlen, w.zref, w.co2_frac, w.o2_frac, w.foln), st, f,
                    torch.from_numpy(w.static_f).cuda(), torch.from_numpy(w.static_i).cuda())
            torch.cuda.synchronize()
            out.append(st.cpu().numpy())
            eng.close()
        same = (out[0].view(np.int32) == out[1].view(np.int32)).all(axis=0)
        print(f"{wl}: {ncol} columns x {nsteps} steps, columns bit-identical {same.mean():.6f}", flush=True)


if __name__ == "__main__":
    main()
