#!/usr/bin/env python3
"""One-year run of a global-0.25-degree subsample on the GPU (strict, parity, fast builds) against the strict and
plain oracles; prints the agreement figures quoted in DESIGN.md.
Usage: python tools/gpu_year_run.py [NCOL] [name=/abs/path/to/experimental_build.so ...]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from noahmp_b200 import _lib, layout as L, synth  # noqa: E402
from noahmp_b200.engine import NoahMP  # noqa: E402
from noahmp_b200.tables import ParamTables  # noqa: E402
from oracle.oracle import Oracle  # noqa: E402


def main():
    ncol = int(sys.argv[1]) if len(sys.argv) > 1 and "=" not in sys.argv[1] else 4096
    extras = dict(a.split("=", 1) for a in sys.argv[1:] if "=" in a)
    nsteps = 8760
    t = ParamTables.default()
    w = synth.make_workload("global025", ncol=ncol, nsteps=nsteps)
    forcing = synth.forcing_all(w)
    ref = {}
    for v in ("strict", "parity"):
        st = w.state.copy()
        t0 = time.time()
        nf = Oracle(t, w.opts, fast=v if v == "strict" else False).run_workload(w, forcing, state=st)
        ref[v] = st
        print(f"oracle {v}: {time.time() - t0:.1f} s, flagged {nf}")
    fo = torch.from_numpy(forcing).cuda()
    sf, si = torch.from_numpy(w.static_f).cuda(), torch.from_numpy(w.static_i).cuda()
    for v in ("strict", "parity", "fast") + tuple(extras):
        if v in extras:   # an experimental build, compared with the plain oracle like `fast`
            os.environ["NOAHMP_B200_LIB"] = os.path.abspath(extras[v])
            _lib._cache.clear()
        eng = NoahMP(t, w.opts, fast="fast" if v in extras else v)
        st = torch.from_numpy(w.state.copy()).cuda()
        bd = torch.zeros(L.NBUDGET, dtype=torch.float64, device="cuda")
        torch.cuda.synchronize()
        t0 = time.time()
        for s0 in range(0, nsteps, 24):
            rp = eng.run_params(w.dt, synth.ZSOIL, w.julian0 + s0 * w.dt / 86400.0, w.yearlen, w.zref, w.co2_frac, w.o2_frac, w.foln)
            eng.run(rp, st, fo[s0:s0 + 24], sf, si, None, bd)
        torch.cuda.synchronize()
        el = time.time() - t0
        g = st.cpu().numpy()
        b = bd.cpu().numpy()
        r = ref["strict"] if v == "strict" else ref["parity"]
        same = (g[:L.NSTATE_CORE].view(np.int32) == r[:L.NSTATE_CORE].view(np.int32)).all(axis=0).mean()
        isn = (g[L.S["ISNOW"]].view(np.int32) == r[L.S["ISNOW"]].view(np.int32)).mean()
        line = [f"GPU {v}: {el:.2f} s ({ncol * nsteps / el:.3g} cell-steps/s), flagged {int(b[L.B['NFLAGGED']])}, "
                f"max|ERRWAT| {b[L.B['ERRWAT_MAXABS']]:.4f} mm, columns bit-identical {same:.4f}, ISNOW equal {isn:.4f}"]
        for p in ("TG", "STC+1", "STC+4", "SMC1", "SMC4", "WA", "SNEQV"):
            a, c = g[L.S[p]].astype(np.float64), r[L.S[p]].astype(np.float64)
            d = np.abs(a - c)
            within = ((d <= 1e-4 * np.maximum(np.abs(a), np.abs(c))) | (d <= 1e-6)).mean()
            line.append(f"{p}: mean diff {a.mean() - c.mean():+.2e}, p95|d| {np.quantile(d, 0.95):.2e}, within 1e-4: {within:.3f}")
        print("\n   ".join(line))
        eng.close()


if __name__ == "__main__":
    main()
