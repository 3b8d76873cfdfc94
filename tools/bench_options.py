#!/usr/bin/env python3
"""Throughput of the fast build for non-default option sets (the run-time-options kernel, 57-plane state) beside the
shipped option set (compile-time options, 44-plane state): same workload and harness as tools/tune_libs.py.

  python tools/bench_options.py [--ncol N]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from noahmp_b200 import layout as L, synth  # noqa: E402
from noahmp_b200.engine import NoahMP, OPTION_ORDER  # noqa: E402
from noahmp_b200.parallel import divergence_order  # noqa: E402
from noahmp_b200.tables import ParamTables  # noqa: E402

CASES = [("shipped options", {}), ("opt_tbot=2 (same physics cost, run-time options)", {"opt_tbot": 2}),
         ("opt_sfc=2 (Chen97)", {"opt_sfc": 2}), ("opt_crs=2 (Jarvis)", {"opt_crs": 2}), ("opt_run=3 (Schaake)", {"opt_run": 3}),
         ("opt_frz=2 opt_inf=2 (Koren99)", {"opt_frz": 2, "opt_inf": 2}), ("opt_veg=2 (dynamic vegetation)", {"opt_veg": 2}),
         ("opt_alb=1 opt_rad=3 opt_stc=2", {"opt_alb": 1, "opt_rad": 3, "opt_stc": 2})]


def main():
    ncol = int(sys.argv[sys.argv.index("--ncol") + 1]) if "--ncol" in sys.argv else 4_000_000
    t = ParamTables.default()
    w = synth.make_workload("conus1km", ncol=ncol, nsteps=24)
    order = divergence_order(w.static_i, w.static_f[L.C["LAT"]], 1)
    dev = torch.device("cuda", 0)
    state0 = torch.from_numpy(np.ascontiguousarray(w.state[:, order])).to(dev)
    sf = torch.from_numpy(np.ascontiguousarray(w.static_f[:, order])).to(dev)
    si = torch.from_numpy(np.ascontiguousarray(w.static_i[order])).to(dev)
    forcing = torch.from_numpy(np.ascontiguousarray(synth.forcing_all(w)[:, :, order])).to(dev)
    print(f"{'options':<52}{'regs':>5}{'lmem':>6}{'cell-steps/s':>14}  flagged")
    for name, over in CASES:
        opts = dict(zip(OPTION_ORDER, w.opts))
        opts.update(over)
        eng = NoahMP(t, opts, fast="fast")
        state = state0.clone()
        bud = torch.zeros(L.NBUDGET, dtype=torch.float64, device=dev)
        rp = eng.run_params(w.dt, synth.ZSOIL, w.julian0, w.yearlen, w.zref, w.co2_frac, w.o2_frac, w.foln)
        eng.run(rp, state, forcing, sf, si)
        torch.cuda.synchronize()
        best = 1e30
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            eng.run(rp, state, forcing, sf, si)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        eng.run(rp, state, forcing, sf, si, None, bud)
        torch.cuda.synchronize()
        info = eng.kernel_info()
        print(f"{name:<52}{info['registers']:>5}{info['local_bytes']:>6}{ncol * 24 / (best * 1e-3):>14.4g}  {int(bud[L.B['NFLAGGED']].item())}",
              flush=True)
        eng.close()


if __name__ == "__main__":
    main()
