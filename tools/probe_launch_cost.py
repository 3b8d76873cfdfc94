#!/usr/bin/env python3
"""What a short launch costs: the step kernel on RESIDENT inputs, 24 model steps as one launch, as 24 one-step launches
(forcing planes / records, with and without the budget sums) and as 6 four-step launches at the reference namelist's
cadence (900 s steps between hourly records).  Explains the distance between the device-timed figure (one 24-step launch)
and the end-to-end figure (noahmp_advance_records_host launches once per record).

  python tools/probe_launch_cost.py [--ncol N]          # gpurun; prints a table, cell-steps/s from CUDA events"""
import copy
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from noahmp_b200 import layout as L, synth  # noqa: E402
from noahmp_b200.engine import NoahMP  # noqa: E402
from noahmp_b200.parallel import divergence_order  # noqa: E402
from noahmp_b200.tables import ParamTables  # noqa: E402


def main():
    ncol = int(sys.argv[sys.argv.index("--ncol") + 1]) if "--ncol" in sys.argv else 4_000_000
    t = ParamTables.default()
    w = synth.make_workload("conus1km", ncol=ncol, nsteps=24)
    order = divergence_order(w.static_i, w.static_f[L.C["LAT"]], 1)
    dev = torch.device("cuda", 0)
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
    state0, sf, si, sincos = d(w.state[:, order]), d(w.static_f[:, order]), d(w.static_i[order]), d(w.sincos[:, order])
    ids = d(order.astype(np.int64))      # the generator hashes a column's id in ITS numbering
    eng = NoahMP(t, w.opts, fast="fast")
    forcing = torch.empty((24, L.NFORCING, ncol), dtype=torch.float32, device=dev)
    coef = synth.synth_params_array(w.synth)
    scal = [synth.step_scalars(w, s) for s in range(24)]
    for s, sc in enumerate(scal):
        eng.synth_forcing_ids(ids, w.synth.seed, s, coef, np.asarray([sc["sd"], sc["cdch"], sc["t_diur"], sc["t_seas"]], np.float32),
                              sincos, forcing[s])
    recs = forcing[:, :8].contiguous()
    geo = torch.stack([sincos[0], sincos[1], torch.zeros_like(sincos[0]), torch.ones_like(sincos[0])]).contiguous()
    budget = torch.zeros(L.NBUDGET, dtype=torch.float64, device=dev)
    rp = lambda dt, j0: eng.run_params(dt, synth.ZSOIL, j0, w.yearlen, w.zref, w.co2_frac, w.o2_frac, w.foln)  # noqa: E731
    tab1 = eng.step_table(list(range(24)), list(range(24)), [0.0] * 24, [(sc["sd"], 1.0, sc["cdch"], 0.0) for sc in scal])
    w900 = copy.copy(w)
    w900.dt = 900.0
    sc900 = [synth.step_scalars(w900, s) for s in range(96)]
    tab4 = eng.step_table([s // 4 for s in range(96)], [(s // 4 + 1) % 24 for s in range(96)], [(s % 4) / 4 for s in range(96)],
                          [(sc["sd"], 1.0, sc["cdch"], 0.0) for sc in sc900])

    def sub(tab, lo, hi):      # a slice of a ctypes array is a list: build the sub-table the C ABI takes
        return (type(tab[0]) * (hi - lo))(*tab[lo:hi])

    def day(j0):
        return w.julian0 + j0

    def one_launch(st, bud):
        eng.run(rp(3600.0, day(0)), st, forcing, sf, si, None, bud)

    def one_step_launches(st, bud):
        for s in range(24):
            eng.run(rp(3600.0, day(s / 24.0)), st, forcing[s:s + 1], sf, si, None, bud)

    def one_step_records(st, bud):
        for s in range(24):
            eng.run_records(rp(3600.0, day(s / 24.0)), st, recs, sub(tab1, s, s + 1), geo, sf, si, None, bud)

    def records_24(st, bud):
        eng.run_records(rp(3600.0, day(0)), st, recs, tab1, geo, sf, si, None, bud)

    def four_step_records_900(st, bud):
        for k in range(24):
            eng.run_records(rp(900.0, day(k / 24.0)), st, recs, sub(tab4, 4 * k, 4 * k + 4), geo, sf, si, None, bud)

    def steps_900_one_launch_each_32(st, bud):
        for k in range(3):
            eng.run_records(rp(900.0, day(k / 3.0)), st, recs, sub(tab4, 32 * k, 32 * k + 32), geo, sf, si, None, bud)

    cases = [("forcing planes, 1 launch x 24 steps", one_launch, None, 24), ("same + budget", one_launch, budget, 24),
             ("forcing planes, 24 launches x 1 step", one_step_launches, None, 24), ("same + budget", one_step_launches, budget, 24),
             ("records, 1 launch x 24 steps + budget", records_24, budget, 24),
             ("records, 24 launches x 1 step", one_step_records, None, 24), ("same + budget", one_step_records, budget, 24),
             ("records dt=900, 24 launches x 4 steps + budget", four_step_records_900, budget, 96),
             ("records dt=900, 3 launches x 32 steps + budget", steps_900_one_launch_each_32, budget, 96)]
    print(f"{ncol} class-sorted CONUS columns, fast build, resident inputs (CUDA events, best of 3)")
    print(f"{'case':<52}{'ms per model day':>18}{'cell-steps/s':>14}{'ms per launch':>15}")
    for name, fn, bud, nst in cases:
        st = state0.clone()
        fn(st, bud)
        torch.cuda.synchronize()
        best = 1e30
        for _ in range(3):
            st.copy_(state0)
            l0 = eng.launch_count()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn(st, bud)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
            nl = eng.launch_count() - l0
        print(f"{name:<52}{best:>18.2f}{ncol * nst / (best * 1e-3):>14.4g}{best / nl:>15.3f}", flush=True)
    eng.close()


if __name__ == "__main__":
    main()
