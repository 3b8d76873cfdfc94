#!/usr/bin/env python3
"""Time several builds of the step kernel on the same resident workload, in one process.

  python tools/tune_libs.py [--ncol N] [--reps R] name=/path/lib.so ...

Builds come from tools/tune_builds.sh (or any `nvcc ... -o x.so noahmp_abi.cu`).  The workload is the benchmark's
(conus1km, class-sorted, 24 model steps per launch, forcing generated on the device by the shipped fast build); every
build starts from the same state, 2 warm-up launches, R timed launches (CUDA events), best and median reported."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from noahmp_b200 import _lib, layout as L, synth  # noqa: E402
from noahmp_b200.engine import NoahMP  # noqa: E402
from noahmp_b200.parallel import divergence_order  # noqa: E402
from noahmp_b200.tables import ParamTables  # noqa: E402


def main():
    argv = sys.argv[1:]
    ncol, reps = 4_000_000, 4
    for flag in ("--ncol", "--reps"):
        if flag in argv:
            i = argv.index(flag)
            val = int(argv[i + 1])
            del argv[i:i + 2]
            if flag == "--ncol":
                ncol = val
            else:
                reps = val
    libs = [("shipped-fast", None)] + [tuple(a.split("=", 1)) for a in argv]
    t = ParamTables.default()
    w = synth.make_workload("conus1km", ncol=ncol, nsteps=24)
    order = divergence_order(w.static_i, w.static_f[L.C["LAT"]], 1)
    dev = torch.device("cuda", 0)
    state0 = torch.from_numpy(np.ascontiguousarray(w.state[:, order])).to(dev)
    sf = torch.from_numpy(np.ascontiguousarray(w.static_f[:, order])).to(dev)
    si = torch.from_numpy(np.ascontiguousarray(w.static_i[order])).to(dev)
    sincos = torch.from_numpy(np.ascontiguousarray(w.sincos[:, order])).to(dev)
    forcing = torch.empty((24, L.NFORCING, ncol), dtype=torch.float32, device=dev)
    eng = NoahMP(t, w.opts, fast="fast")
    coef = synth.synth_params_array(w.synth)
    for s in range(24):
        sc = synth.step_scalars(w, s)
        eng.synth_forcing(ncol, 0, w.synth.seed, s, coef, np.asarray([sc["sd"], sc["cdch"], sc["t_diur"], sc["t_seas"]], np.float32),
                          sincos, forcing[s])
    torch.cuda.synchronize()
    eng.close()
    ref_state = None
    print(f"{'build':<16}{'regs':>5}{'lmem':>6}{'best cs/s':>12}{'median':>12}  state==shipped")
    for name, path in libs:
        if path:
            os.environ["NOAHMP_B200_LIB"] = os.path.abspath(path)
        else:
            os.environ.pop("NOAHMP_B200_LIB", None)
        _lib._cache.clear()
        eng = NoahMP(t, w.opts, fast="fast")
        info = eng.kernel_info()
        state = state0.clone()
        rp = lambda k: eng.run_params(w.dt, synth.ZSOIL, w.julian0 + k * 24 * w.dt / 86400.0, w.yearlen, w.zref, w.co2_frac,  # noqa: E731
                                      w.o2_frac, w.foln)
        for k in range(2):
            eng.run(rp(k), state, forcing, sf, si)
        torch.cuda.synchronize()
        ms = []
        for k in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            eng.run(rp(2 + k), state, forcing, sf, si)
            e1.record()
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        same = ""
        if ref_state is None:
            ref_state = state.clone()
        else:
            same = f"{(state.view(torch.int32) == ref_state.view(torch.int32)).all(dim=0).float().mean().item():.4f}"
        cs = ncol * 24 / (np.asarray(ms) * 1e-3)
        print(f"{name:<16}{info['registers']:>5}{info['local_bytes']:>6}{cs.max():>12.4g}{np.median(cs):>12.4g}  {same}", flush=True)
        eng.close()


if __name__ == "__main__":
    main()
