#!/usr/bin/env python3
"""BASELINE configs[2] and configs[4] at FULL size on one GPU, forcing generated on the device chunk by chunk, with a
strided column sample carried through the oracle for the whole run ("parity on a 4 096-column subsample run
through the oracle for the full year; throughput on the full set").

  python tools/gpu_full_configs.py [global025] [snow1m] [--sample N]

Prints, per config and build: wall time of the whole run (forcing generation included), step-kernel-only throughput,
closure flags, and the agreement of the sampled columns with the oracle (strict build vs strict oracle: bit-identical
fraction; fast build vs plain oracle: fraction of columns inside the 1e-4 tolerance per prognostic group)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from noahmp_b200 import layout as L, synth  # noqa: E402
from noahmp_b200.engine import NoahMP  # noqa: E402
from noahmp_b200.tables import ParamTables  # noqa: E402
from oracle.oracle import Oracle  # noqa: E402

NOMINAL_STEPS = {"global025": 8760, "snow1m": 720}
CHUNK = 24


def sample_of(w, n):
    import copy
    idx = (np.arange(n, dtype=np.int64) * w.ncol) // n
    s = copy.copy(w)
    s.ncol = n
    s.state = np.ascontiguousarray(w.state[:, idx])
    s.static_f = np.ascontiguousarray(w.static_f[:, idx])
    s.static_i = np.ascontiguousarray(w.static_i[idx])
    s.sincos = np.ascontiguousarray(w.sincos[:, idx])
    return s, idx


def gpu_run(t, w, variant):
    dev = torch.device("cuda", 0)
    eng = NoahMP(t, w.opts, fast=variant)
    st = torch.from_numpy(w.state.copy()).to(dev)
    sf, si = torch.from_numpy(w.static_f).to(dev), torch.from_numpy(w.static_i).to(dev)
    sincos = torch.from_numpy(w.sincos).to(dev)
    fo = torch.empty((CHUNK, L.NFORCING, w.ncol), dtype=torch.float32, device=dev)
    bd = torch.zeros(L.NBUDGET, dtype=torch.float64, device=dev)
    coef = synth.synth_params_array(w.synth)
    scal = [np.asarray([sc["sd"], sc["cdch"], sc["t_diur"], sc["t_seas"]], np.float32)
            for sc in (synth.step_scalars(w, s) for s in range(w.nsteps))]
    evs = []
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for s0 in range(0, w.nsteps, CHUNK):
        n = min(CHUNK, w.nsteps - s0)
        for s in range(n):
            eng.synth_forcing(w.ncol, 0, w.synth.seed, s0 + s, coef, scal[s0 + s], sincos, fo[s])
        rp = eng.run_params(w.dt, synth.ZSOIL, w.julian0 + s0 * w.dt / 86400.0, w.yearlen, w.zref, w.co2_frac, w.o2_frac, w.foln)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        eng.run(rp, st, fo[:n], sf, si, None, bd)
        e1.record()
        evs.append((e0, e1))
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    kern = sum(a.elapsed_time(b) for a, b in evs) * 1e-3
    out = st.cpu().numpy(), bd.cpu().numpy(), wall, kern, eng.launch_count()
    eng.close()
    return out


def main():
    argv = sys.argv[1:]
    nsample = 4096
    if "--sample" in argv:
        i = argv.index("--sample")
        nsample = int(argv[i + 1])
        del argv[i:i + 2]
    names = argv or ["global025", "snow1m"]
    t = ParamTables.default()
    for name in names:
        w = synth.make_workload(name, nsteps=NOMINAL_STEPS[name])
        print(f"== {name}: {w.ncol} columns x {w.nsteps} steps of {w.dt:.0f} s = {w.ncol * w.nsteps:.3g} cell-steps", flush=True)
        ws, idx = sample_of(w, min(nsample, w.ncol))
        t0 = time.perf_counter()
        forcing = synth.forcing_all(ws, cols=idx)
        ref = {}
        for v in ("strict", "plain"):
            st = ws.state.copy()
            nf = Oracle(t, w.opts, fast="strict" if v == "strict" else False).run_workload(ws, forcing, state=st)
            ref[v] = st
            print(f"   oracle {v}: {ws.ncol} sampled columns, flagged {nf}", flush=True)
        print(f"   (oracle legs incl. host forcing: {time.perf_counter() - t0:.1f} s)")
        for v in ("fast", "strict"):
            g, b, wall, kern, launches = gpu_run(t, w, v)
            cs = w.ncol * w.nsteps
            r = ref["strict"] if v == "strict" else ref["plain"]
            gs = g[:, idx]
            same = (gs[:L.NSTATE_CORE].view(np.int32) == r[:L.NSTATE_CORE].view(np.int32)).all(axis=0).mean()
            isn = (gs[L.S["ISNOW"]].view(np.int32) == r[L.S["ISNOW"]].view(np.int32)).mean()
            line = [f"   GPU {v}: whole run {wall:.2f} s ({cs / wall:.3g} cell-steps/s incl. forcing generation, {launches} launches), "
                    f"step kernel {kern:.2f} s ({cs / kern:.3g}), flagged column-launches {int(b[L.B['NFLAGGED']])}, "
                    f"max|ERRWAT| {b[L.B['ERRWAT_MAXABS']]:.4f} mm; sample: bit-identical {same:.4f}, ISNOW equal {isn:.4f}"]
            snow = ["SNICE-2", "SNICE-1", "SNICE+0", "STC+0"] if name == "snow1m" else []
            for p in ["TG", "STC+1", "STC+4", "SMC1", "SMC4", "WA", "SNEQV", "SNOWH"] + snow:
                if p not in L.S:
                    continue
                a, c = gs[L.S[p]].astype(np.float64), r[L.S[p]].astype(np.float64)
                d = np.abs(a - c)
                within = ((d <= 1e-4 * np.maximum(np.abs(a), np.abs(c))) | (d <= 1e-6)).mean()
                line.append(f"{p}: within 1e-4 {within:.3f}, mean diff {a.mean() - c.mean():+.2e}")
            print("\n      ".join(line), flush=True)


if __name__ == "__main__":
    main()
