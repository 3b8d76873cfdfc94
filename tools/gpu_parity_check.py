#!/usr/bin/env python3
"""Compare the CUDA step kernel with the CPU oracle on a subsampled workload and print per-plane errors.
Usage: python tools/gpu_parity_check.py NAME NCOL NSTEPS [--fast] [opt=val ...]   (needs a GPU)"""
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

OPT_NAMES = ["veg", "crs", "btr", "run", "sfc", "frz", "inf", "rad", "alb", "snf", "tbot", "stc"]


def compare(name, a, b, is_int=False):
    if is_int:
        ai, bi = a.view(np.int32), b.view(np.int32)
        nd = int((ai != bi).sum())
        return f"{name:9s} int mismatches {nd}/{ai.size}"
    both_nan = np.isnan(a) & np.isnan(b)
    d = np.abs(a.astype(np.float64) - b.astype(np.float64))
    d[both_nan] = 0.0
    rel = d / np.maximum(np.maximum(np.abs(a), np.abs(b)).astype(np.float64), 1e-6)
    rel[both_nan] = 0.0
    nbad = int((~(rel <= 1e-4) & ~(d <= 1e-6)).sum())
    return (f"{name:9s} max|d| {np.nanmax(d):10.3e} maxrel {np.nanmax(rel):10.3e} p99.9rel {np.nanquantile(rel, 0.999):9.2e} "
            f"bitexact {float((a.view(np.int32) == b.view(np.int32)).mean()):6.4f} >1e-4: {nbad}")


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--") and "=" not in a]
    name, ncol, nsteps = args[0], int(args[1]), int(args[2])
    fast = "fast" if "--fast" in sys.argv else ("strict" if "--strict" in sys.argv else False)
    opts = list(synth.DEFAULT_OPTS)
    for a in sys.argv[1:]:
        if "=" in a:
            k, v = a.split("=")
            opts[OPT_NAMES.index(k)] = int(v)
    t = ParamTables.default()
    w = synth.make_workload(name, ncol=ncol, nsteps=nsteps, opts=opts)
    f = synth.forcing_all(w)
    st0 = w.state.copy()
    # --- oracle
    o = Oracle(t, opts, fast="strict" if fast == "strict" else False)
    st_cpu = st0.copy()
    dg_cpu = np.zeros((L.NDIAG, w.ncol), np.float32)
    t0 = time.time()
    nf_cpu = o.run_workload(w, f, dg_cpu, state=st_cpu)
    t_cpu = time.time() - t0
    # --- GPU, device-pointer entry
    eng = NoahMP(t, opts, fast=fast)
    print("kernel info:", eng.kernel_info())
    dev = torch.device("cuda:0")
    st = torch.from_numpy(st0.copy()).to(dev)
    fo = torch.from_numpy(f).to(dev)
    sf = torch.from_numpy(w.static_f).to(dev)
    si = torch.from_numpy(w.static_i).to(dev)
    dg = torch.zeros((L.NDIAG, w.ncol), dtype=torch.float32, device=dev)
    bud = torch.zeros(L.NBUDGET, dtype=torch.float64, device=dev)
    rp = eng.run_params(w.dt, synth.ZSOIL, w.julian0, w.yearlen, w.zref, w.co2_frac, w.o2_frac, w.foln)
    eng.run(rp, st, fo, sf, si, dg, bud)
    torch.cuda.synchronize()
    st_gpu = st.cpu().numpy()
    dg_gpu = dg.cpu().numpy()
    # timing of the plain (no diag) variant
    st2 = torch.from_numpy(st0.copy()).to(dev)
    eng.run(rp, st2, fo, sf, si)
    torch.cuda.synchronize()
    st2 = torch.from_numpy(st0.copy()).to(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    eng.run(rp, st2, fo, sf, si)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"{name} ncol={ncol} nsteps={nsteps} opts={opts} fast={fast}")
    print(f"CPU oracle {t_cpu:.3f}s = {ncol*nsteps/t_cpu:.3g} cell-steps/s ({o.num_threads()} thr); "
          f"GPU {ms:.3f} ms = {ncol*nsteps/ms*1e3:.3g} cell-steps/s; flagged cpu={nf_cpu} gpu={int(bud[1].item())}")
    same_plain = np.array_equal(st2.cpu().numpy()[:L.NSTATE_CORE].view(np.int32), st_gpu[:L.NSTATE_CORE].view(np.int32))
    print("plain-variant state identical to diag-variant state:", same_plain)
    print("--- state")
    nplanes = L.NSTATE if opts != synth.DEFAULT_OPTS else L.NSTATE_CORE
    for p in range(nplanes):
        print(" ", compare(L.STATE[p], st_cpu[p], st_gpu[p], p in L.INT_STATE_PLANES))
    print("--- diagnostics (last step)")
    for p in range(L.NDIAG):
        print(" ", compare(L.DIAG[p], dg_cpu[p], dg_gpu[p], p in L.INT_DIAG_PLANES))
    it_c, it_g = dg_cpu[L.D["ITER"]].view(np.int32), dg_gpu[L.D["ITER"]].view(np.int32)
    for nm, sh in (("loop1", 0), ("stom_sun", 8), ("stom_sha", 16), ("soil_substeps", 24)):
        a, b = (it_c >> sh) & 0xff, (it_g >> sh) & 0xff
        print(f"  iteration-count mismatches {nm}: {int((a != b).sum())}/{a.size}")
    print("budget:", {L.BUDGET[i]: float(bud[i].item()) for i in range(L.NBUDGET)})


if __name__ == "__main__":
    main()
