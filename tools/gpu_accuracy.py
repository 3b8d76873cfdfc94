#!/usr/bin/env python3
"""Full-length accuracy of the GPU builds against the fp64 twin of the reference (tests/golden/twin_*.npz).

    python tools/gpu_accuracy.py [--cases a,b] [--variants fast,parity,strict] [--json out.json] [name=/abs/lib.so ...]

For every case of tools/make_twin_fixtures.py the sampled columns are run on the GPU for the config's full step count
(uniform100k runs ALL 100 000 columns of BASELINE configs[1] and the sample is picked out of the result) and every one of
the 44 live state planes is set against the twin in tolerance units, beside the fp32 reference's own distance
(tools/accuracy.py).  `name=lib.so` adds experimental builds (used to bisect --use_fast_math sub-flags and the
NMP_FAST_ALGEBRA sites).  TEST INFRASTRUCTURE: the fixtures are outputs of the translated reference; nothing here is product.
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import accuracy as A  # noqa: E402
import make_twin_fixtures as M  # noqa: E402
from noahmp_b200 import _lib, layout as L, synth  # noqa: E402
from noahmp_b200.engine import NoahMP  # noqa: E402
from noahmp_b200.tables import ParamTables  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def fixture(name):
    return np.load(os.path.join(GOLDEN, f"twin_{name}.npz"))


def gpu_advance(tables, w, forcing, variant, chunk=24, lib=None, col0=0, device_forcing=None):
    """Advance w.state by all steps of `forcing` ([nsteps][9][ncol] host array) in launches of `chunk` fused steps; returns the
    end state, the flagged column-launch count and the kernel seconds."""
    if lib:
        os.environ["NOAHMP_B200_LIB"] = os.path.abspath(lib)
        _lib._cache.clear()
    try:
        eng = NoahMP(tables, w.opts, device=0, fast=variant if not lib else "fast")
    finally:
        if lib:
            os.environ.pop("NOAHMP_B200_LIB", None)
            _lib._cache.clear()
    dev = "cuda:0"
    st = torch.from_numpy(w.state.copy()).to(dev)
    sf, si = torch.from_numpy(w.static_f).to(dev), torch.from_numpy(w.static_i).to(dev)
    bd = torch.zeros(L.NBUDGET, dtype=torch.float64, device=dev)
    nsteps = forcing.shape[0]
    fo = torch.from_numpy(forcing).to(dev) if forcing.nbytes < 8e9 else None
    torch.cuda.synchronize()
    t0 = time.time()
    for s0 in range(0, nsteps, chunk):
        n = min(chunk, nsteps - s0)
        blk = fo[s0:s0 + n] if fo is not None else torch.from_numpy(forcing[s0:s0 + n]).to(dev)
        rp = eng.run_params(w.dt, synth.ZSOIL, w.julian0 + s0 * w.dt / 86400.0, w.yearlen, w.zref, w.co2_frac, w.o2_frac, w.foln)
        eng.run(rp, st, blk, sf, si, None, bd)
    torch.cuda.synchronize()
    el = time.time() - t0
    out = st.cpu().numpy()
    flagged = int(bd[L.B["NFLAGGED"]].item())
    eng.close()
    return out, flagged, el


def run_case(name, variant, tables=None, lib=None):
    """-> (comparison dict of tools/accuracy.compare, flagged count, seconds).  The GPU runs the sample's columns for the
    full step count; for uniform100k it runs the whole 100 000-column config and the sample is taken from the result."""
    tables = tables or ParamTables.default()
    g = fixture(name)
    workload, ncol, nsteps, keep = M.CASES[name]
    idx = g["columns"]
    if name == "uniform100k":
        w = synth.make_workload(workload, ncol=ncol, nsteps=nsteps)
        st, flagged, el = gpu_advance(tables, w, synth.forcing_all(w), variant, lib=lib)
        st = st[:, idx]
    else:
        w, idx2, forcing = M.case_inputs(name)
        assert np.array_equal(idx, idx2)
        st, flagged, el = gpu_advance(tables, w, forcing, variant, lib=lib)
    cmp = A.compare(st[:L.NSTATE_CORE], g["ref32"], g["ref64"])
    cmp["flagged"] = flagged
    cmp["seconds"] = el
    return cmp


def main():
    argv = sys.argv[1:]
    cases, variants, out_json = list(M.CASES), ["strict", "parity", "fast"], None
    extras = {}
    it = iter(argv)
    for a in it:
        if a == "--cases":
            cases = next(it).split(",")
        elif a == "--variants":
            variants = [v for v in next(it).split(",") if v]
        elif a == "--json":
            out_json = next(it)
        elif "=" in a:
            k, v = a.split("=", 1)
            extras[k] = v
    t = ParamTables.default()
    res = {}
    for name in cases:
        for v in variants + list(extras):
            cmp = run_case(name, v, t, lib=extras.get(v))
            res[f"{name}:{v}"] = {"summary": A.summary(cmp), "planes": cmp["planes"], "flagged": cmp["flagged"]}
            print(f"==== {name} / {v}: pass={cmp['pass']} flagged={cmp['flagged']} ({cmp['seconds']:.2f} s on the GPU)")
            print(A.table(cmp))
            print("worst vs bar:", cmp["worst"], flush=True)
    if out_json:
        with open(out_json, "w") as fh:
            json.dump(res, fh)


if __name__ == "__main__":
    main()
