"""Memory-safety evidence without compute-sanitizer (closed on the GPU pool): the `bounds` build is the strict build with
every data-dependent index of the step kernel range-checked on the device (csrc/common.cuh NMP_CHK: pick(), the snow-pack
scratch of combine / divide func:5236-5533, phenology's month indices func:580-597, the table rows of the type indices,
the root-layer count, the water-table layer func:6545-6551).  A violation is counted and clamped, never made.

Run with `-m gpu`.  The strict suite's cases and BASELINE configs[4] (the snow ensemble: combine / divide every few steps)
run under it: zero violations, and the state equals the strict build's bit for bit (the checks do not touch the arithmetic).
"""
import numpy as np
import pytest
import torch

from noahmp_b200 import layout as L, synth
from noahmp_b200.engine import NoahMP

import make_golden
from test_parity_gpu import STRICT_CASES, gpu_run

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_the_checker_counts_and_clamps(tables):
    eng = NoahMP(tables, device=0, fast="bounds")
    assert eng.debug_bounds(reset=True)["checked"]
    assert eng.debug_bounds_selftest(2)                       # in range: nothing counted
    assert eng.debug_bounds()["hits"] == 0
    assert eng.debug_bounds_selftest(7) and eng.debug_bounds_selftest(-1)
    got = eng.debug_bounds(reset=True)
    assert got["hits"] == 2 and got["sites"] == 1             # two violations, both at site 0 (pick)
    assert eng.debug_bounds()["hits"] == 0
    eng.close()
    other = NoahMP(tables, device=0, fast="strict")           # every other build: not instrumented, says so
    assert other.debug_bounds() == {"checked": False, "hits": 0, "sites": 0} and not other.debug_bounds_selftest(7)
    other.close()


def _run_checked(tables, w, forcing, diag=True):
    eng = NoahMP(tables, w.opts, device=0, fast="bounds")
    eng.debug_bounds(reset=True)
    st = torch.from_numpy(w.state.copy()).to(DEV)
    dg = torch.zeros((L.NDIAG, w.ncol), dtype=torch.float32, device=DEV) if diag else None
    bd = torch.zeros(L.NBUDGET, dtype=torch.float64, device=DEV)
    rp = eng.run_params(w.dt, synth.ZSOIL, w.julian0, w.yearlen, w.zref, w.co2_frac, w.o2_frac, w.foln)
    eng.run(rp, st, torch.from_numpy(forcing).to(DEV), torch.from_numpy(w.static_f).to(DEV), torch.from_numpy(w.static_i).to(DEV), dg, bd)
    torch.cuda.synchronize()
    hits = eng.debug_bounds(reset=True)
    eng.close()
    return st.cpu().numpy(), hits, int(bd[L.B["NFLAGGED"]].item())


@pytest.mark.parametrize("workload,ncol,nsteps,over", STRICT_CASES, ids=lambda v: str(v).replace(" ", ""))
def test_strict_suite_under_the_bounds_build(tables, workload, ncol, nsteps, over):
    opts = make_golden.options(over)
    w = synth.make_workload(workload, ncol=ncol, nsteps=nsteps, opts=opts)
    forcing = synth.forcing_all(w)
    st, hits, nflag = _run_checked(tables, w, forcing)
    assert hits["checked"] and hits["hits"] == 0, hits
    ref, _, bd = gpu_run(tables, w, forcing, "strict", diag=True)
    assert np.array_equal(st.view(np.int32), ref.view(np.int32))
    assert nflag == int(bd[L.B["NFLAGGED"]])


def test_snow_ensemble_full_length_under_the_bounds_build(tables):
    """BASELINE configs[4] at its full 720 steps on a 20 000-column sample: three-layer packs melting, refreezing, merging
    and splitting - the code with run-time loop bounds and merge targets (func:5276-5411, 5443-5526)."""
    w = synth.make_workload("snow1m", ncol=20000, nsteps=24)
    eng = NoahMP(tables, w.opts, device=0, fast="bounds")
    eng.debug_bounds(reset=True)
    st = torch.from_numpy(w.state.copy()).to(DEV)
    sf, si = torch.from_numpy(w.static_f).to(DEV), torch.from_numpy(w.static_i).to(DEV)
    bd = torch.zeros(L.NBUDGET, dtype=torch.float64, device=DEV)
    seen = set()
    for chunk in range(30):
        f = np.stack([synth.forcing_step(w, chunk * 24 + s) for s in range(24)])
        rp = eng.run_params(w.dt, synth.ZSOIL, w.julian0 + chunk * 24 * w.dt / 86400.0, w.yearlen, w.zref, w.co2_frac, w.o2_frac, w.foln)
        eng.run(rp, st, torch.from_numpy(f).to(DEV), sf, si, None, bd)
        seen |= set(np.unique(st[L.S["ISNOW"]].cpu().numpy().view(np.int32)).tolist())
    torch.cuda.synchronize()
    hits = eng.debug_bounds()
    eng.close()
    assert hits["checked"] and hits["hits"] == 0, hits
    assert int(bd[L.B["NFLAGGED"]].item()) == 0
    assert len(seen) >= 3, seen             # packs merged and split on the way (layer counts changed)


def test_type_indices_outside_the_tables_never_reach_a_table_row(tables):
    """Columns whose type indices are outside the tables are flagged up front and run on substitute type 1
    (column_step.cuh); the row checks (site 4) must stay silent for them."""
    w = synth.make_workload("global025", ncol=512, nsteps=4)
    bad = w.static_i.copy()
    bad[::7] = L.pack_types(99, 77, 40, 15, 1, 0)
    w.static_i = bad.astype(np.int32)
    _, hits, nflag = _run_checked(tables, w, synth.forcing_all(w))
    assert hits["hits"] == 0 and nflag == len(bad[::7])
