"""Synthetic workloads: reproducible, inside the validity envelope of SURVEY App. G, rank-count independent."""
import numpy as np
import pytest

from noahmp_b200 import layout as L, synth


@pytest.mark.parametrize("name", ["single", "uniform100k", "global025", "conus1km", "snow1m"])
def test_initial_state_envelope(tables, name):
    w = synth.make_workload(name, ncol=500)
    st = w.state
    sl = (w.static_i >> 8) & 0xff
    lu = w.static_i & 0xff
    assert ((sl >= 1) & (sl <= 19) & (sl != 14)).all() and not np.isin(lu, [16, 24]).any()
    smcmax = np.asarray(tables.soil["SMCMAX"], np.float32)[sl - 1]
    for k in range(1, 5):
        smc, liq = st[L.smc(k)], st[L.soilwat(k)]
        assert (smc > 0).all() and (smc <= smcmax + 1e-6).all() and (liq >= 0).all() and (liq <= smc + 1e-7).all()
    isn = st[L.S["ISNOW"]].view(np.int32)
    z = np.stack([st[L.zsnso(j)] for j in range(-2, 5)])
    for c in range(0, w.ncol, 50):
        act = z[isn[c] + 3:, c]
        assert (np.diff(act) < 0).all() and act[0] < 0            # strictly decreasing layer bottoms
    if name == "snow1m":
        tot = sum(st[L.snice(j)] + st[L.snliq(j)] for j in range(-2, 1))
        assert np.allclose(tot, st[L.S["SNEQV"]], rtol=1e-6) and (isn == -3).all()
        assert np.allclose(-st[L.zsnso(0)], st[L.S["SNOWH"]], rtol=1e-6)
    else:
        assert (isn == 0).all()


def test_forcing_is_deterministic_and_shard_independent():
    w = synth.make_workload("conus1km", ncol=1000, nsteps=3)
    full = synth.forcing_step(w, 2)
    again = synth.forcing_step(w, 2)
    assert np.array_equal(full, again)
    import copy
    part = copy.copy(w)
    part.sincos = np.ascontiguousarray(w.sincos[:, 400:700])
    assert np.array_equal(synth.forcing_step(part, 2, col0=400), full[:, 400:700])
    assert not np.array_equal(synth.forcing_step(w, 1), full)


def test_forcing_ranges():
    w = synth.make_workload("global025", ncol=4000, nsteps=48)
    f = synth.forcing_all(w)
    F = L.F
    assert (f[:, F["SFCPRS"]] > 8e4).all() and (f[:, F["Q2"]] >= 1e-4).all() and (f[:, F["LWDN"]] > 50).all()
    assert (f[:, F["SOLDN"]] >= 0).all() and (np.abs(f[:, F["COSZ"]]) <= 1).all()
    assert ((f[:, F["SOLDN"]] > 0) == (f[:, F["COSZ"]] > 0)).all()
    wet = f[:, F["PRCP"]] > 0
    assert 0.03 < wet.mean() < 0.09
    assert 200 < f[:, F["SFCTMP"]].min() and f[:, F["SFCTMP"]].max() < 330
