"""Host-side driver logic around the path: frequency parsing, calendar stepping, solar scalars, initial
snow layering, restart files, and the numpy statement of the forcing-preparation kernel."""
import datetime as dt
import math

import numpy as np
import pytest

from noahmp_b200 import forcing as F
from noahmp_b200 import layout as L
from noahmp_b200.driver import read_restart, snow_layers_from_depth, write_restart
from oracle.forcing_prep import prepare_forcing


def test_frequency_parsing_and_stepping():
    assert F.parse_frequency("1 hour") == (3600, "seconds") and F.parse_frequency("3 hours") == (10800, "seconds")
    assert F.parse_frequency("30 minute") == (1800, "seconds") and F.parse_frequency(900) == (900, "seconds")
    assert F.parse_frequency("1 month") == (1, "months") and F.parse_frequency("1 year") == (12, "months")
    with pytest.raises(ValueError):
        F.parse_frequency("every now and then")
    t = dt.datetime(2000, 1, 31)
    assert F.advance(t, (1, "months")) == dt.datetime(2000, 2, 29)          # leap year, clipped day
    assert F.advance(t, (3600, "seconds")) == dt.datetime(2000, 1, 31, 1)
    o = dt.datetime(2000, 1, 1)
    assert F.is_boundary(dt.datetime(2000, 1, 1, 3), o, (10800, "seconds"))
    assert not F.is_boundary(dt.datetime(2000, 1, 1, 2), o, (10800, "seconds"))
    assert F.is_boundary(dt.datetime(2000, 3, 1), o, (1, "months")) and not F.is_boundary(dt.datetime(2000, 3, 2), o, (1, "months"))


def test_julian_and_yearlen():
    assert F.julian_of(dt.datetime(2000, 1, 1)) == 0.0 and F.yearlen_of(dt.datetime(2000, 6, 1)) == 366
    assert F.julian_of(dt.datetime(2001, 12, 31, 12)) == 364.5 and F.yearlen_of(dt.datetime(2001, 1, 1)) == 365


def test_cosz_known_geometry():
    """Equinox-ish noon at the equator: sun near the zenith; midnight: below the horizon; polar night at 80N in December."""
    geo = F.geo_planes(np.radians([0.0, 0.0, 80.0, 40.0]), np.radians([0.0, 180.0, 0.0, -105.0]))
    rec = np.zeros((8, 4), np.float32)
    noon = prepare_forcing(rec, rec, 0.0, geo, F.solar_scalars(dt.datetime(2001, 3, 22, 12)))[8]
    assert noon[0] > 0.999 and noon[1] < -0.999
    dec = prepare_forcing(rec, rec, 0.0, geo, F.solar_scalars(dt.datetime(2001, 12, 21, 12)))[8]
    assert dec[2] < 0.0
    # 40N 105W, local solar noon = 19 UTC, June solstice: cos z = cos(40 - 23.44 deg)
    jun = prepare_forcing(rec, rec, 0.0, geo, F.solar_scalars(dt.datetime(2001, 6, 21, 19)))[8]
    assert abs(jun[3] - math.cos(math.radians(40.0 - 23.44))) < 2e-3


def test_prepare_forcing_interpolates_linearly():
    rng = np.random.default_rng(0)
    a, b = rng.normal(size=(8, 100)).astype(np.float32), rng.normal(size=(8, 100)).astype(np.float32)
    geo = F.geo_planes(rng.uniform(-1, 1, 100), rng.uniform(-3, 3, 100))
    s = F.solar_scalars(dt.datetime(2000, 7, 1, 15, 30))
    assert np.array_equal(prepare_forcing(a, b, 0.0, geo, s)[:8], a)
    assert np.allclose(prepare_forcing(a, b, 1.0, geo, s)[:8], b, atol=1e-6)
    mid = prepare_forcing(a, b, 0.25, geo, s)
    assert np.allclose(mid[:8], 0.75 * a + 0.25 * b, atol=1e-6) and (np.abs(mid[8]) <= 1.0 + 1e-6).all()


def test_initial_snow_layering_invariants():
    snowh = np.array([0.0, 0.01, 0.03, 0.08, 0.2, 0.4, 1.5], np.float32)
    sneqv = 250.0 * snowh
    isnow, dz, snice, zs = snow_layers_from_depth(snowh, sneqv, None)
    assert isnow.tolist() == [0, 0, -1, -2, -2, -3, -3]
    thick = dz.sum(axis=0)
    assert np.allclose(thick[isnow < 0], snowh[isnow < 0], rtol=1e-6)
    assert np.allclose(snice.sum(axis=0)[isnow < 0], sneqv[isnow < 0], rtol=1e-5)
    for c in range(snowh.size):
        act = zs[isnow[c] + 3:, c]
        assert (np.diff(act) < 0).all() and act[0] < 0
        assert np.isclose(act[-1], -thick[c] - 2.0, atol=1e-5)
        assert (dz[:isnow[c] + 3, c] == 0).all() and (dz[isnow[c] + 3:, c] >= 0.024).all()


def test_restart_round_trip(tmp_path):
    st = np.random.default_rng(1).normal(size=(L.NSTATE, 7)).astype(np.float32)
    st[L.S["ISNOW"]] = np.array([0, -1, -2, -3, 0, 0, -1], np.int32).view(np.float32)
    p = str(tmp_path / "r" / "RESTART.200001010300.npz")
    write_restart(p, st, dt.datetime(2000, 1, 1, 3), 12)
    st2, when, step = read_restart(p)
    assert np.array_equal(st.view(np.int32), st2.view(np.int32)) and when == dt.datetime(2000, 1, 1, 3) and step == 12
