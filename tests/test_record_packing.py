"""Zero suppression of record planes (noahmp_set_record_packing / noahmp_pack_plane): the host packer against a numpy
statement of the format (no device needed), and - on the GPU - packed records end in the bit-identical state."""
import numpy as np
import pytest

from noahmp_b200.engine import pack_plane


def numpy_pack(plane: np.ndarray):
    """The format as include/noahmp_b200.h states it: one occupancy word per 32 columns (bit b = word 32 g + b is not the
    bit pattern 0), the count of non-zero words in earlier groups, and the non-zero words in order."""
    x = plane.view(np.uint32)
    n = x.shape[0]
    g = (n + 31) // 32
    nz = np.zeros(g * 32, bool)
    nz[:n] = x != 0
    bits = nz.reshape(g, 32)
    occ = (bits.astype(np.uint64) << np.arange(32, dtype=np.uint64)).sum(1).astype(np.uint32)
    cnt = bits.sum(1)
    before = (np.cumsum(cnt) - cnt).astype(np.uint32)
    return occ, before, plane[x != 0]


def numpy_unpack(occ, before, values, n):
    """What the step kernel's load does (csrc/column_step.cuh: ld_record), column by column."""
    i = np.arange(n)
    g, lane = i >> 5, (i & 31).astype(np.uint32)
    m = occ[g]
    present = ((m >> lane) & 1).astype(bool)
    below = m & ((np.uint32(1) << lane) - np.uint32(1))
    rank = np.array([bin(int(v)).count("1") for v in below], np.int64) if n < 5000 else _popcount(below)
    out = np.zeros(n, np.float32)
    out[present] = values[before[g][present].astype(np.int64) + rank[present]]
    return out


def _popcount(v):
    v = v.astype(np.uint64)
    v = v - ((v >> 1) & 0x5555555555555555)
    v = (v & 0x3333333333333333) + ((v >> 2) & 0x3333333333333333)
    v = (v + (v >> 4)) & 0x0F0F0F0F0F0F0F0F
    return ((v * 0x0101010101010101) >> 56).astype(np.int64) & 0xFF


@pytest.mark.parametrize("threads", [1, 3])
@pytest.mark.parametrize("n", [1, 31, 32, 33, 777, 65_536, 300_007])
@pytest.mark.parametrize("fill", [0.0, 0.05, 0.5, 1.0])
def test_packer_matches_the_stated_format_and_unpacks_losslessly(n, fill, threads):
    rng = np.random.default_rng(n + int(100 * fill))
    plane = rng.standard_normal(n).astype(np.float32)
    plane[rng.random(n) >= fill] = 0.0
    if n > 40:
        plane[5] = -0.0                      # -0 is not the bit pattern 0: it travels
        plane[37] = np.float32(1e-42)        # so does a denormal
        plane[n - 1] = np.nan if fill > 0 else 0.0
    occ, before, values = pack_plane(plane, threads)
    rocc, rbefore, rvalues = numpy_pack(plane)
    assert np.array_equal(occ, rocc) and np.array_equal(before, rbefore)
    assert np.array_equal(values.view(np.uint32), rvalues.view(np.uint32))
    back = numpy_unpack(occ, before, values, n)
    assert np.array_equal(back.view(np.uint32), plane.view(np.uint32))
    if n % 32:
        assert occ[-1] >> (n % 32) == 0      # no bit beyond the last column


def test_contiguous_day_night_plane():
    """Short-wave radiation over a grid that the terminator crosses: whole groups full (copied as a block) or empty."""
    n = 200_000
    plane = np.zeros(n, np.float32)
    plane[70_001:150_003] = np.linspace(1.0, 900.0, 80_002, dtype=np.float32)
    occ, before, values = pack_plane(plane, 4)
    assert values.shape[0] == 80_002 and np.array_equal(values, plane[70_001:150_003])
    assert np.array_equal(numpy_unpack(occ, before, values, n), plane)


def test_packer_property_random_shapes():
    """Any size, any fill pattern (runs of zeros and non-zeros of random length, as a terminator or a rain band makes them), any
    thread count: the packed form is the stated one and unpacks to the plane bit for bit."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=60, deadline=None)
    @given(n=st.integers(1, 20_000), seed=st.integers(0, 2**31 - 1), threads=st.integers(1, 6), mean_run=st.integers(1, 400))
    def check(n, seed, threads, mean_run):
        rng = np.random.default_rng(seed)
        plane = rng.standard_normal(n).astype(np.float32)
        pos, zero = 0, bool(rng.integers(2))
        while pos < n:                                   # alternating runs
            run = int(rng.geometric(1.0 / mean_run))
            if zero:
                plane[pos:pos + run] = 0.0
            pos, zero = pos + run, not zero
        occ, before, values = pack_plane(plane, threads)
        rocc, rbefore, rvalues = numpy_pack(plane)
        assert np.array_equal(occ, rocc) and np.array_equal(before, rbefore)
        assert np.array_equal(values.view(np.uint32), rvalues.view(np.uint32))
        assert np.array_equal(numpy_unpack(occ, before, values, n).view(np.uint32), plane.view(np.uint32))

    check()


def test_packer_many_threads_many_calls():
    """The thread pool of the packer: more threads than chunks, repeated calls, a large plane split into many chunks."""
    rng = np.random.default_rng(3)
    plane = rng.standard_normal(3_000_001).astype(np.float32)
    plane[rng.random(plane.shape[0]) < 0.9] = 0.0
    ref = numpy_pack(plane)
    for threads in (2, 7, 16, 64):
        for _ in range(3):
            occ, before, values = pack_plane(plane, threads)
            assert np.array_equal(occ, ref[0]) and np.array_equal(before, ref[1]) and np.array_equal(values, ref[2])
