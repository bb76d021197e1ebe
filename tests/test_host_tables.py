"""The height-score bin thresholds (host libm, no device): the table the binning kernel searches must reproduce
pingpongpro.cpp:591-594 for EVERY float product, i.e. #thresholds <= hs == (int)(0.5 + log10f(hs) / maxHS * 999).
The literal expression is restated here on top of the C library's log10f (ctypes) and numpy float32 arithmetic."""
import ctypes
import ctypes.util

import numpy as np
import pytest

from pingpongpro_b200 import api

_libm = ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6")
_libm.log10f.restype = ctypes.c_float
_libm.log10f.argtypes = [ctypes.c_float]

F = np.float32


def log10f(x) -> np.float32:
    return F(_libm.log10f(ctypes.c_float(float(x))))


def max_height_score(max_freq: float) -> np.float32:
    return log10f(F(max_freq) * F(max_freq))  # :551


def ref_bin(hs, max_hs: np.float32) -> int:
    t = F(F(log10f(hs) / max_hs) * F(999.0))  # :593, float throughout
    return int(0.5 + float(t))                # :594, the 0.5 is a double


def below(x: np.float32) -> np.float32:
    return np.nextafter(F(x), F(0))


MAX_FREQS = [2, 3, 5, 17, 100, 622, 4096, 12345, 250000, 1e6, 9999999, 16777215, 16777216]


@pytest.mark.parametrize("max_freq", MAX_FREQS)
def test_every_threshold_is_the_first_float_of_its_bin(max_freq):
    thr = api.bin_thresholds(max_freq)
    max_hs = max_height_score(max_freq)
    top = F(max_freq) * F(max_freq)
    assert ref_bin(top, max_hs) == 999  # the tallest product sits in the last bin
    assert np.all(np.isfinite(thr)) and np.all(np.diff(thr) >= 0) and thr[0] > 1 and thr[-1] <= top
    for b in range(1, 1000):
        t = thr[b - 1]
        assert ref_bin(t, max_hs) >= b, (max_freq, b)
        if b == 1 or thr[b - 2] < t:
            assert ref_bin(below(t), max_hs) < b, (max_freq, b)


@pytest.mark.parametrize("max_freq", [5, 622, 12345, 1e6, 16777216])
def test_threshold_count_equals_the_reference_bin_for_products(max_freq):
    """The device computes the bin as the number of thresholds <= hs: random products, the floats around every
    threshold, and the extremes."""
    thr = api.bin_thresholds(max_freq)
    max_hs = max_height_score(max_freq)
    rng = np.random.default_rng(int(max_freq))
    a = np.floor(10 ** rng.uniform(0, np.log10(max_freq), 3000)).astype(np.float32)
    b = np.floor(10 ** rng.uniform(0, np.log10(max_freq), 3000)).astype(np.float32)
    hs = np.concatenate([a * b, [F(1), F(max_freq), F(max_freq) * F(max_freq)],
                         thr[::7], np.nextafter(thr[::7], F(0)), np.nextafter(thr[::7], F(np.inf))]).astype(np.float32)
    hs = hs[(hs >= 1) & (hs <= F(max_freq) * F(max_freq))]
    got = np.searchsorted(thr, hs, side="right")
    want = np.array([ref_bin(x, max_hs) for x in hs])
    assert np.array_equal(got, want)


def test_many_maximum_frequencies_validate():
    """ppp_host_bin_thresholds re-checks the floats around every step with the literal expression and fails otherwise."""
    rng = np.random.default_rng(11)
    for mf in np.floor(10 ** rng.uniform(0.31, 7.2247, 300)):
        thr = api.bin_thresholds(float(mf))
        assert thr.shape == (999,) and thr[-1] <= F(mf) * F(mf)


def test_rejects_degenerate_maximum():
    for bad in (1.0, 0.0, -3.0, float("nan")):
        with pytest.raises(api.PppError):
            api.bin_thresholds(bad)


def test_device_restatement_of_log10f_equals_the_host_libm():
    """csrc/log10f_emul.h (evaluated here on the host; the device runs the same IEEE operations) against this host's
    log10f: every 97th float of [1, 2^48) plus whole binades.  The device's threshold tables are built on it and are
    compared with the host's table on every pass (api.cu: settle); this test is why that comparison never fails here."""
    import ctypes

    libm = ctypes.CDLL("libm.so.6")
    libm.log10f.restype = ctypes.c_float
    libm.log10f.argtypes = [ctypes.c_float]
    bits = np.concatenate([np.arange(0x3f800000, 0x3f800000 + (48 << 23), 9973, dtype=np.uint64),
                           np.arange(0x3f800000, 0x3f800000 + (1 << 16), dtype=np.uint64), np.arange(0x4b000000, 0x4b000000 + (1 << 16), dtype=np.uint64)]).astype(np.uint32)
    x = bits.view(np.float32)
    got = api.emulated_log10f(x)
    want = np.array([libm.log10f(float(v)) for v in x], dtype=np.float32)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_device_threshold_tables_equal_the_host_tables():
    rng = np.random.default_rng(11)
    freqs = [2.0, 3.0, 7.0, 1000.0, 12327.0, 10716714.0, 12540968.0, float(1 << 24)] + [float(v) for v in rng.integers(2, 1 << 24, size=400)]
    for mf in freqs:
        a, b = api.bin_thresholds(mf), api.emulated_bin_thresholds(mf)
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), mf


def _float_chain(h0, w):
    h = np.float32(h0)
    for v in w:
        h = np.float32(h + v)
    return h


def test_composed_addition_maps_equal_the_float_additions():
    """csrc/fadd_chain.h: a run of float additions as one integer map per binade (k_fold_huge folds stacks of 10^4..10^6
    reads this way) against the additions themselves (pingpongpro.cpp:487): read weights 1/NH, ties (powers of two),
    sums that start at zero, cross binades and saturate at 2^24."""
    rng = np.random.default_rng(5)
    real_total = n_total = 0
    for trial in range(300):
        n = int(rng.integers(1, 3000))
        kind = trial % 4
        nh = (rng.integers(1, 5, n) if kind == 0 else rng.integers(1, 65, n) if kind == 1 else rng.integers(1, 100000, n) if kind == 2
              else 1 << rng.integers(0, 12, n))
        w = (1.0 / nh.astype(np.float64)).astype(np.float32)
        w[rng.random(n) < 0.15] = 0.0  # the other strand's records are staged as +0.0f
        h0 = [0.0, 1.0, float(rng.integers(1, 1 << 24)), float(np.uint32(0x3f800000 + rng.integers(0, 24 << 23)).view(np.float32)), 16777215.0][trial % 5]
        got, real = api.emulated_fadd_chain(h0, w, int(rng.integers(1, 65)))
        assert np.float32(got).view(np.uint32) == _float_chain(h0, w).view(np.uint32), (trial, h0)
        real_total += real
        n_total += n
    assert real_total < n_total // 10  # nearly everything went through the maps
    # saturation: 2^24 + 1 = 2^24 (the tie rounds to the even mantissa), + 0.5 and + 1/3 do not move it either
    got, real = api.emulated_fadd_chain(16777216.0, np.array([1.0, 0.5, 1.0 / 3, 1.0] * 100, np.float32), 32)
    assert got == np.float32(16777216.0) and real == 0
    # an infinite height (NH:i:0 through the C ABI) stays infinite
    got, _ = api.emulated_fadd_chain(np.inf, np.ones(50, np.float32), 16)
    assert np.isinf(got)
