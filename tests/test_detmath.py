"""Shared deterministic math (include/romis_detmath.h) and the counter-based RNG (include/romis_rng.h)."""
import math

import numpy as np

from oracle.pyoracle import Oracle


def ulp_diff(a: np.float32, b: np.float32) -> int:
    if np.isnan(a) and np.isnan(b):
        return 0
    if a == b:
        return 0
    ia, ib = int(np.float32(a).view(np.int32)), int(np.float32(b).view(np.int32))
    if (ia < 0) != (ib < 0):
        return 1 << 30
    return abs(ia - ib)


def test_powf_expf_within_one_ulp_of_libm():
    orc = Oracle()
    rng = np.random.default_rng(0)
    worst = 0
    for _ in range(20000):
        x = np.float32(rng.uniform(0, 2)); y = np.float32(rng.uniform(0, 300))
        if rng.random() < 0.3:
            y = np.float32(rng.integers(0, 300))          # integer exponents take the squaring path
        ref = np.float32(math.pow(float(x), float(y))) if x > 0 else np.float32(0.0 if y > 0 else 1.0)
        worst = max(worst, ulp_diff(np.float32(orc.lib.orc_powf(float(x), float(y))), ref))
    assert worst <= 1
    worst = 0
    for _ in range(20000):
        x = np.float32(rng.uniform(-80, 80))
        worst = max(worst, ulp_diff(np.float32(orc.lib.orc_expf(float(x))), np.float32(math.exp(float(x)))))
    assert worst <= 1
    orc.close()


def test_powf_special_cases_follow_c99():
    orc = Oracle()
    p = lambda x, y: np.float32(orc.lib.orc_powf(x, y))
    inf, nan = float("inf"), float("nan")
    assert p(nan, 0.0) == 1 and p(1.0, nan) == 1 and np.isnan(p(nan, 2.0)) and np.isnan(p(2.0, nan))
    assert np.isnan(p(-0.5, 0.5))                         # negative base, non-integer exponent
    assert p(-0.5, 3.0) == np.float32(-0.125) and p(-0.5, 2.0) == np.float32(0.25)
    assert p(0.0, -1.0) == inf and p(-0.0, -1.0) == -inf and p(0.0, 2.0) == 0
    assert p(2.0, inf) == inf and p(0.5, inf) == 0 and p(-1.0, inf) == 1
    assert p(inf, -1.0) == 0 and p(-inf, 3.0) == -inf
    assert p(10.0, 39.0) == inf and p(0.1, 50.0) == 0 and p(0.99, 250.0) == np.float32(math.pow(np.float32(0.99), 250.0))
    assert np.float32(orc.lib.orc_expf(-inf)) == 0 and np.float32(orc.lib.orc_expf(inf)) == inf and np.float32(orc.lib.orc_expf(0.0)) == 1
    orc.close()


def test_rng_streams_are_uniform_and_keyed():
    orc = Oracle()
    bits = np.array([orc.lib.orc_rng_bits(99, 3, 2, 1234, 1, c) for c in range(20000)], np.uint64)
    u = (bits >> np.uint64(1)).astype(np.float64) / 2147483648.0
    assert 0.49 < u.mean() < 0.51 and 0.32 < (u < 1 / 3).mean() < 0.345
    assert len(np.unique(bits)) > 19990
    other = np.array([orc.lib.orc_rng_bits(99, 3, 2, 1235, 1, c) for c in range(2000)], np.uint64)
    assert (other == bits[:2000]).sum() == 0                # neighbouring pixels: unrelated streams
    again = np.array([orc.lib.orc_rng_bits(99, 3, 2, 1234, 1, c) for c in range(100)], np.uint64)
    assert np.array_equal(again, bits[:100])                # addressable without state
    # multiply-shift mapping of uniform_int_distribution(-r, r) covers the window evenly
    r = 10
    d = -r + ((bits.astype(np.uint64) * np.uint64(2 * r + 1)) >> np.uint64(32)).astype(np.int64)
    assert d.min() == -r and d.max() == r and np.bincount(d + r).min() > 20000 / 21 * 0.8
    orc.close()


def test_specular_cutoff_claim():
    """romis_specular_cutoff(s) = c promises pow(x, s) in {+-0, NaN} for every |x| <= c (include/romis_gpu.h): checked
    against the oracle's romis_powf on dense samples of [-c, c], the boundary itself, and just outside for tightness.
    (Host-only function of libromis_gpu.so: no GPU call.)"""
    import ctypes
    from romis_b200.api import load_library
    lib = load_library()
    orc = Oracle().lib
    orc.orc_powf.restype = ctypes.c_float; orc.orc_powf.argtypes = [ctypes.c_float, ctypes.c_float]
    rng = np.random.default_rng(3)
    for s in (1.0, 3.5, 10.000004768371582, 32.0, 250.0, 1000.0, 65536.0, 70000.5):
        c = np.float32(lib.romis_specular_cutoff(ctypes.c_float(s)))
        if s < 60:                     # 0.05^s must underflow: s log2(20) > 150
            assert c == 0.0, (s, c)
            continue
        assert 0.05 <= c < 1.0
        xs = np.concatenate([rng.uniform(-c, c, 20000), c * (1 - np.logspace(-7, -1, 2000)), [c, -c, 0.0]]).astype(np.float32)
        xs = xs[np.abs(xs) <= c]
        for x in xs:
            r = np.float32(orc.orc_powf(ctypes.c_float(float(x)), ctypes.c_float(s)))
            assert r == 0.0 or np.isnan(r), (s, float(x), float(r))
        # tight to within 0.2 %: just above the bound the lobe is alive
        assert np.float32(orc.orc_powf(ctypes.c_float(float(c) * 1.002), ctypes.c_float(s))) != 0.0
    for s in (0.0, -2.0, 0.5, float("nan"), float("inf")):
        assert lib.romis_specular_cutoff(ctypes.c_float(s)) == 0.0


def test_pow_with_exponent_one_is_the_identity():
    """The tone-mapping shortcut of the kernels (tone_map, csrc/device_common.cuh) skips pow(x, 1 / gamma) for gamma = 1:
    romis_powf(x, 1) must return x bit for bit, for every kind of x."""
    orc = Oracle()
    rng = np.random.default_rng(11)
    xs = np.concatenate([
        rng.uniform(0, 1, 20000).astype(np.float32),                       # the range 1 - exp(-exposure * c) lives in
        rng.uniform(-4, 4, 5000).astype(np.float32),
        rng.integers(0, 1 << 32, 20000, dtype=np.uint64).astype(np.uint32).view(np.float32),    # any bit pattern
        np.array([0.0, -0.0, 1.0, -1.0, np.inf, -np.inf, 1e-45, -1e-45, 1.17549435e-38, 3.4028235e38, 5.9e-39], np.float32)])
    for x in xs:
        r = np.float32(orc.lib.orc_powf(float(x), 1.0))
        if np.isnan(x):
            assert np.isnan(r)
        else:
            assert r.view(np.uint32) == x.view(np.uint32), f"pow({x!r}, 1) = {r!r}"
    orc.close()


def test_uniform_float_conversion_is_one_multiplication():
    """Device form of romis_rand_to_unit: float(r) * 2^-31 has the bits of ((float(r) - 0) / (2^31 - 0)) * (1 - 0) + 0 for r >= 0."""
    rng = np.random.default_rng(12)
    r = np.concatenate([rng.integers(0, 1 << 31, 200000, dtype=np.int64), np.array([0, 1, 2, (1 << 31) - 1, (1 << 31) - 64, (1 << 24) + 1])])
    f = r.astype(np.float32)                                               # int -> float, round to nearest even, as the cast
    host = (((f - np.float32(0)) / (np.float32(2147483648.0) - np.float32(0))) * (np.float32(1) - np.float32(0))) + np.float32(0)
    dev = f * np.float32(4.656612873077392578125e-10)
    assert np.array_equal(host.view(np.uint32), dev.view(np.uint32))
    assert host.min() >= 0 and host.max() <= 1 and not np.signbit(host).any()
