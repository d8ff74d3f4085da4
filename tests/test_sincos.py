"""shapes_sincos (include/shapes_sincos.h): the cos/sin shared by the host and the device-resident
world.  Host evaluation only (the library's exported function needs no GPU); the device copy is
compared bit for bit in tests/test_gpu_world.py."""
import numpy as np

from shapes_b200 import engine


def _ulps(a, b):
    ia = a.view(np.int64).copy(); ib = b.view(np.int64).copy()
    ia[ia < 0] = np.int64(-2**63) - ia[ia < 0]
    ib[ib < 0] = np.int64(-2**63) - ib[ib < 0]
    return np.abs(ia - ib)


def test_sincos_within_one_ulp_of_libm(oracle, product_lib):
    rng = np.random.default_rng(5)
    for scale in (0.78, 3.2, 100.0, 1e5, 1e9, 1e12):
        x = rng.uniform(-scale, scale, 200_000)
        c, s = engine.sincos(x)
        lc, ls = oracle.cos_sin(x)
        uc, us = _ulps(c, lc), _ulps(s, ls)
        assert uc.max() <= 1 and us.max() <= 1, (scale, uc.max(), us.max())
        assert (uc != 0).mean() < 0.06 and (us != 0).mean() < 0.06
    # neighbours of multiples of pi/2, where the reduction cancels
    k = np.arange(-20000, 20001, dtype=np.float64) * (np.pi / 2)
    x = np.concatenate([np.nextafter(k, np.inf), np.nextafter(k, -np.inf), k])
    c, s = engine.sincos(x)
    lc, ls = oracle.cos_sin(x)
    assert _ulps(c, lc).max() <= 1 and _ulps(s, ls).max() <= 1


def test_sincos_special_values(product_lib):
    c, s = engine.sincos(np.array([0.0, -0.0, np.inf, -np.inf, np.nan, 2.0**61, np.nextafter(1e12, np.inf), -1.0e13]))
    assert c[0] == 1.0 and s[0] == 0.0 and not np.signbit(s[0])
    assert c[1] == 1.0 and s[1] == 0.0 and np.signbit(s[1])
    assert np.isnan(c[2:]).all() and np.isnan(s[2:]).all()       # beyond the proven range (|x| > 1e12): NaN, never a drift
    c, s = engine.sincos(np.array([1e12, -1e12]))
    assert np.isfinite(c).all() and np.isfinite(s).all()
    # unit circle to rounding, symmetric in sign
    x = np.linspace(-50, 50, 10001)
    c, s = engine.sincos(x)
    assert np.abs(c * c + s * s - 1.0).max() < 4e-16
    c2, s2 = engine.sincos(-x)
    assert np.array_equal(c, c2) and np.array_equal(s, -s2)
