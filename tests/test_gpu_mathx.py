"""csrc/mathx.cuh: the forward kernels' log / atan2 (CUDA's own algorithms with the polynomial coefficients in the
constant bank instead of 64-bit immediates) must return, bit for bit, what the CUDA library returns -- on the main path
because the operations are the same, off it because the library routine itself is called."""
import numpy as np
import pytest

import tomofastx_b200 as tfx

pytestmark = pytest.mark.gpu


def _same_bits(a, b):
    nan = np.isnan(a) & np.isnan(b)
    return np.all((a.view(np.int64) == b.view(np.int64)) | nan)


def _inputs(rng, n):
    mag = 10.0 ** rng.uniform(-12, 12, n)
    x = mag * rng.choice([-1.0, 1.0], n)
    y = x * 10.0 ** rng.uniform(-9, 9, n) * rng.choice([-1.0, 1.0], n)
    return y, x


def test_log_and_atan2_match_the_cuda_library_bit_for_bit():
    rng = np.random.default_rng(12345)
    y, x = _inputs(rng, 2_000_000)
    # the forward kernels' own argument ranges: distances of 1e0..1e5 m, products of two of them
    y2 = rng.uniform(-1e5, 1e5, 1_000_000) * rng.uniform(-1e5, 1e5, 1_000_000)
    x2 = rng.uniform(-5e3, 5e3, 1_000_000) * rng.uniform(1.0, 2e5, 1_000_000)
    # ratios near 1 and near 0 (both reflections of the atan argument reduction), mantissas near sqrt(2) for log
    x3 = np.ldexp(rng.uniform(1.41421, 1.41422, 200_000), rng.integers(-300, 300, 200_000))
    y3 = x3 * (1.0 + rng.uniform(-1e-12, 1e-12, 200_000))
    special = np.array([0.0, -0.0, 1.0, -1.0, np.inf, -np.inf, np.nan, 5e-324, 2.2250738585072014e-308, 1e-310,
                        1.7976931348623157e308, 1e-290, 1e290, 2.0 ** -921, 2.0 ** -922, 2.0 ** 1021, 2.0 ** 1022, 0.5, 2.0])
    ys, xs = (a.ravel() for a in np.meshgrid(special, special))
    y = np.concatenate([y, y2, y3, ys, np.zeros(1000), rng.uniform(-1e4, 1e4, 1000)])
    x = np.concatenate([x, x2, x3, xs, rng.uniform(-1e4, 1e4, 1000), np.zeros(1000)])
    mine_log, lib_log, mine_at, lib_at = tfx.debug_math(y, x)
    assert _same_bits(mine_log, lib_log), int(np.sum(mine_log.view(np.int64) != lib_log.view(np.int64)))
    assert _same_bits(mine_at, lib_at), int(np.sum(mine_at.view(np.int64) != lib_at.view(np.int64)))
    # and both are what numpy (glibc) computes, to rounding
    ok = np.isfinite(x) & (x > 0)
    assert np.allclose(mine_log[ok], np.log(x[ok]), rtol=4e-16, atol=5e-324)
    fin = np.isfinite(x) & np.isfinite(y)
    assert np.allclose(mine_at[fin], np.arctan2(y[fin], x[fin]), rtol=1e-15, atol=1e-300)
