"""Host model of the fused Haar pass of csrc/wavelet.cu (haar_axes12_fused): along axis 2 the three lowest scales are
lifted inside groups of 8 consecutive positions (a group that starts at a multiple of 8 is closed under scales 1-3) and
the remaining scales are the Haar transform of every 8th element -- a line of length ceil(n / 8) -- with nscale - 3 scales,
where nscale is taken from the FULL length (wavelet_transform.F90:88-101). Bit-identical to the one-line transform for
every length, forward and inverse, and to the oracle."""
import numpy as np
import pytest

SQ2 = np.sqrt(2.0)


def nscale_of(n):
    return int(np.floor(np.log2(n))) if n >= 2 else 0


def haar_scales(x, scales, forward, length_for_pairs=None):
    """The reference's lifting steps for the given scales (1-based), in the reference's operation order."""
    n = x.size if length_for_pairs is None else length_for_pairs
    for s in (scales if forward else reversed(list(scales))):
        step, half = 1 << s, 1 << (s - 1)
        lo = np.arange(0, n, step)
        lo = lo[lo + half < n]                       # a pair exists when its high element is inside the line (:97-101)
        hi = lo + half
        if forward:                                  # :103-149
            h = x[hi] - x[lo]
            l = x[lo] + h * 0.5
            x[lo] = l * SQ2
            x[hi] = h / SQ2
        else:                                        # :186-232
            l = x[lo] / SQ2
            h = x[hi] * SQ2
            l = l - h * 0.5
            x[lo] = l
            x[hi] = h + l
    return x


@pytest.mark.parametrize("n", [2, 3, 7, 8, 9, 15, 16, 17, 31, 40, 67, 100, 128, 300, 513, 1000, 1024])
def test_low_scales_in_groups_of_8_plus_high_scales_on_every_8th(n):
    rng = np.random.default_rng(n)
    x = rng.standard_normal(n)
    ns = nscale_of(n)
    for forward in (True, False):
        want = haar_scales(x.copy(), range(1, ns + 1), forward)
        got = x.copy()
        low = range(1, min(3, ns) + 1)
        m = (n + 7) // 8
        if forward:
            haar_scales(got, low, True)                                   # fused kernel: groups of 8 columns
            if ns > 3:
                sub = got[::8].copy()
                assert sub.size == m
                got[::8] = haar_scales(sub, range(1, ns - 3 + 1), True)   # the pass over every 8th row
        else:
            if ns > 3:
                sub = got[::8].copy()
                got[::8] = haar_scales(sub, range(1, ns - 3 + 1), False)
            haar_scales(got, low, False)
        assert np.array_equal(got, want), (n, forward)


@pytest.mark.parametrize("n", [5, 16, 67, 130])
def test_model_matches_the_oracle(oracle, n):
    rng = np.random.default_rng(100 + n)
    x = rng.standard_normal(n)
    ns = nscale_of(n)
    assert np.array_equal(haar_scales(x.copy(), range(1, ns + 1), True), oracle.forward_wavelet(x.copy(), 1, n, 1, 1))
    assert np.array_equal(haar_scales(x.copy(), range(1, ns + 1), False), oracle.inverse_wavelet(x.copy(), 1, n, 1, 1))
