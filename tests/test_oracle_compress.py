"""Oracle row compression (oracle/tfx_oracle.c orc_compress_row, sensitivity_gravmag.F90:230-311) against the rules it
restates, on hand-made lines with compression_type = 0 in the transform slot replaced by an identity-like case:
threshold = |x| of rank N - nel_compressed in ascending |x| order (:244-249), floor 1e-30 (:252-256), strict `>`
(:261), real(4) values (:265)."""
import numpy as np
import pytest


def _compress(oracle, line, nel, ctype=1, shape=None):
    n = line.size
    nx, ny, nz = shape or (n, 1, 1)
    return oracle.compress_row(np.array(line, dtype=np.float64), nx, ny, nz, ctype, nel)


def test_threshold_is_the_kth_value_and_the_comparison_is_strict(oracle):
    # n1 = n2 = n3 = 1 would leave nothing to transform; use a 1 x 1 x n line with Haar and undo the transform instead:
    # choose the WAVELET coefficients, inverse-transform them into the line, so the pipeline sees known coefficients.
    rng = np.random.default_rng(0)
    n = 64
    coef = rng.permutation(np.arange(1, n + 1)).astype(np.float64) * rng.choice([-1.0, 1.0], n)   # |c| = 1..64, distinct
    line = oracle.inverse_wavelet(coef.copy(), n, 1, 1, 1)
    nel = 10
    r = _compress(oracle, line, nel, 1, (n, 1, 1))
    back = r["line_w"]                                          # transformed line as the pipeline saw it
    assert np.allclose(back, coef, rtol=1e-12)
    # sorted ascending |c|: position N - nel (1-based) is the threshold; entries strictly above it are kept
    assert r["threshold"] == pytest.approx(n - nel, rel=1e-12)
    kept = np.sort(np.flatnonzero(np.abs(back) > r["threshold"]) + 1)
    assert np.array_equal(r["cols"], kept) and len(kept) == nel
    assert r["vals"].dtype == np.float32 and np.array_equal(r["vals"], back[kept - 1].astype(np.float32))
    assert r["cost_full"] == pytest.approx(np.sum(back ** 2), rel=1e-13)
    assert r["cost_discarded"] == pytest.approx(np.sum(back[np.abs(back) <= r["threshold"]] ** 2), rel=1e-13)


def test_ties_at_the_threshold_are_dropped(oracle):
    """A line of period 2 gives eight bit-identical first-level detail coefficients (same arithmetic on the same
    inputs), a single coarse coefficient and exact zeros elsewhere: when the rank of the threshold lands on the tied
    value, the strict comparison (:261) drops ALL of them and fewer than nel_compressed entries are stored."""
    n = 16
    line = np.tile([1.0, 3.0], n // 2)
    r = _compress(oracle, line, 4, 1, (n, 1, 1))
    w = r["line_w"]
    details = w[1::2]
    assert np.all(details == details[0]) and details[0] != 0.0           # exact ties
    assert np.count_nonzero(w) == 9                                        # 8 details + the coarse coefficient
    assert r["threshold"] == abs(details[0])
    assert len(r["cols"]) == 1 and r["cols"][0] == 1                       # only the coarse coefficient is strictly above
    assert r["cost_discarded"] == pytest.approx(8 * details[0] ** 2, rel=1e-14)


def test_threshold_floor_and_keep_everything(oracle):
    n = 16
    coef = np.zeros(n)
    coef[3], coef[9] = 2.0, -1.0e-31                            # one real entry, one below the 1e-30 floor
    line = oracle.inverse_wavelet(coef.copy(), n, 1, 1, 1)
    r = _compress(oracle, line, 8, 1, (n, 1, 1))               # rank lands among the (numerically) zero entries
    assert r["threshold"] == pytest.approx(1.0e-30)
    assert len(r["cols"]) >= 1 and np.abs(r["vals"]).max() == pytest.approx(2.0, rel=1e-6)
    assert np.all(np.abs(r["line_w"][r["cols"] - 1]) > 1.0e-30)
    # nel_compressed >= N keeps every entry above the floor (:245-247)
    coef = np.arange(1.0, n + 1.0)
    line = oracle.inverse_wavelet(coef.copy(), n, 1, 1, 1)
    r = _compress(oracle, line, n, 1, (n, 1, 1))
    assert len(r["cols"]) == n
