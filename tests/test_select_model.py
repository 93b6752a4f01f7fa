"""Host model (numpy) of the batched row pipeline's integer logic in csrc/sensit.cu: the exact k-th order statistic of |x|
by MSD radix select on the IEEE bit patterns (two 12-bit passes -> candidate list -> 8-bit passes on the list, or further
passes over the line when the bucket does not fit the list) and the chunked ordered compaction. Same digits, same pick
rule, same fallback schedule as the kernels; checked against np.sort / np.nonzero on adversarial lines (ties, zeros,
signed zeros, subnormals, a single exponent). Reference rule: threshold = the (N - nel_compressed)-th smallest |x|, floored
at 1e-30, keep |x| > threshold strictly, columns ascending (sensitivity_gravmag.F90:240-272)."""
import numpy as np
import pytest

MASK63 = np.uint64(0x7FFFFFFFFFFFFFFF)


def _pick(hist, rank):
    """k_sel_pick: the digit whose bucket holds the wanted rank, the rank inside it, the bucket size."""
    cum = np.cumsum(hist)
    d = int(np.searchsorted(cum, rank, side="right"))
    return d, rank - (int(cum[d - 1]) if d > 0 else 0), int(hist[d])


def kth_abs_bits(x, rank, cand_cap=16384):
    b = x.view(np.uint64) & MASK63
    prefix, himask = np.uint64(0), np.uint64(0)
    passes_over_line = 0
    for shift, bits in ((51, 12), (39, 12)):                      # k_sel_hist12<1>, k_sel_hist12<2>
        act = (b & himask) == prefix
        digit = (b[act] >> np.uint64(shift)) & np.uint64((1 << bits) - 1)
        d, rank, count = _pick(np.bincount(digit.astype(np.int64), minlength=1 << bits), rank)
        prefix |= np.uint64(d) << np.uint64(shift)
        himask |= np.uint64(((1 << bits) - 1) << shift)
        passes_over_line += 1
    if count <= cand_cap:                                         # k_sel_collect + k_sel_finish
        cand = b[(b & himask) == prefix]
        passes_over_line += 1
        assert cand.size == count
        sft = 39
        while sft > 0:
            bits = min(8, sft)
            sft -= bits
            act = (cand & himask) == prefix
            digit = (cand[act] >> np.uint64(sft)) & np.uint64((1 << bits) - 1)
            d, rank, _ = _pick(np.bincount(digit.astype(np.int64), minlength=1 << bits), rank)
            prefix |= np.uint64(d) << np.uint64(sft)
            himask |= np.uint64(((1 << bits) - 1) << sft)
    else:                                                         # the kRest schedule of assemble_rows_device
        for shift, bits in ((27, 12), (15, 12), (3, 12), (0, 3)):
            act = (b & himask) == prefix
            digit = (b[act] >> np.uint64(shift)) & np.uint64((1 << bits) - 1)
            d, rank, count = _pick(np.bincount(digit.astype(np.int64), minlength=1 << bits), rank)
            prefix |= np.uint64(d) << np.uint64(shift)
            himask |= np.uint64(((1 << bits) - 1) << shift)
            passes_over_line += 1
    return np.array([prefix], dtype=np.uint64).view(np.float64)[0], passes_over_line


def compact(x, thr, chunk=4096):
    """k_cmp_count / k_cmp_scan / k_cmp_write: per-chunk counts, exclusive scan, ordered write."""
    n = x.size
    nchunks = (n + chunk - 1) // chunk
    cnt = np.array([np.count_nonzero(np.abs(x[c * chunk:(c + 1) * chunk]) > thr) for c in range(nchunks)])
    off = np.concatenate([[0], np.cumsum(cnt)[:-1]])
    out = np.full(int(cnt.sum()), -1, dtype=np.int64)
    for c in range(nchunks):
        seg = x[c * chunk:(c + 1) * chunk]
        # 16 rows of 256 threads; inside a row the ballot rank, rows and warps in element order
        keep = np.abs(seg) > thr
        pos = off[c] + np.cumsum(keep) - keep
        out[pos[keep]] = c * chunk + np.nonzero(keep)[0]
    return out


def _lines(rng):
    n = 50_000
    yield "lognormal", rng.standard_normal(n) * 10.0 ** rng.uniform(-12, 3, n)
    yield "one_exponent", rng.uniform(1.0, 2.0, n) * rng.choice([-1.0, 1.0], n)          # all in one first-pass bucket
    ties = np.round(rng.standard_normal(n), 1)                                            # ~80 distinct values
    yield "ties", ties
    z = rng.standard_normal(n); z[rng.random(n) < 0.7] = 0.0; z[::7] = -0.0
    yield "mostly_zero", z
    yield "all_equal", np.full(n, -3.25)
    sub = rng.standard_normal(n) * 1e-310
    yield "subnormal", sub
    yield "short", rng.standard_normal(37)


@pytest.mark.parametrize("cand_cap", [16384, 1])
def test_radix_select_is_the_exact_order_statistic(cand_cap):
    rng = np.random.default_rng(7)
    for name, x in _lines(rng):
        x = np.ascontiguousarray(x, dtype=np.float64)
        n = x.size
        srt = np.sort(np.abs(x))
        for rate in (0.05, 0.3, 0.999):
            nel = int(rate * n)
            if nel >= n or nel < 1:
                continue
            rank = n - nel - 1                                   # 0-based rank of sorted(N - nel_compressed)
            got, passes = kth_abs_bits(x, rank, cand_cap)
            assert got == srt[rank], (name, rate)
            assert passes <= 3 or cand_cap == 1 or name in ("ties", "mostly_zero", "all_equal"), (name, passes)
            thr = max(got, 1e-30)
            cols = compact(x, thr)
            assert np.array_equal(cols, np.nonzero(np.abs(x) > thr)[0]), name
            assert cols.size <= nel                               # strict '>' never keeps more than nel_compressed
