"""Host model of the T16 builder's bank dealing (csrc/t16.cu, t16_fill_kernel): the closed-form position of an entry in
the dealt order and the packet-row permutation, restated in numpy. Checks the invariants the kernels rely on -- the
mapping is a bijection onto the segment, and the 16 entries a half-warp gathers with one instruction hit 16 different
shared-memory banks as long as every bank class still has entries. (The device code itself is covered by the product
parity tests; this pins the index arithmetic on the CPU.)"""
import numpy as np
import pytest


def swz(i):
    return i ^ (((i >> 4) ^ (i >> 8) ^ (i >> 12)) & 15)


def dealt_positions(keys, stride):
    """keys: ascending in-tile indices of one long segment. Returns (pos, bank): the slot of every entry in the
    segment and its bank class, following t16_fill_kernel."""
    kswz = swz(np.asarray(keys, dtype=np.int64))
    c = kswz & 15
    n = c.size
    cnt = np.bincount(c, minlength=16)
    g = np.zeros(n, dtype=np.int64)                         # index of the entry inside its class, in segment order
    run = np.zeros(16, dtype=np.int64)
    for k in range(n):
        g[k] = run[c[k]]
        run[c[k]] += 1
    q = np.array([np.minimum(cnt, g[k]).sum() + np.count_nonzero(cnt[:c[k]] > g[k]) for k in range(n)])
    blk = 16 * stride
    nfull = n // blk * blk
    pos = np.where(q < nfull, q // blk * blk + stride * (q & 15) + ((q >> 4) % stride), q)
    return pos, c, cnt


@pytest.mark.parametrize("stride", [2, 4])
@pytest.mark.parametrize("n,tile", [(300, 8192), (1000, 8192), (4097, 16384), (257, 512)])
def test_dealing_is_a_bijection_and_conflict_free_while_all_classes_last(stride, n, tile):
    rng = np.random.default_rng(n + stride)
    keys = np.sort(rng.choice(tile, size=n, replace=False))
    pos, bank, cnt = dealt_positions(keys, stride)
    assert np.array_equal(np.sort(pos), np.arange(n))        # every slot of the segment is written exactly once
    slot_bank = np.empty(n, dtype=np.int64)
    slot_bank[pos] = bank
    # a half-warp reads, for component c of packet row r, the slots r*blk + stride*j + c, j = 0..15
    blk = 16 * stride
    full_rounds = cnt.min()                                  # rounds of the dealing in which all 16 classes take part
    clean = 0
    for r in range(n // blk):
        for comp in range(stride):
            unit = r * blk + stride * np.arange(16) + comp
            first_q = r * blk + 16 * comp                    # the unit holds dealt entries first_q .. first_q + 15
            if first_q + 16 <= 16 * full_rounds:
                assert len(set(slot_bank[unit])) == 16
                clean += 1
    assert clean >= (16 * full_rounds) // 16 - stride        # essentially every unit of the balanced part


def test_wavelet_strides_do_not_collapse_onto_one_bank():
    """Coefficients of one wavelet level sit at power-of-two strides; the XOR swizzle of the tile placement spreads them
    over the bank classes so that the dealing has something to deal."""
    for level in range(1, 10):
        keys = np.arange(2 ** (level - 1), 8192, 2 ** level)
        if keys.size < 64:
            continue
        classes = np.bincount(swz(keys) & 15, minlength=16)
        assert classes.max() <= 2.1 * classes.mean() + 2, (level, classes)
