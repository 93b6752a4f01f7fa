"""get_load_balancing_nelements (sensitivity_gravmag.F90:470-524): the C-ABI host routine against the
oracle restatement; no GPU needed (integer work on the host)."""
import numpy as np
import pytest

import tomofastx_b200 as tfx
from oracle import partition as orp


@pytest.mark.parametrize("nbproc", [1, 2, 3, 5, 8])
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_load_balancing_matches_oracle(nbproc, seed):
    rng = np.random.default_rng(seed)
    n = 400
    nnz = rng.integers(0, 50, n).astype(np.int32)
    nnz[: n // 16] += 300                       # coarse wavelet scales are dense columns
    if seed == 2:
        nnz[n // 2:] = 0                        # a long empty tail: the last ranks still get cells
    try:
        want = orp.get_load_balancing_nelements(nnz, nbproc)
    except RuntimeError as e:
        with pytest.raises(tfx.TfxError, match="get_load_balancing_nelements"):
            tfx.get_load_balancing_nelements(nnz, nbproc)
        return
    got = tfx.get_load_balancing_nelements(nnz, nbproc)
    assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1])
    assert got[1].sum() == n and got[0].sum() == int(nnz.sum())
    if nbproc > 1 and seed != 2:
        # the point of the exercise: nnz per rank within one column of the ideal share
        assert np.abs(got[0] - nnz.sum() / nbproc).max() <= nnz.max() + nbproc


def test_load_balancing_error_when_a_rank_gets_nothing():
    nnz = np.zeros(4, dtype=np.int32)
    nnz[3] = 10                                  # everything in the last cell: ranks cannot all be served
    with pytest.raises(RuntimeError):
        orp.get_load_balancing_nelements(nnz, 3)
    with pytest.raises(tfx.TfxError, match="Wrong cpu in get_load_balancing_nelements"):
        tfx.get_load_balancing_nelements(nnz, 3)


def test_partition_from_a_strided_station_sample_balances_the_full_kernel(oracle):
    """The row-blocked assembly of the big compressed configs (bench.py --comp-batch) has to fix the
    column partition BEFORE the kernel exists, so it balances on the per-column nnz counts of a strided ~1/16 sample of
    the stations (the reference has the counts of all rows on disk first, sensitivity_gravmag.F90:381-392,610-625).
    On a regular station lattice the sample must predict the full distribution: the partition computed from it
    leaves the ranks of the FULL kernel within a few per cent of the mean."""
    from tests.synth import depth_weight_type1, regular_grid, station_lattice
    nx, ny, nz, nd, rate = 16, 16, 8, 256, 0.1
    N = nx * ny * nz
    grid = regular_grid(nx, ny, nz)
    xyz = station_lattice(nd, 100.0 * nx, 100.0 * ny, z=-0.1)
    cw = depth_weight_type1(grid, 2.0, 0.0, 4.0e3)
    nel = int(rate * N)
    counts = np.zeros((nd, N), dtype=np.int32)
    for i in range(nd):
        line = oracle.graviprism_z(grid, *(float(a[i]) for a in xyz)) * cw
        r = oracle.compress_row(line, nx, ny, nz, 1, nel)
        counts[i, r["cols"] - 1] = 1
    full = counts.sum(axis=0)
    sample = counts[::17].sum(axis=0)      # odd stride, like bench.py: no aliasing with the 16-wide station lattice
    for nbproc in (2, 4, 8):
        _, nel_at = orp.get_load_balancing_nelements(sample, nbproc)
        cum = np.concatenate([[0], np.cumsum(nel_at)])
        per_rank = np.array([full[cum[r]:cum[r + 1]].sum() for r in range(nbproc)])
        assert per_rank.sum() == full.sum()
        assert per_rank.max() / per_rank.mean() < 1.10, (nbproc, per_rank)
