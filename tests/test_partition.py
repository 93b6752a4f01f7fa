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
