"""calculate_depth_weight on the device (csrc/weights.cu) against the oracle restatement of
weights_gravmag.f90:46-250 (depth, distance and minimum-distance weighting)."""
import numpy as np
import pytest

import tomofastx_b200 as tfx
from tests.synth import depth_weight_type1, regular_grid, station_lattice

pytestmark = pytest.mark.gpu

# f64 throughout; CUDA's pow is accurate to 2 ulp (libm: < 1 ulp) and type 2 sums ndata terms in the same order
RTOL = 1e-13


def _irregular_grid(nx, ny, nz, seed=3):
    g = regular_grid(nx, ny, nz)
    rng = np.random.default_rng(seed)
    g[5] = g[4] + rng.uniform(20.0, 80.0, g[4].size)          # Z2: cells of different thickness / volume
    return g


@pytest.mark.parametrize("wtype,power,beta,Z0", [(1, 2.0, 1.0, 0.0), (1, 3.0, 1.0, 12.5), (2, 3.0, 1.5, 0.0),
                                                  (2, 2.0, 1.0, 0.0), (3, 2.0, 1.0, 0.0), (3, 3.5, 1.0, 0.0)])
def test_depth_weight_vs_oracle(oracle, wtype, power, beta, Z0):
    grid = _irregular_grid(9, 7, 5)
    xd, yd, zd = station_lattice(23, 900.0, 700.0, z=-5.0)
    want = oracle.depth_weight(wtype, grid, xd, yd, zd, power, beta, Z0)
    got = tfx.calculate_depth_weight(wtype, grid, (xd, yd, zd), power, beta, Z0)
    assert np.allclose(got, want, rtol=RTOL, atol=0.0)
    if wtype == 1 and Z0 == 0.0:
        assert np.allclose(got, depth_weight_type1(grid, power), rtol=RTOL)


def test_depth_weight_slabs_match_full(oracle):
    """A rank's slab (nsmaller, nelements) is the same cells of the full result up to the normalisation constant
    (the maximum is global in the reference: mpi_allreduce MAX, weights_gravmag.f90:237-240)."""
    grid = _irregular_grid(8, 6, 4)
    xd, yd, zd = station_lattice(700, 800.0, 600.0, z=-0.1)    # > one shared-memory chunk of stations
    full = tfx.calculate_depth_weight(2, grid, (xd, yd, zd), 3.0, 1.5)
    want = oracle.depth_weight(2, grid, xd, yd, zd, 3.0, 1.5)
    assert np.allclose(full, want, rtol=RTOL)
    n0, n = 50, 77
    slab = tfx.calculate_depth_weight(2, grid, (xd, yd, zd), 3.0, 1.5, nsmaller=n0, nelements=n)
    ratio = slab / full[n0:n0 + n]
    assert np.allclose(ratio, ratio[0], rtol=1e-14)
    assert slab.min() == pytest.approx(1.0, rel=1e-15)           # normalised by the slab's own maximum


def test_depth_weight_device_pointers():
    grid = regular_grid(16, 16, 8)
    xd, yd, zd = station_lattice(64, 1600.0, 1600.0)
    want = tfx.calculate_depth_weight(2, grid, (xd, yd, zd), 2.0, 1.0)
    bufs = []
    for a in list(grid) + [xd, yd, zd]:
        b = tfx.Buffer(a.size); tfx.copy(b, a, a.size); bufs.append(b)
    out = tfx.Buffer(grid[0].size)
    tfx.calculate_depth_weight(2, bufs[:6], bufs[6:], 2.0, 1.0, column_weight=out)
    assert np.array_equal(out.numpy(), want)


def test_depth_weight_aborts():
    grid = regular_grid(4, 4, 2)
    xyz = station_lattice(4, 400.0, 400.0)
    with pytest.raises(tfx.TfxError, match="Not known depth weight type"):
        tfx.calculate_depth_weight(4, grid, xyz, 2.0)
    with pytest.raises(tfx.TfxError, match="non-positive depth"):
        tfx.calculate_depth_weight(1, grid, xyz, 2.0, Z0=-1000.0)
