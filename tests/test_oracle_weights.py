"""Oracle depth / distance weighting (oracle/tfx_oracle.c orc_depth_weight, weights_gravmag.f90:46-250) against an
independent vectorised numpy derivation of the same formulas ("parity unpinned" by reference tests)."""
import numpy as np
import pytest

from tests.synth import regular_grid, station_lattice


def _setup():
    grid = regular_grid(7, 6, 5)
    rng = np.random.default_rng(2)
    grid[5] = grid[4] + rng.uniform(20.0, 80.0, grid[4].size)        # cells of different thickness
    xd, yd, zd = station_lattice(19, 700.0, 600.0, z=-3.0)
    return grid, xd, yd, zd


def _finish(w, vol):
    w = w * np.sqrt(vol)                  # :170-175
    w = w / w.max()                       # normalize_depth_weight (:228-250)
    return 1.0 / w                        # :189-195


@pytest.mark.parametrize("power", [2.0, 3.0])
def test_distance_weighting(oracle, power):
    grid, xd, yd, zd = _setup()
    X1, X2, Y1, Y2, Z1, Z2 = grid
    beta, R0, f = 1.5, 0.1, 0.25
    vol = np.abs((X2 - X1) * (Y2 - Y1) * (Z2 - Z1))
    px = np.stack([X1 + f * np.abs(X2 - X1), X2 - f * np.abs(X2 - X1)])       # 2 points per axis inside the cell
    py = np.stack([Y1 + f * np.abs(Y2 - Y1), Y2 - f * np.abs(Y2 - Y1)])
    pz = np.stack([Z1 + f * np.abs(Z2 - Z1), Z2 - f * np.abs(Z2 - Z1)])
    wr = np.zeros(X1.size)
    for j in range(xd.size):
        integral = np.zeros(X1.size)
        for a in range(2):
            for b in range(2):
                for c in range(2):
                    R = np.sqrt((px[a] - xd[j]) ** 2 + (py[b] - yd[j]) ** 2 + (pz[c] - zd[j]) ** 2)
                    integral += 1.0 / (R + R0) ** power
        wr += (integral * vol / 8.0) ** 2                                        # Li & Oldenburg (2000), Eq. 19
    want = _finish((1.0 / np.sqrt(vol)) * wr ** (beta / 4.0), vol)
    got = oracle.depth_weight(2, grid, xd, yd, zd, power, beta, 0.0)
    assert np.allclose(got, want, rtol=1e-12)
    assert got.min() == pytest.approx(1.0, rel=1e-14)                            # normalised: the largest weight is 1


def test_minimum_distance_weighting(oracle):
    grid, xd, yd, zd = _setup()
    X1, X2, Y1, Y2, Z1, Z2 = grid
    vol = np.abs((X2 - X1) * (Y2 - Y1) * (Z2 - Z1))
    cx, cy, cz = 0.5 * (X1 + X2), 0.5 * (Y1 + Y2), 0.5 * (Z1 + Z2)
    d = np.sqrt((cx[:, None] - xd) ** 2 + (cy[:, None] - yd) ** 2 + (cz[:, None] - zd) ** 2).min(axis=1)
    want = _finish(np.sqrt(1.0 / (d + 0.01) ** 2.5), vol)
    got = oracle.depth_weight(3, grid, xd, yd, zd, 2.5, 1.0, 0.0)
    assert np.allclose(got, want, rtol=1e-13)


def test_depth_weighting_and_its_aborts(oracle):
    grid, xd, yd, zd = _setup()
    X1, X2, Y1, Y2, Z1, Z2 = grid
    vol = np.abs((X2 - X1) * (Y2 - Y1) * (Z2 - Z1))
    want = _finish((0.5 * (Z1 + Z2) + 12.5) ** (-3.0 / 2.0), vol)
    assert np.allclose(oracle.depth_weight(1, grid, xd, yd, zd, 3.0, 1.0, 12.5), want, rtol=1e-13)
    with pytest.raises(RuntimeError):
        oracle.depth_weight(1, grid, xd, yd, zd, 3.0, 1.0, -1.0e4)              # non-positive depth (:218-221)
    with pytest.raises(RuntimeError):
        oracle.depth_weight(4, grid, xd, yd, zd, 3.0)                           # unknown type (:163-165)
