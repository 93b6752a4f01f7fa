"""Generates tests/golden/twobody_induced.npz from the reference's input fixtures for config D
(parfiles/Parfile_2body_induced.txt). Run once in the development container:

    python tests/golden/make_2body_fixture.py /root/reference

The model grid (67 x 67 x 30 cells, padded: cell sizes grow towards the rim) is a tensor product of node coordinates,
and the 1681 stations are a 41 x 41 lattice at z = -5: only the node arrays and the lattice parameters are stored; the
script asserts that the reconstruction reproduces the reference files bit for bit. The synthetic model (3 magnetisation
components per cell) is a uniform background plus two single-cell bodies: stored as the background vector and the
(cell index, 3 values) pairs of the cells that differ from it.
"""
import os
import sys

import numpy as np


def main(ref):
    d = os.path.join(ref, "data", "gravmag", "2body_magnet", "induced")
    grid = np.loadtxt(os.path.join(d, "meshgrid_padded_2depth_true-grid.txt"), skiprows=1)
    vals = np.loadtxt(os.path.join(d, "meshgrid_padded_2depth_true-values.txt"), skiprows=1)
    obs = np.loadtxt(os.path.join(d, "dummy.obs"), skiprows=1)
    nx, ny, nz = 67, 67, 30
    assert grid.shape[0] == nx * ny * nz
    xn = np.concatenate([grid[:nx, 0], grid[nx - 1:nx, 1]])
    yn = np.concatenate([grid[:nx * ny:nx, 2], grid[nx * (ny - 1):nx * (ny - 1) + 1, 3]])
    zn = np.concatenate([grid[::nx * ny, 4], grid[nx * ny * (nz - 1):nx * ny * (nz - 1) + 1, 5]])
    k, j, i = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    i, j, k = i.ravel(), j.ravel(), k.ravel()
    assert np.array_equal(grid[:, 0], xn[i]) and np.array_equal(grid[:, 1], xn[i + 1])
    assert np.array_equal(grid[:, 2], yn[j]) and np.array_equal(grid[:, 3], yn[j + 1])
    assert np.array_equal(grid[:, 4], zn[k]) and np.array_equal(grid[:, 5], zn[k + 1])
    assert np.array_equal(grid[:, 6], i + 1) and np.array_equal(grid[:, 7], j + 1) and np.array_equal(grid[:, 8], k + 1)
    ns = 41
    sy, sx = np.meshgrid(obs[0, 1] + (obs[ns, 1] - obs[0, 1]) * np.arange(ns), obs[0, 0] + (obs[1, 0] - obs[0, 0]) * np.arange(ns),
                         indexing="ij")
    assert np.array_equal(obs[:, 0], sx.ravel()) and np.array_equal(obs[:, 1], sy.ravel()) and np.all(obs[:, 2] == obs[0, 2])
    vals = vals.reshape(nx * ny * nz, -1)
    background = vals[0].copy()
    nzc = np.flatnonzero(np.any(vals != background, axis=1))
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "twobody_induced.npz")
    np.savez_compressed(out, nx=nx, ny=ny, nz=nz, xn=xn, yn=yn, zn=zn, station_x0=obs[0, 0], station_dx=obs[1, 0] - obs[0, 0],
                        station_y0=obs[0, 1], station_dy=obs[ns, 1] - obs[0, 1], station_z=obs[0, 2], nstations_side=ns,
                        model_background=background, model_cells=nzc.astype(np.int32), model_values=vals[nzc],
                        inclination=-60.0, declination=2.0, intensity_nT=55000.0, xaxis_declination=0.0,
                        compression_rate=0.3, depth_weighting=np.array([2.0, 3.0, 1.5]))
    print("wrote", out, os.path.getsize(out), "bytes;", nzc.size, "body cells, value columns:", vals.shape[1])


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
