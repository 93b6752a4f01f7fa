"""Generates tests/golden/mansf_slice.npz from the reference's input fixtures for config A
(parfiles/Parfile_mansf_slice.txt). Run once in the development container:

    python tests/golden/make_mansf_fixture.py /root/reference

The model grid (2 x 128 x 32 cells of 127 x 127 x 90 m starting at x = 8001) and the 256 stations are
exactly regular, so only their generating parameters are stored; the script asserts that the regular
reconstruction reproduces the reference files bit for bit. The true model has three lithologies
(0 / 110 / 240 kg/m3) and is stored as float64 values.
"""
import os
import sys

import numpy as np


def main(ref):
    d = os.path.join(ref, "data", "gravmag", "mansf_slice")
    vals = np.loadtxt(os.path.join(d, "true_model_grav_3litho-values.txt"), skiprows=1)
    grid = np.loadtxt(os.path.join(d, "true_model_grav_3litho-grid.txt"), skiprows=1)
    data = np.loadtxt(os.path.join(d, "data_grid.txt"), skiprows=1)
    nx, ny, nz = 2, 128, 32
    k, j, i = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    assert np.array_equal(grid[:, 0], 8001.0 + 127.0 * i.ravel()) and np.array_equal(grid[:, 1], grid[:, 0] + 127.0)
    assert np.array_equal(grid[:, 2], 127.0 * j.ravel()) and np.array_equal(grid[:, 3], grid[:, 2] + 127.0)
    assert np.array_equal(grid[:, 4], 90.0 * k.ravel()) and np.array_equal(grid[:, 5], grid[:, 4] + 90.0)
    ys, xs = np.meshgrid(63.5 + 127.0 * np.arange(128), np.array([8064.5, 8191.5]), indexing="ij")
    assert np.array_equal(data[:, 0], xs.ravel()) and np.array_equal(data[:, 1], ys.ravel())
    assert np.all(data[:, 2] == -0.1)
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "mansf_slice.npz")
    np.savez_compressed(out, model=vals.astype(np.float64), nx=nx, ny=ny, nz=nz, x0=8001.0, dx=127.0, dy=127.0,
                        dz=90.0, station_x=np.array([8064.5, 8191.5]), station_y0=63.5, station_dy=127.0,
                        station_z=-0.1, admm_bounds=np.array([-20., 20., 90., 130., 220., 260.]))
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
