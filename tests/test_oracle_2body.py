"""Oracle-only look at config D (Parfile_2body_induced: 67 x 67 x 30 padded grid, 1681 stations, 3 magnetisation
components, TMI data, compression rate 0.3) from the committed fixture tests/golden/twobody_induced.npz (generated from
the reference's input files by tests/golden/make_2body_fixture.py). Pins the magnetic kernel + compression restatement
against the surveyor's sanity figures of SURVEY.md section 8a/a12 (stations 1 and 841, no depth weight: 40 401 entries
kept per component, compression error ~1-4e-5 with Haar, ~1-4e-6 with D4) -- NOT reference output."""
import os

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "twobody_induced.npz")


@pytest.fixture(scope="module")
def cfg():
    z = np.load(GOLDEN)
    nx, ny, nz = int(z["nx"]), int(z["ny"]), int(z["nz"])
    k, j, i = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    i, j, k = i.ravel(), j.ravel(), k.ravel()
    xn, yn, zn = z["xn"], z["yn"], z["zn"]
    grid = [np.ascontiguousarray(a) for a in (xn[i], xn[i + 1], yn[j], yn[j + 1], zn[k], zn[k + 1])]
    ns = int(z["nstations_side"])
    sy, sx = np.meshgrid(float(z["station_y0"]) + float(z["station_dy"]) * np.arange(ns),
                         float(z["station_x0"]) + float(z["station_dx"]) * np.arange(ns), indexing="ij")
    return dict(nx=nx, ny=ny, nz=nz, N=nx * ny * nz, grid=grid, sx=sx.ravel(), sy=sy.ravel(), sz=float(z["station_z"]),
                mi=float(z["inclination"]), md=float(z["declination"]), theta=float(z["xaxis_declination"]),
                intensity=float(z["intensity_nT"]), rate=float(z["compression_rate"]), xn=xn, yn=yn, zn=zn)


def test_fixture_shape(cfg):
    assert cfg["N"] == 134670 and cfg["sx"].size == 1681
    assert int(cfg["rate"] * cfg["N"]) == 40401                       # nel_compressed (sensitivity_gravmag.F90:64-77)
    # padded grid: 50 m core cells, growing towards the rim; every station is above the top face (no inside-cell branch)
    assert np.isclose(np.diff(cfg["xn"]).min(), 50.0) and np.diff(cfg["xn"]).max() > 200.0
    assert cfg["sz"] < cfg["zn"].min()


@pytest.mark.parametrize("station", [1, 841])
def test_compression_error_haar_vs_d4(oracle, cfg, station):
    i = station - 1
    lines = oracle.magprism(cfg["grid"], float(cfg["sx"][i]), float(cfg["sy"][i]), cfg["sz"], 3, 1, cfg["mi"], cfg["md"],
                            cfg["theta"], cfg["intensity"])
    assert lines.shape == (1, 3, cfg["N"]) and np.all(np.isfinite(lines))
    nel = int(cfg["rate"] * cfg["N"])
    err = {}
    for ctype in (1, 2):
        e = []
        for kcomp in range(3):
            r = oracle.compress_row(lines[0, kcomp].copy(), cfg["nx"], cfg["ny"], cfg["nz"], ctype, nel)
            assert len(r["cols"]) == nel                               # no ties at the threshold: exactly 40 401 kept
            assert np.all(np.diff(r["cols"]) > 0)
            e.append(np.sqrt(r["cost_discarded"] / r["cost_full"]))
        err[ctype] = np.array(e)
    assert np.all((err[1] > 5e-6) & (err[1] < 8e-5)), err             # Haar: ~1-4e-5
    assert np.all((err[2] > 5e-7) & (err[2] < 8e-6)), err             # D4: ~1-4e-6
    assert np.all(err[2] < 0.3 * err[1])                               # the reason config D uses D4
