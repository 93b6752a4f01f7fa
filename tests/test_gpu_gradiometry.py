"""Gravity gradiometry kernels on the device (data_type = 2, sensitivity_gravmag.F90:199-213): gradiprism_zz
(gravity_field.f90:314-364, one data component) and gradiprism_full (:207-309, six data components XX, YY, ZZ, XY, YZ, ZX)
against the oracle, raw lines and through the compressed row pipeline (one matrix row per station and component,
sensitivity_gravmag.F90:759-762)."""
import numpy as np
import pytest

import tomofastx_b200 as tfx
from tests.synth import make_problem

pytestmark = pytest.mark.gpu


def _pb(ndc, ctype=0, rate=0.25):
    pb = make_problem(nx=9, ny=8, nz=5, ndata=7, compression_type=ctype, rate=rate, ndata_components=ndc)
    pb.par.data_type = 2
    return pb


def test_gzz_lines(oracle):
    pb = _pb(1)
    lines = tfx.sensit_lines(pb.par, pb.grid, pb.data_xyz)                      # (station, 1, 1, cell)
    for i in range(pb.ndata):
        want = oracle.gradiprism_zz(pb.grid, *(float(a[i]) for a in pb.data_xyz))
        assert np.abs(lines[i, 0, 0] - want).max() <= 1e-11 * np.abs(want).max()


def test_full_tensor_lines(oracle):
    pb = _pb(6)
    lines = tfx.sensit_lines(pb.par, pb.grid, pb.data_xyz)                      # (station, 6, 1, cell)
    assert lines.shape == (pb.ndata, 6, 1, pb.N)
    for i in range(pb.ndata):
        want = oracle.gradiprism_full(pb.grid, *(float(a[i]) for a in pb.data_xyz))
        for d in range(6):
            assert np.abs(lines[i, d, 0] - want[d]).max() <= 1e-11 * np.abs(want[d]).max(), (i, d)
    # the zz component of the full tensor is gradiprism_zz
    pz = _pb(1)
    assert np.array_equal(tfx.sensit_lines(pz.par, pz.grid, pz.data_xyz)[:, 0], lines[:, 2])


@pytest.mark.parametrize("ctype", [0, 1])
def test_full_tensor_matrix_rows(oracle, ctype):
    """calculate_sensit with six data components: row (idata - 1)*6 + d holds component d of station idata."""
    pb = _pb(6, ctype)
    pb.dw = np.linspace(0.5, 1.5, pb.ndata * 6).reshape(pb.ndata, 6)
    S, nnz_col, cerr, tot = tfx.calculate_sensit(pb.par, pb.grid, pb.data_xyz, pb.cw, pb.dw)
    nel = pb.nel_compressed
    assert S.get_total_row_number() == 6 * pb.ndata
    assert abs(tot - 6 * pb.ndata * nel) <= 2 * 6 * pb.ndata
    sa, ija, ijl, rowptr = S.export()
    rows = {int(rowptr[i]): (ija[ijl[i] - 1:ijl[i + 1] - 1], sa[ijl[i] - 1:ijl[i + 1] - 1]) for i in range(len(rowptr))}
    flips = 0
    for i in range(pb.ndata):
        want = oracle.gradiprism_full(pb.grid, *(float(a[i]) for a in pb.data_xyz))
        for d in range(6):
            r = oracle.compress_row(want[d] * pb.cw, pb.nx, pb.ny, pb.nz, ctype, nel)
            cols, vals = rows[i * 6 + d + 1]
            diff = set(cols) ^ set(r["cols"])
            flips += len(diff)
            if not diff:
                w = (r["vals"] * np.float32(pb.dw[i, d])).astype(np.float32)
                assert np.allclose(vals, w, rtol=3e-6, atol=1e-6 * np.abs(w).max())
    assert flips <= 2 * 6 * pb.ndata
    x = np.random.default_rng(0).standard_normal(pb.ncolumns)
    assert np.all(np.isfinite(S.mult_vector(x)))


def test_gradiometry_aborts():
    pb = _pb(3)
    with pytest.raises(tfx.TfxError, match="Wrong number of gravity gradiometry data components"):
        tfx.sensit_lines(pb.par, pb.grid, pb.data_xyz)
    pb = _pb(6)
    x, y, z = pb.data_xyz
    x[0], y[0], z[0] = 0.0, 0.0, 300.0          # on the vertical line through a grid edge, below the grid
    with pytest.raises(tfx.TfxError, match="gradiprism_full"):
        tfx.sensit_lines(pb.par, pb.grid, (x, y, z))
