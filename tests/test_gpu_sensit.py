"""Device sensitivity assembly (csrc/assembly.cu, csrc/sensit.cu) vs the oracle.

The forward kernels are sums of 8 signed corner terms with heavy cancellation; CUDA's log/atan2 differ
from glibc's in the last ulp, which the cancellation amplifies. Tolerances are therefore relative to the
largest |value| of the line (stated per test). No reference test covers these kernels ("parity unpinned"):
the oracle is a line-by-line restatement."""
import numpy as np
import pytest

import tomofastx_b200 as tfx
from tests.synth import make_problem

pytestmark = pytest.mark.gpu


def test_gravity_lines_vs_oracle(oracle):
    pb = make_problem(nx=12, ny=10, nz=6, ndata=9)
    got = tfx.sensit_lines(pb.par, pb.grid, pb.data_xyz)
    for i in range(pb.ndata):
        want = oracle.graviprism_z(pb.grid, *(float(a[i]) for a in pb.data_xyz))
        err = np.abs(got[i, 0, 0] - want).max() / np.abs(want).max()
        assert err < 1e-12, err


@pytest.mark.parametrize("nmc,ndc", [(1, 1), (3, 1), (1, 3), (3, 3)])
def test_magnetic_lines_vs_oracle(oracle, nmc, ndc):
    pb = make_problem(nx=9, ny=8, nz=5, ndata=6, problem_type=2, nmodel_components=nmc, ndata_components=ndc)
    got = tfx.sensit_lines(pb.par, pb.grid, pb.data_xyz)
    p = pb.par
    for i in range(pb.ndata):
        want = oracle.magprism(pb.grid, *(float(a[i]) for a in pb.data_xyz), nmc, ndc, p.mi, p.md, p.theta, p.intensity)
        err = np.abs(got[i] - want).max() / np.abs(want).max()
        assert err < 1e-11, err


def test_magnetic_station_inside_cell(oracle):
    # borehole branch: six sub-prisms around a void (magnetic_field.f90:139-224)
    pb = make_problem(nx=5, ny=5, nz=4, ndata=1, problem_type=2, nmodel_components=3, ndata_components=1)
    pb.data_xyz = (np.array([237.3]), np.array([151.9]), np.array([61.0]))
    got = tfx.sensit_lines(pb.par, pb.grid, pb.data_xyz)
    p = pb.par
    want = oracle.magprism(pb.grid, 237.3, 151.9, 61.0, 3, 1, p.mi, p.md, p.theta, p.intensity)
    assert np.abs(got[0] - want).max() / np.abs(want).max() < 1e-11


def test_station_on_cell_boundary_aborts():
    pb = make_problem(nx=4, ny=4, nz=2, ndata=1)
    # station on the line through a cell edge along x (y and z on cell faces): for the cells ahead
    # Rs + XX = |XX| + XX = 0 (gravity_field.f90:173-178)
    pb.data_xyz = (np.array([150.0]), np.array([100.0]), np.array([50.0]))
    with pytest.raises(tfx.TfxError, match="coincides with model grid boundary"):
        tfx.sensit_lines(pb.par, pb.grid, pb.data_xyz)


def _rows(sa, ija, ijl, rowptr):
    return {int(rowptr[i]): (ija[ijl[i] - 1:ijl[i + 1] - 1], sa[ijl[i] - 1:ijl[i + 1] - 1]) for i in range(len(rowptr))}


def test_dense_assembly_vs_oracle(oracle):
    pb = make_problem(nx=10, ny=9, nz=5, ndata=20, compression_type=0)
    S, nnz_col, cerr, tot = tfx.calculate_sensit(pb.par, pb.grid, pb.data_xyz, pb.cw, pb.dw)
    assert S.storage_kind() == 1 and tot == pb.ndata * pb.N and np.all(nnz_col == pb.ndata)
    So = pb.oracle_matrix(oracle)
    want = _rows(*So.arrays())
    got = _rows(*S.export())
    for r in want:
        assert np.array_equal(got[r][0], want[r][0])
        # f32 values: equal up to f32 rounding of a value that differs by ~1e-13 relative in f64
        assert np.allclose(got[r][1], want[r][1], rtol=3e-7, atol=1e-7 * np.abs(want[r][1]).max())


@pytest.mark.parametrize("ctype", [1, 2])
def test_compressed_assembly_vs_oracle(oracle, ctype):
    pb = make_problem(nx=12, ny=10, nz=6, ndata=12, compression_type=ctype, rate=0.2)
    S, nnz_col, cerr, tot = tfx.calculate_sensit(pb.par, pb.grid, pb.data_xyz, pb.cw, pb.dw)
    assert S.storage_kind() == 0
    So = pb.oracle_matrix(oracle)
    want = _rows(*So.arrays())
    got = _rows(*S.export())
    assert set(want) == set(got)
    mismatched = 0
    for r in want:
        if np.array_equal(got[r][0], want[r][0]):
            assert np.allclose(got[r][1], want[r][1], rtol=3e-6, atol=1e-6 * np.abs(want[r][1]).max())
        else:
            # a coefficient within rounding of the threshold may flip: at most a couple per row
            mismatched += len(set(got[r][0]) ^ set(want[r][0]))
    assert mismatched <= 2 * pb.ndata, mismatched
    assert abs(tot - So.nel) <= mismatched
    # the products agree to compression-noise-free precision when the patterns agree
    x = pb.model_scaled_w(oracle)
    d_want = So.mult_vector(x)
    d_got = S.mult_vector(x)
    assert np.allclose(d_got, d_want, rtol=1e-6, atol=1e-8 * np.abs(d_want).max())


def test_magnetic_compressed_assembly_three_components(oracle):
    # config D shape in miniature: magnetisation model (3 comps), TMI data, columns shifted to problem 2
    pb = make_problem(nx=8, ny=7, nz=4, ndata=6, compression_type=2, rate=0.3, problem_type=2, nmodel_components=3)
    S, nnz_col, cerr, tot = tfx.calculate_sensit(pb.par, pb.grid, pb.data_xyz, pb.cw, pb.dw)
    So = pb.oracle_matrix(oracle)
    want = _rows(*So.arrays())
    got = _rows(*S.export())
    bad = sum(len(set(got[r][0]) ^ set(want[r][0])) for r in want)
    assert bad <= 2 * pb.ndata * 3
    assert int(min(want[1][0])) > 3 * pb.N          # param_shift(2) = nelements * ncomponents
    x = pb.model_scaled_w(oracle)
    assert np.allclose(S.mult_vector(x), So.mult_vector(x), rtol=1e-6, atol=1e-9)
