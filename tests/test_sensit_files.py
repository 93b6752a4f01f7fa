"""The reference's on-disk sensitivity formats (csrc/sensit_io.cu) against the numpy restatement
(oracle/sensit_files.py). Metadata / nnz / depth-weight files are host-only (no GPU); the stream files go
through the device matrix (gpu-marked)."""
import os

import numpy as np
import pytest

import tomofastx_b200 as tfx
from oracle import sensit_files as osf
from tests.synth import make_problem


def _par(nx=5, ny=4, nz=3, ndata=7, problem_type=1, nmc=1, ndc=1, ctype=1):
    p = tfx.SensitParams()
    p.problem_type, p.nx, p.ny, p.nz, p.ndata = problem_type, nx, ny, nz, ndata
    p.ndata_components, p.nmodel_components, p.compression_type, p.compression_rate = ndc, nmc, ctype, 0.3
    p.problem_weight = 1.0
    return p


@pytest.mark.parametrize("problem_type", [1, 2])
def test_metadata_round_trip_and_checks(tmp_path, problem_type):
    p = _par(problem_type=problem_type)
    d = str(tmp_path / "SENSIT")
    tfx.write_sensit_metadata(p, d, 3, 2, 1.2345678901234567e-3, 123456789012, np.arange(60, dtype=np.int32))
    assert os.path.exists(os.path.join(d, "sensit_%s_meta.txt" % osf.SUFFIX[problem_type]))
    assert tfx.read_sensitivity_metadata(p, d, 2) == (3, 1.2345678901234567e-3, 123456789012)
    assert np.array_equal(osf.read_nnz(d, problem_type), np.arange(60))
    assert np.array_equal(tfx.read_sensit_nnz(p, d), np.arange(60))
    # the reference's consistency checks (sensitivity_gravmag.F90:1014-1027)
    with pytest.raises(tfx.TfxError, match="does not match the Parfile"):
        tfx.read_sensitivity_metadata(p, d, 1)
    q = _par(problem_type=problem_type, ctype=2)
    with pytest.raises(tfx.TfxError, match="Compression type is inconsistent"):
        tfx.read_sensitivity_metadata(q, d, 2)
    q = _par(problem_type=problem_type, ndata=8)
    with pytest.raises(tfx.TfxError, match="does not match the Parfile"):
        tfx.read_sensitivity_metadata(q, d, 2)
    with pytest.raises(tfx.TfxError, match="Error in opening the sensitivity metadata file"):
        tfx.read_sensitivity_metadata(p, str(tmp_path / "nowhere"), 2)


def test_reads_gfortran_style_metadata(tmp_path):
    p = _par()
    d = str(tmp_path)
    osf.write_meta(d, 1, p.nx, p.ny, p.nz, p.ndata, 4, 1, 1, 2.154e-3, 1, 1, 314368, gfortran_style=True)
    assert tfx.read_sensitivity_metadata(p, d, 1) == (4, 2.154e-3, 314368)
    osf.write_nnz(d, 1, np.arange(60)[::-1])
    assert np.array_equal(tfx.read_sensit_nnz(p, d), np.arange(60)[::-1])
    with pytest.raises(tfx.TfxError, match="Wrong file header"):
        tfx.read_sensit_nnz(_par(nx=6), d)


def test_depth_weight_round_trip(tmp_path):
    p = _par(problem_type=2)
    d = str(tmp_path / "a" / "SENSIT")
    cw = np.random.default_rng(3).uniform(0.1, 5.0, 60)
    tfx.write_depth_weight(p, d, cw)
    assert np.array_equal(osf.read_weight(d, 2), cw)            # bit-exact, big-endian real(8)
    assert np.array_equal(tfx.read_depth_weight(p, d), cw)
    osf.write_weight(d, 2, cw[::-1])
    assert np.array_equal(tfx.read_depth_weight(p, d), cw[::-1])
    with pytest.raises(tfx.TfxError, match="Depth weight file header does not match"):
        tfx.read_depth_weight(_par(problem_type=2, nx=6), d)


def _rows(sa, ija, ijl, rowptr):
    return {int(rowptr[i]): (ija[ijl[i] - 1:ijl[i + 1] - 1], sa[ijl[i] - 1:ijl[i + 1] - 1]) for i in range(len(rowptr))}


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["grav_haar", "mag3_d4"])
def test_stream_file_written_from_device_rows(tmp_path, oracle, case):
    """Device rows -> file -> numpy reader: header, record order (idata, d, k), 1-based cells, unweighted f32."""
    if case == "grav_haar":
        pb = make_problem(nx=12, ny=10, nz=6, ndata=11, compression_type=1, rate=0.2)
    else:
        pb = make_problem(nx=8, ny=7, nz=4, ndata=6, compression_type=2, rate=0.3, problem_type=2, nmodel_components=3)
    nmc, N = pb.par.nmodel_components, pb.N
    d = str(tmp_path / "SENSIT")
    rows, nnz_col, cerr, tot = tfx.sensit_assemble_rows(pb.par, pb.grid, pb.data_xyz, pb.cw, pb.dw)
    tfx.write_sensit_file(rows, d)
    tfx.write_sensit_metadata(pb.par, d, 1, 1, cerr, tot, nnz_col)
    hdr, recs = osf.read_rank_file(d, pb.par.problem_type, 1, 0)
    assert hdr == (pb.ndata, pb.ndata, N, 0, 1)
    assert [(r[0], r[2], r[1]) for r in recs] == [(i, 1, k) for i in range(1, pb.ndata + 1) for k in range(1, nmc + 1)]
    assert sum(len(r[3]) for r in recs) == tot
    # same content as the device matrix (columns (k-1)*N + cell + param_shift)
    S, _, _, _ = tfx.calculate_sensit(pb.par, pb.grid, pb.data_xyz, pb.cw, pb.dw)
    got = _rows(*S.export())
    for i in range(1, pb.ndata + 1):
        cols = np.concatenate([r[3] + (r[1] - 1) * N + pb.par.param_shift for r in recs if r[0] == i])
        vals = np.concatenate([r[4] for r in recs if r[0] == i])
        assert np.array_equal(cols, got[i][0]) and np.array_equal(vals, got[i][1])
        for r in recs:
            assert np.all(np.diff(r[3]) > 0) and r[3].min() >= 1 and r[3].max() <= N
    # and the whole chain through the file: read_sensitivity_kernel on 1 rank reproduces the matrix
    M = tfx.read_sensitivity_kernel(pb.par, d, pb.dw, 1, 1 if pb.par.problem_type == 1 else 2, [N])
    back = _rows(*M.export())
    assert set(back) == set(got)
    for i in got:
        assert np.array_equal(back[i][0], got[i][0]) and np.array_equal(back[i][1], got[i][1])
    # rows that carry weights cannot be written (the file stores the unweighted kernel)
    pb.par.problem_weight = 2.0
    rows_w, _, _, _ = tfx.sensit_assemble_rows(pb.par, pb.grid, pb.data_xyz, pb.cw, pb.dw)
    with pytest.raises(tfx.TfxError, match="unweighted"):
        tfx.write_sensit_file(rows_w, d)


@pytest.mark.gpu
@pytest.mark.parametrize("nbproc_read", [1, 3])
def test_read_sensitivity_kernel_from_reference_style_files(tmp_path, oracle, nbproc_read):
    """Files written by the numpy restatement as 2 writer ranks -> read as 1 or 3 column slabs with problem and
    data weights applied in real(4) (sensitivity_gravmag.F90:837-843), against the oracle's slab rule."""
    from oracle import partition as orp
    pb = make_problem(nx=8, ny=7, nz=4, ndata=6, compression_type=2, rate=0.3, problem_type=2, nmodel_components=3)
    nmc, N = 3, pb.N
    d = str(tmp_path)
    shift = pb.par.param_shift
    pb.par.param_shift = 0
    full = _rows(*pb.oracle_matrix(oracle).arrays())               # unweighted (problem_weight = 1, data weight 1)
    pb.par.param_shift = shift
    nnz_col = np.zeros(N, dtype=np.int32)
    split = [range(1, 4), range(4, 7)]                              # even data split over 2 writer ranks
    for rank, rr in enumerate(split):
        recs = []
        for i in rr:
            c, v = full[i]
            for k in range(1, nmc + 1):
                sel = (c > (k - 1) * N) & (c <= k * N)
                recs.append((i, k, 1, c[sel] - (k - 1) * N, v[sel]))
                np.add.at(nnz_col, c[sel] - (k - 1) * N - 1, 1)
        osf.write_rank_file(d, 2, 2, rank, pb.ndata, N, recs)
    osf.write_meta(d, 2, pb.nx, pb.ny, pb.nz, pb.ndata, 2, 1, 2, 1e-4, nmc, 1, int(nnz_col.sum()), gfortran_style=True)
    osf.write_nnz(d, 2, nnz_col)
    assert np.array_equal(tfx.read_sensit_nnz(pb.par, d), nnz_col)
    nnz_at, nel_at = tfx.get_load_balancing_nelements(nnz_col, nbproc_read)
    pb.par.problem_weight = 0.75
    dw = np.linspace(0.5, 2.0, pb.ndata).reshape(pb.ndata, 1)
    total = 0
    for r in range(nbproc_read):
        M = tfx.read_sensitivity_kernel(pb.par, d, dw, 1, 2, nel_at, r, nbproc_read)
        got = _rows(*M.export())
        want = orp.column_slab(full, N, nmc, nel_at, r, 2)
        assert set(got) == set(want)
        for i in got:
            w = np.float32(0.75 * dw[i - 1, 0])
            assert np.array_equal(got[i][0], want[i][0])
            assert np.array_equal(got[i][1], (want[i][1] * w).astype(np.float32))      # bit-exact real(4) product
        assert M.get_number_elements() == nnz_at[r] and M.get_ncolumns() == 2 * nmc * nel_at[r]
        total += M.get_number_elements()
    assert total == nnz_col.sum()
    # header checks of the reader (:749-752, :772-789)
    bad = make_problem(nx=8, ny=7, nz=4, ndata=5, compression_type=2, rate=0.3, problem_type=2, nmodel_components=3)
    with pytest.raises(tfx.TfxError, match="does not match the Parfile"):
        tfx.read_sensitivity_kernel(bad.par, d, np.ones((5, 1)), 1, 2, [N])
