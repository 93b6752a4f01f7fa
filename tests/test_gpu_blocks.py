"""Row-blocked sensitivity matrices (option sensit_row_blocks): the kernel assembled batch of stations after batch of
stations, each batch an independent row block with its own product layouts -- bounded build memory for the big
compressed configs. Products, part_mult_vector, calculate_data and the LSQR solve must match the matrix assembled in one
piece from the same rows."""
import copy

import numpy as np
import pytest

import tomofastx_b200 as tfx
from tests.synth import make_problem

pytestmark = pytest.mark.gpu


@pytest.fixture
def blocks_on():
    tfx.set_option("sensit_row_blocks", 1)
    yield
    tfx.set_option("sensit_row_blocks", 0)
    tfx.set_option("t16_min_nnz", 1 << 22)


def _batch_par(pb, n):
    par = copy.copy(pb.par)
    par.ndata = n
    return par


def _assemble(pb, batches, blocked, t16):
    tfx.set_option("sensit_row_blocks", 1 if blocked else 0)
    tfx.set_option("t16_min_nnz", 0 if t16 else 1 << 22)
    N, ndc = pb.N, pb.ndc
    nel_at = np.array([N], dtype=np.int32)
    S = tfx.SparseMatrix(pb.ndata * ndc, pb.ncolumns, pb.ndata * ndc * N)
    x, y, z = pb.data_xyz
    d0 = 0
    for n in batches:
        sl = slice(d0, d0 + n)
        rows, _, _, _ = tfx.sensit_assemble_rows(_batch_par(pb, n), pb.grid, (x[sl], y[sl], z[sl]), pb.cw, pb.dw[sl])
        tfx.sensit_repartition_into(S, rows, 1, nel_at)
        d0 += n
    S.finalize()
    return S


@pytest.mark.parametrize("t16", [False, True])
@pytest.mark.parametrize("case", ["grav_haar", "mag3_d4"])
def test_blocked_matrix_matches_single_piece(oracle, blocks_on, case, t16):
    if case == "grav_haar":
        pb = make_problem(nx=12, ny=10, nz=6, ndata=23, compression_type=1, rate=0.25)
    else:
        pb = make_problem(nx=8, ny=7, nz=4, ndata=11, compression_type=2, rate=0.3, problem_type=2, nmodel_components=3)
    batches = [pb.ndata // 3, pb.ndata // 3 + 2, pb.ndata - 2 * (pb.ndata // 3) - 2]
    S1 = _assemble(pb, [pb.ndata], False, t16)
    Sb = _assemble(pb, batches, True, t16)
    assert Sb.get_number_elements() == S1.get_number_elements()
    assert Sb.storage_kind() == (2 if t16 else 0)
    rng = np.random.default_rng(3)
    x = rng.standard_normal(pb.ncolumns); u = rng.standard_normal(pb.ndata * pb.ndc)
    # forward: every row lives in exactly one block with the same entries in the same order
    assert np.allclose(Sb.mult_vector(x), S1.mult_vector(x), rtol=1e-13, atol=1e-300)
    # transposed: the sum over rows is split by block
    t1 = S1.trans_mult_vector(u)
    assert np.allclose(Sb.trans_mult_vector(u), t1, rtol=1e-12, atol=1e-13 * np.abs(t1).max())
    acc = rng.standard_normal(pb.ncolumns)
    a1, ab = acc.copy(), acc.copy()
    S1.add_trans_mult_vector(u, a1); Sb.add_trans_mult_vector(u, ab)
    assert np.allclose(ab, a1, rtol=1e-12, atol=1e-13 * np.abs(a1).max())
    # a window of rows crossing a block boundary
    n0, n = batches[0] * pb.ndc - 1, 4
    xm = rng.standard_normal(pb.N * pb.par.nmodel_components)
    assert np.allclose(Sb.part_mult_vector(xm, n, n0, 0), S1.part_mult_vector(xm, n, n0, 0), rtol=1e-13, atol=1e-300)
    d1 = tfx.calculate_data(S1, pb.m_true, pb.ndata, pb.ndc, 1.0, pb.cw, pb.dw, pb.par.compression_type, pb.nx, pb.ny, pb.nz)
    db = tfx.calculate_data(Sb, pb.m_true, pb.ndata, pb.ndc, 1.0, pb.cw, pb.dw, pb.par.compression_type, pb.nx, pb.ny, pb.nz)
    assert np.allclose(db, d1, rtol=1e-13, atol=1e-300)
    # LSQR on both
    b = S1.mult_vector(rng.standard_normal(pb.ncolumns))
    hist = []
    for S in (S1, Sb):
        uu = b.copy(); xx = np.zeros(pb.ncolumns)
        tfx.lsqr_solve(len(uu), pb.ncolumns, 25, 1e-13, 0.0, S, uu, xx)
        h, it, _ = tfx.last_history()
        hist.append((h, it, xx))
    assert hist[0][1] == hist[1][1]
    n = min(10, len(hist[0][0]))
    assert np.allclose(hist[1][0][:n], hist[0][0][:n], rtol=1e-9)
    with pytest.raises(tfx.TfxError, match="row-blocked"):
        Sb.export()


def test_blocked_matrix_row_count_is_checked(blocks_on):
    pb = make_problem(nx=6, ny=5, nz=4, ndata=8, compression_type=1, rate=0.3)
    tfx.set_option("sensit_row_blocks", 1)
    S = tfx.SparseMatrix(pb.ndata, pb.ncolumns, pb.ndata * pb.N)
    x, y, z = pb.data_xyz
    rows, _, _, _ = tfx.sensit_assemble_rows(_batch_par(pb, 5), pb.grid, (x[:5], y[:5], z[:5]), pb.cw, pb.dw[:5])
    tfx.sensit_repartition_into(S, rows, 1, np.array([pb.N], dtype=np.int32))
    with pytest.raises(tfx.TfxError, match="total number of rows"):
        S.finalize()


@pytest.mark.parametrize("t16", [False, True])
def test_joint_blocked_matrix_part_products(blocks_on, t16):
    """Joint matrix_sensit built from row blocks: the blocks of problem 1 hold columns 1..N, those of problem 2 columns
    N+1..2N (sensitivity_gravmag.F90:685-686). part_mult_vector / calculate_data of one problem (model.F90:288: line_start,
    param_shift of that problem, x of nelements) must only touch that problem's blocks -- the other problem's columns lie
    outside x (out-of-bounds read through x[col - param_shift])."""
    g = make_problem(nx=10, ny=8, nz=5, ndata=9, compression_type=1, rate=0.25, problem_type=1)
    m = make_problem(nx=10, ny=8, nz=5, ndata=7, compression_type=1, rate=0.25, problem_type=2, nmodel_components=1)
    N = g.N
    m.par.param_shift = N
    m.par.ncolumns = g.par.ncolumns = 2 * N
    nel_at = np.array([N], dtype=np.int32)
    mats = []
    for blocked in (False, True):
        tfx.set_option("sensit_row_blocks", 1 if blocked else 0)
        tfx.set_option("t16_min_nnz", 0 if t16 else 1 << 22)
        S = tfx.SparseMatrix(g.ndata + m.ndata, 2 * N, 2 * int(0.25 * N) * (g.ndata + m.ndata))
        for slot, pb, batches in ((1, g, [4, 5]), (2, m, [3, 4])):
            x, y, z = pb.data_xyz
            d0 = 0
            for n in (batches if blocked else [pb.ndata]):
                sl = slice(d0, d0 + n)
                rows, _, _, _ = tfx.sensit_assemble_rows(_batch_par(pb, n), pb.grid, (x[sl], y[sl], z[sl]), pb.cw, pb.dw[sl])
                tfx.sensit_repartition_into(S, rows, slot, nel_at)
                d0 += n
        S.finalize()
        mats.append(S)
    S1, Sb = mats
    assert Sb.get_number_elements() == S1.get_number_elements()
    rng = np.random.default_rng(11)
    for line_start, nd, shift in ((1, g.ndata, 0), (g.ndata + 1, m.ndata, N), (3, 4, 0), (g.ndata + 2, 3, N)):
        xm = rng.standard_normal(N)
        want = S1.part_mult_vector(xm, nd, line_start, shift)
        got = Sb.part_mult_vector(xm, nd, line_start, shift)
        assert np.all(np.isfinite(got)) and np.abs(want).max() > 0
        assert np.allclose(got, want, rtol=1e-13, atol=1e-300)
    for pb, line_start, shift in ((g, 1, 0), (m, g.ndata + 1, N)):
        ct = pb.par.compression_type
        d1 = tfx.calculate_data(S1, pb.m_true, pb.ndata, 1, 1.0, pb.cw, pb.dw, ct, pb.nx, pb.ny, pb.nz, line_start, shift)
        db = tfx.calculate_data(Sb, pb.m_true, pb.ndata, 1, 1.0, pb.cw, pb.dw, ct, pb.nx, pb.ny, pb.nz, line_start, shift)
        assert np.abs(d1).max() > 0 and np.allclose(db, d1, rtol=1e-13, atol=1e-300)
    if t16:
        # a window that spans both problems cannot be served from one nelements-sized x: loud error, no wild read
        with pytest.raises(tfx.TfxError, match="outside"):
            Sb.part_mult_vector(rng.standard_normal(N), g.ndata + 2, 1, 0)


def test_dense_row_blocks_match_single_dense_block(oracle):
    """Uncompressed gravity kernels with more data rows than the register-resident sweep holds (kDenseMaxRows = 10 240)
    are stored as several dense row blocks (4 B per entry). Option dense_block_rows lowers the block size so that a small
    problem exercises the path: products, part_mult_vector, calculate_data and the LSQR solve (split path over the blocks)
    against the single block (fused path) and the oracle."""
    pb = make_problem(nx=10, ny=9, nz=5, ndata=50, compression_type=0)
    S1, nnz1, _, tot1 = tfx.calculate_sensit(pb.par, pb.grid, pb.data_xyz, pb.cw, pb.dw)
    try:
        tfx.set_option("dense_block_rows", 16)
        Sb, nnzb, _, totb = tfx.calculate_sensit(pb.par, pb.grid, pb.data_xyz, pb.cw, pb.dw)
    finally:
        tfx.set_option("dense_block_rows", 0)
    assert S1.storage_kind() == 1 and Sb.storage_kind() == 1
    assert tot1 == totb == pb.ndata * pb.N and np.array_equal(nnz1, nnzb)
    assert Sb.device_bytes() >= 4 * pb.ndata * pb.N
    So = pb.oracle_matrix(oracle)
    rng = np.random.default_rng(5)
    x = rng.standard_normal(pb.ncolumns); u = rng.standard_normal(pb.ndata)
    want = So.mult_vector(x)
    assert np.array_equal(Sb.mult_vector(x), S1.mult_vector(x))          # same columns per CTA, same order per row
    assert np.allclose(Sb.mult_vector(x), want, rtol=1e-6, atol=1e-7 * np.abs(want).max())   # f32 values vs the oracle's
    t1 = S1.trans_mult_vector(u)
    assert np.allclose(Sb.trans_mult_vector(u), t1, rtol=1e-12, atol=1e-13 * np.abs(t1).max())
    acc = rng.standard_normal(pb.ncolumns)
    a1, ab = acc.copy(), acc.copy()
    S1.add_trans_mult_vector(u, a1); Sb.add_trans_mult_vector(u, ab)
    assert np.allclose(ab, a1, rtol=1e-12, atol=1e-13 * np.abs(a1).max())
    accd = rng.standard_normal(pb.ndata)
    d1, db = accd.copy(), accd.copy()
    S1.add_mult_vector(x, d1); Sb.add_mult_vector(x, db)
    assert np.allclose(db, d1, rtol=1e-13, atol=1e-300)
    xm = rng.standard_normal(pb.N)
    assert np.allclose(Sb.part_mult_vector(xm, 9, 12, 0), So.part_mult_vector(xm, 9, 12, 0), rtol=1e-6,
                       atol=1e-7 * np.abs(want).max())
    d1 = tfx.calculate_data(S1, pb.m_true, pb.ndata, pb.ndc, 1.0, pb.cw, pb.dw, 0, pb.nx, pb.ny, pb.nz)
    db = tfx.calculate_data(Sb, pb.m_true, pb.ndata, pb.ndc, 1.0, pb.cw, pb.dw, 0, pb.nx, pb.ny, pb.nz)
    assert np.allclose(db, d1, rtol=1e-13, atol=1e-300)
    b = S1.mult_vector(rng.standard_normal(pb.ncolumns))
    out = []
    for S in (S1, Sb):
        uu = b.copy(); xx = np.zeros(pb.ncolumns)
        tfx.lsqr_solve(len(uu), pb.ncolumns, 25, 1e-13, 0.0, S, uu, xx)
        h, it, fused = tfx.last_history()
        out.append((h, it, xx, fused))
    assert out[0][3] == 1 and out[1][3] == 0                              # fused sweep vs split path over the blocks
    assert out[0][1] == out[1][1]
    n = min(10, len(out[0][0]))
    assert np.allclose(out[1][0][:n], out[0][0][:n], rtol=1e-9)
    # (25 LSQR iterations amplify the last-bit differences of the two summation orders)
    assert np.allclose(out[1][2], out[0][2], rtol=1e-4, atol=1e-5 * np.abs(out[0][2]).max())
