"""T16 tiled layouts (csrc/t16.cu): forward / transposed products and LSQR on the 16-bit tiled copies of the
matrix, forced on small matrices with option t16_min_nnz = 0 and several tile sizes (one tile -> DIRECT,
many tiles -> TILES, many tiles with a long output side -> DIRECT tile after tile). Checked against the
oracle like tests/test_gpu_sparse.py (tolerance 1e-12 relative to the absolute-value row sums)."""
import numpy as np
import pytest

import tomofastx_b200 as tfx
from tests.test_gpu_sparse import check_products, random_matrix
from tests.test_gpu_lsqr import _sensit_case, assert_history

pytestmark = pytest.mark.gpu


@pytest.fixture
def force_t16():
    tfx.set_option("t16_min_nnz", 0)
    tfx.set_option("dense_detect", 0)
    yield
    tfx.set_option("t16_min_nnz", 1 << 22)
    tfx.set_option("t16_tile", 0)
    tfx.set_option("dense_detect", 1)


CASES = {
    # name: (nl, ncol, row length fn, empty-row probability)
    "compressed_like": (300, 40000, lambda i: 2000, 0.0),
    "short_rows": (5000, 3000, lambda i: 1 + i % 6, 0.0),
    "skewed": (300, 40000, lambda i: 30000 if i % 50 == 0 else 1 + i % 40, 0.0),
    "with_empty_rows": (400, 900, lambda i: 50, 0.4),
    "very_long_rows": (7, 60000, lambda i: 50000, 0.0),
    "odd_lengths": (33, 1000, lambda i: 1 + 2 * (i % 17), 0.0),
    "single_entry": (3, 3, lambda i: 1, 0.0),
    "many_rows": (20000, 500, lambda i: 20, 0.1),          # transposed layout needs two row tiles
}


@pytest.mark.parametrize("tile", [0, 64, 1024])
@pytest.mark.parametrize("case", sorted(CASES))
def test_products_vs_oracle(oracle, force_t16, case, tile):
    nl, ncol, fn, ep = CASES[case]
    rng = np.random.default_rng(sum(ord(ch) for ch in case) + tile)
    tfx.set_option("t16_tile", tile)
    mo, mg, rows = random_matrix(oracle, rng, nl, ncol, fn, ep)
    assert mg.storage_kind() == 2
    check_products(mo, mg, rng, nl, ncol)


def test_unsorted_rows_stay_on_generic_kernels(oracle, force_t16):
    mg = tfx.SparseMatrix(2, 5, 6)
    mg.add(1.0, 3); mg.add(2.0, 1); mg.new_row()          # columns not ascending
    mg.add(1.0, 2); mg.new_row()
    mg.finalize()
    assert mg.storage_kind() == 0
    assert np.allclose(mg.mult_vector(np.arange(1.0, 6.0)), [5.0, 2.0])


def test_part_mult_vector_t16(oracle, force_t16):
    rng = np.random.default_rng(12)
    nel, nd = 700, 40
    mo = oracle.SparseMatrix(nd, 2 * nel, nd * 100)
    mg = tfx.SparseMatrix(nd, 2 * nel, nd * 100)
    for i in range(nd):
        cols = np.sort(rng.choice(nel, size=100, replace=False)).astype(np.int32) + 1 + nel   # second problem
        vals = rng.standard_normal(100).astype(np.float32)
        for m in (mo, mg):
            m.add_row(vals, cols); m.new_row()
    mo.finalize(); mg.finalize()
    assert mg.storage_kind() == 2
    x = rng.standard_normal(nel)
    want = mo.part_mult_vector(x, 10, 5, nel)
    got = mg.part_mult_vector(x, 10, 5, nel)
    assert np.allclose(got, want, rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("tile", [0, 32])
def test_lsqr_solve_sensit_on_t16(oracle, force_t16, tile):
    tfx.set_option("t16_tile", tile)
    rng = np.random.default_rng(123)
    nx, ny, nz, ndata = 6, 5, 4, 24
    So, Sg, Co, Cg, b, N, ncol = _sensit_case(oracle, rng, nx, ny, nz, ndata, 0.3, "damping", False)
    tfx.set_option("dense_detect", 0)
    assert Sg.storage_kind() == 2
    niter = 400
    xr, hr, itr = oracle.lsqr_solve_sensit(niter, 1e-13, 0.0, 0.0, So, Co, b, N, nx, ny, nz, 1, 1, True)
    u = b.copy(); x = np.zeros(ncol)
    tfx.lsqr_solve_sensit(len(b), ncol, niter, 1e-13, 0.0, 0.0, Sg, Cg, u, x, [1, 0], N, nx, ny, nz, 1, 1, True)
    h, it, fused = tfx.last_history()
    assert not fused
    assert_history(h, hr, first=8)
    assert abs(h[-1] - hr[-1]) <= 1e-6 * hr[-1]
    assert np.allclose(x, xr, rtol=1e-6, atol=1e-8 * np.abs(xr).max())


def test_compressed_assembly_builds_t16(oracle, force_t16):
    from tests.synth import make_problem
    pb = make_problem(nx=16, ny=12, nz=6, ndata=30, compression_type=1, rate=0.2)
    S_gpu, nnz_col, cerr, tot = tfx.calculate_sensit(pb.par, pb.grid, pb.data_xyz, pb.cw, pb.dw)
    assert S_gpu.storage_kind() == 2
    S_orc = pb.oracle_matrix(oracle)
    rng = np.random.default_rng(3)
    x = rng.standard_normal(pb.ncolumns); y = rng.standard_normal(pb.ndata)
    w1, g1 = S_orc.mult_vector(x), S_gpu.mult_vector(x)
    assert np.allclose(g1, w1, rtol=1e-11, atol=1e-13 * np.abs(w1).max())
    w2, g2 = S_orc.trans_mult_vector(y), S_gpu.trans_mult_vector(y)
    assert np.allclose(g2, w2, rtol=1e-11, atol=1e-13 * np.abs(w2).max())


def test_direct_mode_tile_after_tile(oracle, force_t16):
    # transposed layout: 40 gathered rows in tiles of 16 (3 tiles), 300k outputs (> 2^18) -> DIRECT per tile
    tfx.set_option("t16_tile", 16)
    rng = np.random.default_rng(77)
    nl, ncol = 40, 300000
    mo, mg, rows = random_matrix(oracle, rng, nl, ncol, lambda i: 5000, 0.0)
    assert mg.storage_kind() == 2
    check_products(mo, mg, rng, nl, ncol)


def test_adjoint_identity_real_compressed_matrix(force_t16):
    # size-independent property on a real Haar-compressed gravity matrix: <S x, u> == <x, S^T u>; the two
    # products run on two different device copies (F and T layouts), so this ties them together.
    from tests.synth import make_problem
    tfx.set_option("t16_tile", 0)
    pb = make_problem(nx=64, ny=48, nz=16, ndata=300, compression_type=1, rate=0.05)
    S, nnz_col, cerr, tot = tfx.calculate_sensit(pb.par, pb.grid, pb.data_xyz, pb.cw, pb.dw)
    assert S.storage_kind() == 2 and tot <= 300 * int(0.05 * pb.N)
    rng = np.random.default_rng(8)
    x = rng.uniform(-1, 1, pb.ncolumns); u = rng.uniform(-1, 1, pb.ndata)
    lhs = float(np.dot(S.mult_vector(x), u)); rhs = float(np.dot(x, S.trans_mult_vector(u)))
    assert abs(lhs - rhs) <= 1e-11 * max(abs(lhs), abs(rhs))
    # linearity: S(a x1 + b x2) == a S x1 + b S x2
    x2 = rng.uniform(-1, 1, pb.ncolumns)
    y = S.mult_vector(2.0 * x - 3.0 * x2)
    y_lin = 2.0 * S.mult_vector(x) - 3.0 * S.mult_vector(x2)
    assert np.allclose(y, y_lin, rtol=1e-10, atol=1e-12 * np.abs(y_lin).max())
    # the CSR copies can be dropped; products keep working, export does not
    b0 = S.device_bytes(); S.drop_csr()
    assert S.device_bytes() < b0
    assert np.allclose(S.mult_vector(x), (y_lin + 3.0 * S.mult_vector(x2)) / 2.0, rtol=1e-9, atol=1e-12 * np.abs(y_lin).max())
