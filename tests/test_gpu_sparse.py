"""GPU sparse products (csrc/csr.cu) through the C ABI vs the oracle, mirroring sparse_matrix.f90 and
tests_sparse_matrix.f90. Products are f32 x f64 -> f64 sums in a different (tree) order than the
reference's sequential loop: tolerance 1e-12 relative to the row's absolute-value sum."""
import numpy as np
import pytest

import tomofastx_b200 as tfx
from tests.conftest import TOL, comparable

pytestmark = pytest.mark.gpu


def random_matrix(orc, rng, nl, ncol, row_len_fn, empty_prob=0.0):
    rows = []
    nnz = 0
    for i in range(nl):
        if rng.random() < empty_prob:
            rows.append((np.zeros(0, np.int32), np.zeros(0, np.float32)))
            continue
        L = min(ncol, max(1, int(row_len_fn(i))))
        cols = np.sort(rng.choice(ncol, size=L, replace=False)).astype(np.int32) + 1
        vals = rng.standard_normal(L).astype(np.float32)
        vals[vals == 0] = 1.0
        rows.append((cols, vals))
        nnz += L
    nl_empty = sum(1 for c, _ in rows if len(c) == 0)
    mo = orc.SparseMatrix(nl, ncol, max(nnz, 1), nl_empty)
    mg = tfx.SparseMatrix(nl, ncol, max(nnz, 1), 0, nl_empty)
    for cols, vals in rows:
        if len(cols):
            mo.add_row(vals, cols)
            mg.add_row(vals, cols)
            mo.new_row()
            mg.new_row()
        else:                         # the reference adds empty rows with add_empty_rows (damping.F90:158,179)
            mo.add_empty_rows(1)
            mg.add_empty_rows(1)
    mo.finalize()
    mg.finalize()
    return mo, mg, rows


def check_products(mo, mg, rng, nl, ncol):
    x = rng.standard_normal(ncol)
    y = rng.standard_normal(nl)
    absrow = np.zeros(nl)
    sa, ija, ijl, rowptr = mo.arrays()
    for i in range(len(rowptr)):
        seg = slice(ijl[i] - 1, ijl[i + 1] - 1)
        absrow[rowptr[i] - 1] = np.sum(np.abs(sa[seg].astype(np.float64) * x[ija[seg] - 1]))
    want = mo.mult_vector(x)
    got = mg.mult_vector(x)
    assert np.all(np.abs(got - want) <= 1e-12 * (absrow + 1e-300) + 1e-300)
    b0 = rng.standard_normal(nl)
    want2 = b0.copy(); mo.add_mult_vector(x, want2)
    got2 = mg.add_mult_vector(x, b0.copy())
    assert np.allclose(got2, want2, rtol=1e-12, atol=1e-12 * (np.abs(b0).max() + absrow.max()))
    wt = mo.trans_mult_vector(y)
    gt = mg.trans_mult_vector(y)
    scale = np.abs(wt).max() + 1.0
    assert np.allclose(gt, wt, rtol=1e-11, atol=1e-12 * scale * max(1, nl ** 0.5))
    c0 = rng.standard_normal(ncol)
    wt2 = c0.copy(); mo.add_trans_mult_vector(y, wt2)
    gt2 = mg.add_trans_mult_vector(y, c0.copy())
    assert np.allclose(gt2, wt2, rtol=1e-11, atol=1e-12 * scale * max(1, nl ** 0.5))


@pytest.mark.parametrize("case", ["uniform_long", "short_rows", "skewed", "with_empty_rows", "very_long_rows",
                                  "single_row", "single_entry"])
def test_products_vs_oracle(oracle, case):
    rng = np.random.default_rng(sum(ord(ch) for ch in case))
    if case == "uniform_long":
        nl, ncol, fn, ep = 64, 20000, (lambda i: 3000), 0.0
    elif case == "short_rows":
        nl, ncol, fn, ep = 5000, 3000, (lambda i: 1 + i % 6), 0.0
    elif case == "skewed":
        nl, ncol, fn, ep = 300, 40000, (lambda i: 30000 if i % 50 == 0 else 1 + i % 40), 0.0
    elif case == "with_empty_rows":
        nl, ncol, fn, ep = 400, 900, (lambda i: 50), 0.4
    elif case == "very_long_rows":
        nl, ncol, fn, ep = 7, 60000, (lambda i: 50000), 0.0      # > kItemLen: several work items per row
    elif case == "single_row":
        nl, ncol, fn, ep = 1, 10, (lambda i: 4), 0.0
    else:
        nl, ncol, fn, ep = 3, 3, (lambda i: 1), 0.0
    tfx.set_option("dense_detect", 0)
    mo, mg, rows = random_matrix(oracle, rng, nl, ncol, fn, ep)
    assert mg.get_number_elements() == mo.nel
    check_products(mo, mg, rng, nl, ncol)
    tfx.set_option("dense_detect", 1)


def test_all_rows_empty(oracle):
    mg = tfx.SparseMatrix(5, 7, 1, 0, 5)
    mg.add_empty_rows(5)
    mg.finalize()
    assert np.array_equal(mg.mult_vector(np.ones(7)), np.zeros(5))
    assert np.array_equal(mg.trans_mult_vector(np.ones(5)), np.zeros(7))


def test_part_mult_vector(oracle):
    # joint-matrix layout: two problems stacked in rows, columns shifted by param_shift (model.F90:288)
    rng = np.random.default_rng(11)
    nel, nd1, nd2 = 500, 20, 30
    ncol = 2 * nel
    mo = oracle.SparseMatrix(nd1 + nd2, ncol, (nd1 + nd2) * 100)
    mg = tfx.SparseMatrix(nd1 + nd2, ncol, (nd1 + nd2) * 100)
    for i in range(nd1 + nd2):
        shift = 0 if i < nd1 else nel
        cols = np.sort(rng.choice(nel, size=100, replace=False)).astype(np.int32) + 1 + shift
        vals = rng.standard_normal(100).astype(np.float32)
        for m in (mo, mg):
            m.add_row(vals, cols)
            m.new_row()
    mo.finalize(); mg.finalize()
    x = rng.standard_normal(nel)
    for (ls, nd, ps) in [(1, nd1, 0), (nd1 + 1, nd2, nel)]:
        want = mo.part_mult_vector(x, nd, ls, ps)
        got = mg.part_mult_vector(x, nd, ls, ps)
        assert np.allclose(got, want, rtol=1e-12, atol=1e-12)
    with pytest.raises(tfx.TfxError, match="Wrong line index"):
        mg.part_mult_vector(x, nd2 + 1, nd1 + 1, nel)


def test_normalize_columns_style_check(oracle):
    # tests_sparse_matrix.f90:39-113 without normalize_columns (test-only routine): CSR build with zero
    # entries dropped and mult_vector against the dense matrix.
    ncolumns, nrows = 10, 30
    A = np.zeros((nrows, ncolumns))
    counter = 0
    m = tfx.SparseMatrix(nrows, ncolumns, ncolumns * nrows)
    for j in range(nrows):
        for i in range(ncolumns):
            counter += 1
            A[j, i] = float(counter) if (i + 1) <= ncolumns // 2 else 0.0
            m.add(A[j, i], i + 1)
        m.new_row()
    m.finalize()
    assert m.get_number_elements() == nrows * (ncolumns // 2)
    for i in range(ncolumns):
        vi = np.zeros(ncolumns); vi[i] = 1.0
        col = m.mult_vector(vi)
        assert np.allclose(col, A[:, i], rtol=1e-7)


def test_finalize_errors():
    m = tfx.SparseMatrix(2, 3, 4)
    m.add(1.0, 1); m.new_row()
    with pytest.raises(tfx.TfxError, match="total number of rows"):
        m.finalize()
    m2 = tfx.SparseMatrix(1, 3, 4)
    m2.add(1.0, 7); m2.new_row()
    with pytest.raises(tfx.TfxError, match="column-index validation"):
        m2.finalize()


def test_device_pointer_vectors(oracle):
    import torch
    rng = np.random.default_rng(2)
    tfx.set_option("dense_detect", 0)
    mo, mg, _ = random_matrix(oracle, rng, 50, 800, lambda i: 200)
    tfx.set_option("dense_detect", 1)
    x = rng.standard_normal(800)
    xd = torch.from_numpy(x).cuda()
    bd = torch.zeros(50, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()              # libtfx runs on its own non-blocking stream: order torch's work first
    mg.mult_vector(xd, bd)
    assert np.allclose(bd.cpu().numpy(), mo.mult_vector(x), rtol=1e-12, atol=1e-12)
