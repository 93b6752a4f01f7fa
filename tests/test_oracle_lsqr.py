"""Oracle vs the reference's LSQR / sparse-matrix known answers
(src/tests/tests_lsqr.f90, src/tests/tests_sparse_matrix.f90)."""
import numpy as np
import pytest

from tests.conftest import TOL, comparable


def _build(oracle, rows, ncols):
    m = oracle.SparseMatrix(len(rows), ncols, len(rows) * ncols)
    for r in rows:
        for i, v in enumerate(r):
            m.add(v, i + 1)
        m.new_row()
    m.finalize()
    return m


def test_lsqr_determined(oracle):
    # tests_lsqr.f90:71-122
    n = 1440
    m = _build(oracle, [[float(j)] * n for j in range(1, n + 1)], n)
    b = np.array([float(j * n) for j in range(1, n + 1)])
    x, hist, it = oracle.lsqr_solve(100, 1e-13, 0.0, m, b)
    assert all(comparable(v, 1.0, TOL) for v in x)


def test_lsqr_overdetermined_1(oracle):
    # tests_lsqr.f90:144-218
    nrows = 1000
    bb = (1.0, -3.0, 0.0)
    rows, rhs = [], []
    for i in range(1, nrows + 1):
        xi = float(i) / float(nrows)
        rows.append([xi ** 0, xi ** 1, xi ** 2])
        rhs.append(bb[0] + bb[1] * xi + bb[2] * xi ** 2)
    m = _build(oracle, rows, 3)
    x, hist, it = oracle.lsqr_solve(100, 1e-14, 0.0, m, np.array(rhs))
    assert comparable(x[0], bb[0], TOL)
    assert comparable(x[1], bb[1], TOL)
    assert abs(x[2]) < TOL


def test_lsqr_overdetermined_2(oracle):
    # tests_lsqr.f90:227-351 (Wunsch 5x3), single-precision matrix tolerance 1e-2
    a = [[1.2550, 1.6731, -1.3927], [0.4891, 0.0943, -0.7829], [-0.1755, 1.8612, 1.0972],
         [0.4189, 0.2469, -0.5990], [-0.2900, 0.7677, 0.8188]]
    b = np.array([0.3511, -1.6710, 6.838, -0.8843, 3.7018])
    m = _build(oracle, a, 3)
    x, hist, it = oracle.lsqr_solve(100, 1e-13, 0.0, m, b)
    # the reference compares against single-precision literals 157.611 etc.
    assert abs(x[0] - np.float32(157.611)) < 1e-2
    assert abs(x[1] + np.float32(38.0747)) < 1e-2
    assert abs(x[2] - np.float32(96.0291)) < 1e-2


def test_lsqr_underdetermined_1(oracle):
    # tests_lsqr.f90:366-447 -- min-norm solution (0, 1, 1), |x1| < 1e-15 absolute.
    m = _build(oracle, [[1.0, 1.0, 0.0], [2.0, 1.0, -1.0]], 3)
    x, hist, it = oracle.lsqr_solve(100, 1e-13, 0.0, m, np.array([1.0, 0.0]))
    assert abs(x[0]) < 1e-15
    assert comparable(x[1], 1.0, TOL)
    assert comparable(x[2], 1.0, TOL)


def test_lsqr_underdetermined_2(oracle):
    # tests_lsqr.f90:461-518
    m = _build(oracle, [[0.25] * 4], 4)
    x, hist, it = oracle.lsqr_solve(100, 1e-14, 0.0, m, np.array([1.0]))
    assert all(comparable(v, 1.0, TOL) for v in x)


def test_lsqr_underdetermined_3(oracle):
    # tests_lsqr.f90:532-624
    m = _build(oracle, [[1.0, 1.0, 1.0, 1.0], [1.0, -1.0, -1.0, 1.0]], 4)
    x, hist, it = oracle.lsqr_solve(100, 1e-14, 0.0, m, np.array([1.0, -1.0]))
    for got, want in zip(x, (0.0, 0.5, 0.5, 0.0)):
        assert comparable(got, want, TOL)


def test_normalize_columns(oracle):
    # tests_sparse_matrix.f90:39-113
    ncolumns, nrows = 10, 30
    A = np.zeros((nrows, ncolumns))
    counter = 0
    for j in range(nrows):
        for i in range(ncolumns):
            counter += 1
            A[j, i] = float(counter) if (i + 1) <= ncolumns // 2 else 0.0
    m = _build(oracle, A.tolist(), ncolumns)
    assert m.nel == nrows * (ncolumns // 2)          # zero values are not stored (:219)
    cn = m.normalize_columns()
    for i in range(ncolumns):
        assert comparable(cn[i], np.linalg.norm(A[:, i]), TOL)
        vi = np.zeros(ncolumns)
        vi[i] = 1.0
        col = m.mult_vector(vi)
        want = 1.0 if np.linalg.norm(A[:, i]) != 0 else 0.0
        assert comparable(np.linalg.norm(col), want, TOL)


def test_sensit_variant_equals_plain_when_no_constraints(oracle):
    # lsqr_solve_sensit (lsqr_solver2.F90:47) with an all-empty constraint matrix and no wavelet
    # must walk the same iterates as lsqr_solve (:321).
    rng = np.random.default_rng(7)
    nrows, ncols = 12, 20
    A = rng.standard_normal((nrows, ncols))
    m = _build(oracle, A.tolist(), ncols)
    Cm = oracle.SparseMatrix(5, ncols, 1, 5)
    Cm.add_empty_rows(5)
    Cm.finalize()
    b = rng.standard_normal(nrows)
    x1, h1, it1 = oracle.lsqr_solve(30, 1e-13, 0.0, m, b)
    x2, h2, it2 = oracle.lsqr_solve_sensit(30, 1e-13, 0.0, 0.0, m, Cm, np.concatenate([b, np.zeros(5)]),
                                           ncols // 2, ncols // 2, 1, 1, 1, 0, True)
    assert it1 == it2
    # v = -beta v + (S^T u) is associated differently in the two routines (:236 vs :414): residuals
    # agree to rounding until convergence noise takes over.
    big = h1 > 1e-9
    np.testing.assert_allclose(h1[big], h2[big], rtol=1e-6)
    np.testing.assert_allclose(x1, x2, rtol=1e-9, atol=1e-12)


def _random_sensit(oracle, rng, nrows, nx, ny, nz):
    N = nx * ny * nz
    A = rng.standard_normal((nrows, N))
    S = oracle.SparseMatrix(nrows, 2 * N, nrows * N)
    for i in range(nrows):
        S.add_row(A[i].astype(np.float32), np.arange(1, N + 1, dtype=np.int32))
        S.new_row()
    S.finalize()
    return S, A.astype(np.float32).astype(np.float64), N


@pytest.mark.parametrize("wtype", [1, 2])
def test_wavelet_in_loop_equals_wavelet_domain_solve(oracle, wtype):
    """lsqr_solve_sensit with WAVELET_DOMAIN = .false. (transforms inside the loop, lsqr_solver2.F90:200-207,228-235)
    solves min |S W x - b| for the physical-domain x; with an orthonormal W its iterates are the inverse transforms of
    the iterates of the wavelet-domain solve min |S y - b| and the residual histories coincide."""
    rng = np.random.default_rng(40 + wtype)
    nx, ny, nz, nrows = 6, 5, 4, 14
    S, A, N = _random_sensit(oracle, rng, nrows, nx, ny, nz)
    Cm = oracle.SparseMatrix(1, 2 * N, 1, 1)
    Cm.add_empty_rows(1)
    Cm.finalize()
    b = np.concatenate([rng.standard_normal(nrows), [0.0]])
    y, hy, ity = oracle.lsqr_solve_sensit(12, 1e-13, 0.0, 0.0, S, Cm, b, N, nx, ny, nz, 1, wtype, True)
    x, hx, itx = oracle.lsqr_solve_sensit(12, 1e-13, 0.0, 0.0, S, Cm, b, N, nx, ny, nz, 1, wtype, False)
    assert itx == ity
    np.testing.assert_allclose(hx, hy, rtol=1e-9)
    np.testing.assert_allclose(x[:N], oracle.inverse_wavelet(y[:N].copy(), nx, ny, nz, wtype), rtol=1e-8, atol=1e-11)


def test_target_misfit_exits_early(oracle):
    """The misfit exit (lsqr_solver2.F90:168-189): RMSE of S x against the incoming right-hand side, checked BEFORE the
    iteration body -- the run stops at the first iteration whose incoming x is good enough."""
    rng = np.random.default_rng(50)
    nx, ny, nz, nrows = 5, 4, 3, 10
    S, A, N = _random_sensit(oracle, rng, nrows, nx, ny, nz)
    Cm = oracle.SparseMatrix(1, 2 * N, 1, 1)
    Cm.add_empty_rows(1)
    Cm.finalize()
    b = np.concatenate([rng.standard_normal(nrows), [0.0]])
    x_full, h_full, it_full = oracle.lsqr_solve_sensit(40, 1e-13, 0.0, 0.0, S, Cm, b, N, nx, ny, nz, 1, 0, True)
    rmse = lambda x: np.sqrt(np.sum((A @ x[:N] - b[:nrows]) ** 2) / nrows)
    target = 10.0 * rmse(x_full) + 1e-3
    x, h, it = oracle.lsqr_solve_sensit(40, 1e-13, 0.0, target, S, Cm, b, N, nx, ny, nz, 1, 0, True)
    assert 0 < it < it_full
    assert rmse(x) <= target
    # the iterations before the exit are those of the run without a target
    np.testing.assert_allclose(h, h_full[:it], rtol=1e-12)


def test_soft_thresholding_shrinks_the_iterate(oracle):
    """gamma > 0 applies ISTA's proximal step to x after every update (lsqr_solver2.F90:272-275,478-494): entries are
    pulled towards zero by gamma and small ones vanish; gamma = 0 is the plain solver."""
    rng = np.random.default_rng(60)
    nx, ny, nz, nrows = 5, 4, 3, 10
    S, A, N = _random_sensit(oracle, rng, nrows, nx, ny, nz)
    Cm = oracle.SparseMatrix(1, 2 * N, 1, 1)
    Cm.add_empty_rows(1)
    Cm.finalize()
    b = np.concatenate([rng.standard_normal(nrows), [0.0]])
    x0, _, _ = oracle.lsqr_solve_sensit(1, 1e-13, 0.0, 0.0, S, Cm, b, N, nx, ny, nz, 1, 0, True)
    gamma = 0.5 * np.abs(x0[:N]).max()
    xg, _, _ = oracle.lsqr_solve_sensit(1, 1e-13, gamma, 0.0, S, Cm, b, N, nx, ny, nz, 1, 0, True)
    want = np.sign(x0) * np.maximum(np.abs(x0) - gamma, 0.0)             # one iteration: the prox of the plain iterate
    np.testing.assert_allclose(xg, want, rtol=1e-13, atol=1e-300)
    assert np.count_nonzero(xg[:N]) < np.count_nonzero(x0[:N])
