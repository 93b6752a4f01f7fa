"""LSQR on the device (csrc/lsqr.cu) through the C ABI: the reference's own unit tests
(src/tests/tests_lsqr.f90) plus lsqr_solve_sensit cases the reference never tests, checked against the
oracle's per-iteration residual history (relative tolerance 1e-6, the north-star parity bar)."""
import numpy as np
import pytest

import tomofastx_b200 as tfx
from tests.conftest import TOL, comparable

pytestmark = pytest.mark.gpu


def build_pair(orc, rows, ncols, dense_detect=1):
    tfx.set_option("dense_detect", dense_detect)
    mo = orc.SparseMatrix(len(rows), ncols, len(rows) * ncols)
    mg = tfx.SparseMatrix(len(rows), ncols, len(rows) * ncols)
    for r in rows:
        for i, v in enumerate(r):
            mo.add(v, i + 1)
            mg.add(v, i + 1)
        mo.new_row(); mg.new_row()
    mo.finalize(); mg.finalize()
    tfx.set_option("dense_detect", 1)
    return mo, mg


def gpu_solve(mg, b, niter, rmin, gamma=0.0):
    u = np.array(b, dtype=np.float64)
    x = np.zeros(mg.get_ncolumns())
    tfx.lsqr_solve(len(u), len(x), niter, rmin, gamma, mg, u, x)
    h, it, fused = tfx.last_history()
    return x, h, it, fused


def assert_history(h, h_ref, floor=1e-9, rtol=1e-6, first=None):
    """Residual histories. LSQR's mid-phase iterates are chaotic under last-bit perturbations (see
    tests/test_oracle_mansf.py::test_residual_history_is_chaotic_under_rounding), so the fast kernels
    (tree-order sums) are compared on the first `first` iterations only; the strict_order mode, which
    reproduces the reference's summation order, is compared on all of them."""
    n = min(len(h), len(h_ref))
    if first is not None:
        n = min(n, first)
    big = h_ref[:n] > floor
    assert np.allclose(h[:n][big], h_ref[:n][big], rtol=rtol), (h, h_ref)


@pytest.fixture(params=["fast", "strict"])
def order(request):
    tfx.set_option("strict_order", 1 if request.param == "strict" else 0)
    yield request.param
    tfx.set_option("strict_order", 0)


@pytest.mark.parametrize("dense", [0, 1])
def test_lsqr_determined(oracle, dense):
    n = 1440
    mo, mg = build_pair(oracle, [[float(j)] * n for j in range(1, n + 1)], n, dense)
    assert mg.storage_kind() == dense
    b = np.array([float(j * n) for j in range(1, n + 1)])
    x, h, it, fused = gpu_solve(mg, b, 100, 1e-13)
    assert fused == bool(dense)
    assert all(comparable(v, 1.0, TOL) for v in x)
    xr, hr, itr = oracle.lsqr_solve(100, 1e-13, 0.0, mo, b)
    assert_history(h, hr)


@pytest.mark.parametrize("dense", [0, 1])
def test_lsqr_overdetermined_1(oracle, dense):
    nrows = 1000
    bb = (1.0, -3.0, 0.0)
    rows, rhs = [], []
    for i in range(1, nrows + 1):
        xi = float(i) / float(nrows)
        rows.append([xi ** 0, xi ** 1, xi ** 2])
        rhs.append(bb[0] + bb[1] * xi + bb[2] * xi ** 2)
    mo, mg = build_pair(oracle, rows, 3, dense)
    x, h, it, fused = gpu_solve(mg, rhs, 100, 1e-14)
    assert comparable(x[0], bb[0], TOL) and comparable(x[1], bb[1], TOL) and abs(x[2]) < TOL
    xr, hr, itr = oracle.lsqr_solve(100, 1e-14, 0.0, mo, np.array(rhs))
    assert_history(h, hr)


@pytest.mark.parametrize("dense", [0, 1])
def test_lsqr_overdetermined_2(oracle, dense):
    a = [[1.2550, 1.6731, -1.3927], [0.4891, 0.0943, -0.7829], [-0.1755, 1.8612, 1.0972],
         [0.4189, 0.2469, -0.5990], [-0.2900, 0.7677, 0.8188]]
    b = [0.3511, -1.6710, 6.838, -0.8843, 3.7018]
    mo, mg = build_pair(oracle, a, 3, dense)
    x, h, it, fused = gpu_solve(mg, b, 100, 1e-13)
    assert abs(x[0] - np.float32(157.611)) < 1e-2
    assert abs(x[1] + np.float32(38.0747)) < 1e-2
    assert abs(x[2] - np.float32(96.0291)) < 1e-2


@pytest.mark.parametrize("dense", [0, 1])
def test_lsqr_underdetermined_1(oracle, dense):
    mo, mg = build_pair(oracle, [[1.0, 1.0, 0.0], [2.0, 1.0, -1.0]], 3, dense)
    x, h, it, fused = gpu_solve(mg, [1.0, 0.0], 100, 1e-13)
    # the reference demands |x1| < 1e-15 absolute (tests_lsqr.f90:431); a different summation order
    # leaves a few ulps of 1.0
    assert abs(x[0]) < 1e-14
    assert comparable(x[1], 1.0, TOL) and comparable(x[2], 1.0, TOL)


@pytest.mark.parametrize("dense", [0, 1])
def test_lsqr_underdetermined_2(oracle, dense):
    mo, mg = build_pair(oracle, [[0.25] * 4], 4, dense)
    x, h, it, fused = gpu_solve(mg, [1.0], 100, 1e-14)
    assert all(comparable(v, 1.0, TOL) for v in x)


@pytest.mark.parametrize("dense", [0, 1])
def test_lsqr_underdetermined_3(oracle, dense):
    mo, mg = build_pair(oracle, [[1.0, 1.0, 1.0, 1.0], [1.0, -1.0, -1.0, 1.0]], 4, dense)
    x, h, it, fused = gpu_solve(mg, [1.0, -1.0], 100, 1e-14)
    for got, want in zip(x, (0.0, 0.5, 0.5, 0.0)):
        assert abs(got - want) < 1e-14


def test_zero_rhs_returns_zero_model(oracle):
    mo, mg = build_pair(oracle, [[1.0, 2.0], [3.0, 4.0]], 2)
    x, h, it, fused = gpu_solve(mg, [0.0, 0.0], 10, 1e-13)     # "|b| = 0, the model is exact" (:123-126)
    assert it == 0 and np.array_equal(x, np.zeros(2))


def test_wrong_sizes_abort(oracle):
    mo, mg = build_pair(oracle, [[1.0, 2.0], [3.0, 4.0]], 2)
    with pytest.raises(tfx.TfxError, match="Wrong matrix size in lsqr_solve"):
        tfx.lsqr_solve(3, 2, 10, 1e-13, 0.0, mg, np.ones(3), np.zeros(2))


def _sensit_case(orc, rng, nx, ny, nz, ndata, rate, ncons_kind, dense):
    """Random S (ndata x 2N, only the first N columns used) + a constraint matrix."""
    N = nx * ny * nz
    ncol = 2 * N
    nel = N if dense else max(1, int(rate * N))
    tfx.set_option("dense_detect", 1 if dense else 0)
    So = orc.SparseMatrix(ndata, ncol, ndata * nel)
    Sg = tfx.SparseMatrix(ndata, ncol, ndata * nel)
    for i in range(ndata):
        cols = (np.arange(N) if dense else np.sort(rng.choice(N, size=nel, replace=False))).astype(np.int32) + 1
        vals = rng.standard_normal(nel).astype(np.float32)
        for m in (So, Sg):
            m.add_row(vals, cols); m.new_row()
    So.finalize(); Sg.finalize()
    tfx.set_option("dense_detect", 1)
    if ncons_kind == "damping":          # alpha*I rows like damping.F90:158-179 (one block of N rows)
        Co = orc.SparseMatrix(N, ncol, N); Cg = tfx.SparseMatrix(N, ncol, N)
        for p in range(N):
            for m in (Co, Cg):
                m.add(0.5 * (1 + (p % 3)), p + 1); m.new_row()
    else:                                # gradient-like rows: two entries per row
        Co = orc.SparseMatrix(N, ncol, 2 * N); Cg = tfx.SparseMatrix(N, ncol, 2 * N)
        for p in range(N):
            for m in (Co, Cg):
                m.add(-0.5, p + 1)
                if p + 1 < N:
                    m.add(0.5, p + 2)
                m.new_row()
    Co.finalize(); Cg.finalize()
    b = np.concatenate([rng.standard_normal(ndata), 0.01 * rng.standard_normal(N)])
    return So, Sg, Co, Cg, b, N, ncol


@pytest.mark.parametrize("dense", [False, True])
@pytest.mark.parametrize("cons", ["damping", "gradient"])
def test_lsqr_solve_sensit_with_constraints(oracle, dense, cons, order):
    rng = np.random.default_rng(99)
    nx, ny, nz, ndata = 6, 5, 4, 24
    So, Sg, Co, Cg, b, N, ncol = _sensit_case(oracle, rng, nx, ny, nz, ndata, 0.3, cons, dense)
    niter = 400                       # run to convergence (rank 120): the converged solution is well defined
    xr, hr, itr = oracle.lsqr_solve_sensit(niter, 1e-13, 0.0, 0.0, So, Co, b, N, nx, ny, nz, 1, 1, True)
    u = b.copy(); x = np.zeros(ncol)
    tfx.lsqr_solve_sensit(len(b), ncol, niter, 1e-13, 0.0, 0.0, Sg, Cg, u, x, [1, 0], N, nx, ny, nz, 1, 1, True)
    h, it, fused = tfx.last_history()
    if order == "strict":
        assert not fused and it == itr
        assert_history(h, hr, rtol=1e-9)
        assert np.allclose(x, xr, rtol=1e-9, atol=1e-12 * np.abs(xr).max())
    else:
        assert fused == dense
        assert_history(h, hr, first=8)
        assert abs(h[-1] - hr[-1]) <= 1e-6 * hr[-1]
        assert np.allclose(x, xr, rtol=1e-6, atol=1e-8 * np.abs(xr).max())


@pytest.mark.parametrize("wtype", [1, 2])
def test_lsqr_solve_sensit_wavelet_in_loop(oracle, wtype, order):
    # WAVELET_DOMAIN = .false. with compression: 2 transforms per iteration (lsqr_solver2.F90:200-206,230-234)
    rng = np.random.default_rng(5)
    nx, ny, nz, ndata = 6, 5, 4, 20
    So, Sg, Co, Cg, b, N, ncol = _sensit_case(oracle, rng, nx, ny, nz, ndata, 0.4, "gradient", False)
    niter = 400
    xr, hr, itr = oracle.lsqr_solve_sensit(niter, 1e-13, 0.0, 0.0, So, Co, b, N, nx, ny, nz, 1, wtype, False)
    u = b.copy(); x = np.zeros(ncol)
    tfx.lsqr_solve_sensit(len(b), ncol, niter, 1e-13, 0.0, 0.0, Sg, Cg, u, x, [1, 0], N, nx, ny, nz, 1, wtype, False)
    h, it, fused = tfx.last_history()
    assert not fused
    if order == "strict":
        assert it == itr
        assert_history(h, hr, rtol=1e-9)
        assert np.allclose(x, xr, rtol=1e-9, atol=1e-12 * np.abs(xr).max())
    else:
        assert_history(h, hr, first=8)
        assert abs(h[-1] - hr[-1]) <= 1e-6 * hr[-1]
        assert np.allclose(x, xr, rtol=1e-6, atol=1e-8 * np.abs(xr).max())


def test_lsqr_solve_sensit_soft_threshold_and_misfit(oracle):
    # branch coverage in the reference's exact arithmetic (strict_order): soft threshold and misfit exit
    tfx.set_option("strict_order", 1)
    try:
        _soft_threshold_and_misfit(oracle)
    finally:
        tfx.set_option("strict_order", 0)


def test_misfit_exit_fast_kernels(oracle):
    rng = np.random.default_rng(8)
    nx, ny, nz, ndata = 5, 4, 3, 18
    So, Sg, Co, Cg, b, N, ncol = _sensit_case(oracle, rng, nx, ny, nz, ndata, 0.5, "damping", False)
    target = 0.5 * np.sqrt(np.mean(b[:ndata] ** 2))
    xr, hr, itr = oracle.lsqr_solve_sensit(200, 1e-13, 0.0, target, So, Co, b, N, nx, ny, nz, 1, 1, True)
    u = b.copy(); x = np.zeros(ncol)
    tfx.lsqr_solve_sensit(len(b), ncol, 200, 1e-13, 0.0, target, Sg, Cg, u, x, [1, 0], N, nx, ny, nz, 1, 1, True)
    h, it, fused = tfx.last_history()
    assert it == itr and 0 < it < 200
    assert np.allclose(x, xr, rtol=1e-6, atol=1e-9)


def _soft_threshold_and_misfit(oracle):
    rng = np.random.default_rng(8)
    nx, ny, nz, ndata = 5, 4, 3, 18
    So, Sg, Co, Cg, b, N, ncol = _sensit_case(oracle, rng, nx, ny, nz, ndata, 0.5, "damping", False)
    # soft thresholding (gamma /= 0, :272-275)
    xr, hr, itr = oracle.lsqr_solve_sensit(25, 1e-13, 1e-3, 0.0, So, Co, b, N, nx, ny, nz, 1, 1, True)
    u = b.copy(); x = np.zeros(ncol)
    tfx.lsqr_solve_sensit(len(b), ncol, 25, 1e-13, 1e-3, 0.0, Sg, Cg, u, x, [1, 0], N, nx, ny, nz, 1, 1, True)
    h, it, fused = tfx.last_history()
    assert it == itr
    assert_history(h, hr, rtol=1e-9)
    assert np.allclose(x, xr, rtol=1e-9, atol=1e-12)
    # misfit early exit (:168-189): pick a target the solve reaches after a few iterations
    target = 0.5 * np.sqrt(np.mean(b[:ndata] ** 2))
    xr, hr, itr = oracle.lsqr_solve_sensit(200, 1e-13, 0.0, target, So, Co, b, N, nx, ny, nz, 1, 1, True)
    assert 0 < itr < 200
    u = b.copy(); x = np.zeros(ncol)
    tfx.lsqr_solve_sensit(len(b), ncol, 200, 1e-13, 0.0, target, Sg, Cg, u, x, [1, 0], N, nx, ny, nz, 1, 1, True)
    h, it, fused = tfx.last_history()
    assert it == itr
    assert np.allclose(x, xr, rtol=1e-6, atol=1e-9)


def test_fused_equals_split_on_same_matrix(oracle):
    # the single-sweep reformulation (S vhat)/alpha must walk the same iterates as the two-product path
    rng = np.random.default_rng(21)
    nrows, ncols = 300, 2000
    A = (rng.standard_normal((nrows, ncols)) / (1 + np.arange(ncols) * 0.01)).astype(np.float32)
    b = rng.standard_normal(nrows)
    res = {}
    for dense in (0, 1):
        tfx.set_option("dense_detect", dense)
        m = tfx.SparseMatrix(nrows, ncols, nrows * ncols)
        cols = np.arange(1, ncols + 1, dtype=np.int32)
        for i in range(nrows):
            m.add_row(A[i], cols); m.new_row()
        m.finalize()
        x, h, it, fused = gpu_solve(m, b, 60, 1e-13)
        assert fused == bool(dense)
        res[dense] = (x, h, it)
    tfx.set_option("dense_detect", 1)
    assert res[0][2] == res[1][2]
    big = res[0][1] > 1e-4        # below that the iterates sit in LSQR's rounding-chaotic regime
    assert np.allclose(res[0][1][big], res[1][1][big], rtol=1e-6)
    assert np.allclose(res[0][0], res[1][0], rtol=1e-6, atol=1e-9)


@pytest.mark.parametrize("strict", [1, 0])
def test_small_rhobar_exit_iteration_count(oracle, strict):
    """The small-|rhobar| exit (`abs(rhobar) < 1.e-30`): lsqr_solve leaves the loop BEFORE `iter = iter + 1`
    (lsqr_solver2.F90:459-465), lsqr_solve_sensit after it (:281-289) -- so on this exit lsqr_solve prints one iteration
    less than it executed. A = 1e-31 * I: alpha = |A^T u| = 1e-31, so rhobar = -c * alpha is below 1e-30 after the first
    iteration whatever the rounding (the oracle executes exactly one iteration as well)."""
    n = 2
    a = np.float32(1e-31)
    A = tfx.SparseMatrix.from_arrays(n, n, np.full(n, a, dtype=np.float32), np.arange(1, n + 1, dtype=np.int32),
                                     np.arange(1, n + 2, dtype=np.int64), np.arange(1, n + 1, dtype=np.int32))
    Ao = oracle.SparseMatrix(n, n, n)
    for i in range(n):
        Ao.add(float(a), i + 1); Ao.new_row()
    Ao.finalize()
    C = tfx.SparseMatrix(0, n, 1); C.finalize()
    b = np.array([1.0, 2.0])
    xo, ho, ito = oracle.lsqr_solve(10, 0.0, 0.0, Ao, b)
    assert ito == 1 and len(ho) == 1                     # the oracle counts the executed loop bodies
    tfx.set_option("strict_order", strict)
    try:
        u = b.copy(); x = np.zeros(n)
        tfx.lsqr_solve(n, n, 10, 0.0, 0.0, A, u, x)
        h, it, _ = tfx.last_history()
        assert tfx.last_iterations() == (1, 0) and it == 1 and len(h) == 1     # executed 1, the reference prints iter - 1 = 0
        assert np.allclose(x, xo, rtol=1e-12)
        u = b.copy(); x = np.zeros(n)
        tfx.lsqr_solve_sensit(n, n, 10, 0.0, 0.0, 0.0, A, C, u, x, [1, 0], n, 1, 1, n, 1, 0, True)
        assert tfx.last_iterations() == (1, 1)                                  # here the reference prints 1
        assert np.allclose(x, xo, rtol=1e-12)
    finally:
        tfx.set_option("strict_order", 0)
