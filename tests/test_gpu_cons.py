"""Constraint-matrix producers on the device (csrc/cons.cu) against the oracle restatements of damping.F90,
damping_gradient.F90, cross_gradient.F90 and admm_method.F90. The device writes the rows in the
reference's add() order with explicit round-to-nearest arithmetic, so the CSR arrays and right-hand sides are compared
bit for bit (except the Lp-norm multiplier, which goes through pow())."""
import numpy as np
import pytest

import tomofastx_b200 as tfx

pytestmark = pytest.mark.gpu

NX, NY, NZ = 7, 6, 5
N = NX * NY * NZ


def _setup(seed=0):
    rng = np.random.default_rng(seed)
    dX = rng.uniform(50.0, 150.0, NX); dY = rng.uniform(50.0, 150.0, NY); dZ = rng.uniform(20.0, 80.0, NZ)
    m1 = rng.standard_normal(N); m2 = rng.standard_normal(N)
    m1[10:20] = 0.0; m2[15:40] = 0.0                      # flat patches: zero derivatives -> entries dropped (add(), :219)
    cw1 = rng.uniform(0.5, 2.0, N); cw2 = rng.uniform(0.5, 2.0, N)
    lw = rng.uniform(0.5, 1.5, N)
    return rng, dX, dY, dZ, m1, m2, cw1, cw2, lw


def _same_matrix(Sg, So):
    sa, ija, ijl, rowptr = Sg.export()
    sa_o, ija_o, ijl_o, rowptr_o = So.arrays()
    assert np.array_equal(rowptr, rowptr_o)
    assert np.array_equal(ijl, ijl_o)
    assert np.array_equal(ija, ija_o)
    assert np.array_equal(sa, sa_o)


def test_damping_gradient_add(oracle):
    rng, dX, dY, dZ, m1, m2, cw1, cw2, lw = _setup(2)
    beta, pw, shift = 2e-2, 1.3, N                       # second model component
    nl = 3 * N
    So = oracle.SparseMatrix(nl, 2 * N, 2 * nl)
    Sg = tfx.SparseMatrix(nl, 2 * N, 2 * nl)
    bo, bg = np.zeros(nl), np.zeros(nl)
    for direction in (1, 2, 3):
        co = oracle.damping_gradient_add(So, bo, beta, pw, NX, NY, NZ, dX, dY, dZ, 0, N, m1, cw1, lw, shift, direction)
        cg = tfx.damping_gradient_add(Sg, bg, beta, pw, NX, NY, NZ, dX, dY, dZ, m1, cw1, lw, shift, direction)
        assert cg == pytest.approx(co, rel=1e-13)
    So.finalize(); Sg.finalize()
    _same_matrix(Sg, So)
    assert np.array_equal(bg, bo)
    with pytest.raises(tfx.TfxError, match="Wrong direction"):
        tfx.damping_gradient_add(tfx.SparseMatrix(N, 2 * N, 2 * N), np.zeros(N), beta, pw, NX, NY, NZ, dX, dY, dZ, m1, cw1,
                                 lw, 0, 4)


@pytest.mark.parametrize("der_type", [1, 2])
@pytest.mark.parametrize("keep", [(0, 0), (0, 1)])
def test_cross_gradient_calculate(oracle, der_type, keep):
    rng, dX, dY, dZ, m1, m2, cw1, cw2, lw = _setup(3)
    nl = 3 * N + 4
    So = oracle.SparseMatrix(nl, 2 * N, 8 * 3 * N)
    Sg = tfx.SparseMatrix(nl, 2 * N, 8 * 3 * N)
    bo, bg = np.zeros(nl), np.zeros(nl)
    # four host-built rows first (like a block produced by unported host code), then the device block
    for S in (So, Sg):
        for r in range(4):
            S.add(1.0 + r, 1 + r); S.add(-2.0, 2 * N - r); S.new_row()
    cost_o, cg_o, nnz_o, nne_o = oracle.cross_gradient_calculate(So, bo, NX, NY, NZ, dX, dY, dZ, 0, N, m1, m2, cw1, cw2,
                                                                  der_type, 0.37, keep)
    cost_g, cg_g = tfx.cross_gradient_calculate(Sg, bg, NX, NY, NZ, dX, dY, dZ, m1, m2, cw1, cw2, der_type, 0.37, keep)
    So.finalize(); Sg.finalize()
    _same_matrix(Sg, So)
    assert np.array_equal(bg, bo)
    assert np.array_equal(cg_g, cg_o)
    assert np.allclose(cost_g, cost_o, rtol=1e-13)
    x = rng.standard_normal(2 * N); u = rng.standard_normal(nl)
    assert np.allclose(Sg.mult_vector(x), So.mult_vector(x), rtol=1e-12, atol=1e-16)
    assert np.allclose(Sg.trans_mult_vector(u), So.trans_mult_vector(u), rtol=1e-12, atol=1e-16)


def test_cross_gradient_unsupported_derivative_aborts():
    rng, dX, dY, dZ, m1, m2, cw1, cw2, lw = _setup(4)
    S = tfx.SparseMatrix(3 * N, 2 * N, 8 * 3 * N)
    with pytest.raises(tfx.TfxError, match="Unsupported derivative type"):
        tfx.cross_gradient_calculate(S, np.zeros(3 * N), NX, NY, NZ, dX, dY, dZ, m1, m2, cw1, cw2, 3, 1.0)


def test_admm_iterate(oracle):
    rng = np.random.default_rng(5)
    n, nl = 500, 3
    lo = np.sort(rng.uniform(-10, 10, (n, nl)), axis=1)
    xmin, xmax = lo, lo + rng.uniform(0.1, 2.0, (n, nl))
    x = rng.uniform(-15, 15, n)
    z_o, u_o = rng.standard_normal(n), rng.standard_normal(n)
    z_g, u_g = z_o.copy(), u_o.copy()
    for _ in range(3):
        x0_o = oracle.admm_iterate(xmin, xmax, x, z_o, u_o)
        x0_g = tfx.admm_iterate_admm_arrays(xmin, xmax, x, z_g, u_g)
        assert np.array_equal(z_g, z_o) and np.array_equal(u_g, u_o) and np.array_equal(x0_g, x0_o)
        x = x + 0.1 * rng.standard_normal(n)


def test_reset_and_rebuild_device_built_matrix(oracle):
    """matrix_cons is reset and rebuilt before every solve (joint_inverse_problem.F90:364-373)."""
    rng, dX, dY, dZ, m1, m2, cw1, cw2, lw = _setup(6)
    S = tfx.SparseMatrix(N, N, N)
    b = np.zeros(N)
    for it in range(3):
        S.reset()
        tfx.damping_add(S, b, 1e-2 * (it + 1), 1.0, 2.0, 0, NX, NY, NZ, cw1, m1, m2, 0, True)
        S.finalize()
        So = oracle.SparseMatrix(N, N, N)
        bo = np.zeros(N)
        oracle.damping_add(So, bo, 1e-2 * (it + 1), 1.0, 2.0, 0, NX, NY, NZ, 0, N, cw1, m1, m2, 0, True)
        So.finalize()
        _same_matrix(S, So)
        assert np.array_equal(b, bo)


# ---- the reference's own known answers for the producers (src/tests/tests_inversion.f90) on the device ------------
def test_ref_golden_add_damping_identity_matrix():
    """test_add_damping_identity_matrix (tests_inversion.f90:50-127): I * (1..N) = b, assert at :117."""
    from tests.conftest import TOL, comparable
    from tests.test_oracle_ref_goldens_cons import damping_identity_case
    nx, ny, nz, ntot, nel = damping_identity_case(1)
    M = tfx.SparseMatrix(ntot, nel, nel)
    b_rhs = np.zeros(ntot)
    model = np.zeros(ntot)
    tfx.damping_add(M, b_rhs, 1.0, 1.0, 2.0, 0, nx, ny, nz, np.ones(ntot), model, model, 0, True)
    M.finalize()
    b = M.mult_vector(np.arange(1, ntot + 1, dtype=np.float64))
    for i in range(ntot):
        assert comparable(b[i], float(i + 1), TOL), i


@pytest.mark.parametrize("der_type", [1, 2])
def test_ref_golden_cross_gradient_457904(oracle, der_type):
    """test_cross_gradient_calculate (tests_inversion.f90:143-253): 457904 stored elements (:244-246); the device-built
    CSR is also bit-identical to the oracle's on this case."""
    from tests.test_oracle_ref_goldens_cons import CG_GOLDEN_NNZ, cross_gradient_case
    nx, ny, nz, n, m1, m2 = cross_gradient_case()
    one = np.ones(n)
    dX, dY, dZ = np.ones(nx), np.ones(ny), np.ones(nz)
    Sg = tfx.SparseMatrix(3 * n, 2 * n, 8 * 3 * n)
    bg = np.zeros(3 * n)
    tfx.cross_gradient_calculate(Sg, bg, nx, ny, nz, dX, dY, dZ, m1, m2, one, one, der_type, 1.0)
    Sg.finalize()
    assert Sg.get_number_elements() == CG_GOLDEN_NNZ
    So = oracle.SparseMatrix(3 * n, 2 * n, 8 * 3 * n)
    bo = np.zeros(3 * n)
    oracle.cross_gradient_calculate(So, bo, nx, ny, nz, dX, dY, dZ, 0, n, m1, m2, one, one, der_type, 1.0)
    So.finalize()
    _same_matrix(Sg, So)
    assert np.array_equal(bg, bo)
