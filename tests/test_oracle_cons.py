"""Self-consistency of the oracle's constraint producers (oracle/tfx_oracle.c: damping.F90, damping_gradient.F90,
cross_gradient.F90, admm_method.F90 restatements). No reference test covers them ("parity unpinned"), so the
restatements are checked against the mathematics they implement: the matrix rows must be the Jacobian of the
right-hand side they come with."""
import numpy as np
import pytest

NX, NY, NZ = 6, 5, 4
N = NX * NY * NZ


def _grid(rng):
    return rng.uniform(50, 150, NX), rng.uniform(50, 150, NY), rng.uniform(20, 80, NZ)


def test_damping_rows_and_rhs(oracle):
    rng = np.random.default_rng(0)
    m, ref, cw, lw = rng.standard_normal(N), rng.standard_normal(N), rng.uniform(0.5, 2, N), rng.uniform(0.5, 1.5, N)
    alpha, pw = 3e-2, 0.7
    C = oracle.SparseMatrix(N, N, N)
    b = np.zeros(N)
    cost = oracle.damping_add(C, b, alpha, pw, 2.0, 0, NX, NY, NZ, 0, N, cw, m, ref, 0, True, lw)
    C.finalize()
    sa, ija, ijl, rowptr = C.arrays()
    assert np.array_equal(ija, np.arange(1, N + 1)) and np.array_equal(rowptr, np.arange(1, N + 1))
    assert np.allclose(sa, (alpha * pw * lw).astype(np.float32))
    # alpha I d(Wm) = -alpha (m - m_ref)/cw : the block's residual at the current model (damping.F90:88-93)
    assert np.allclose(b, -alpha * pw * lw * (m - ref) / cw, rtol=1e-15)
    assert cost == pytest.approx(np.sum(b * b), rel=1e-14)
    # wavelet domain: the right-hand side is the transform of the scaled difference, the cost is invariant (orthonormal)
    C2 = oracle.SparseMatrix(N, N, N)
    b2 = np.zeros(N)
    cost2 = oracle.damping_add(C2, b2, alpha, pw, 2.0, 1, NX, NY, NZ, 0, N, cw, m, ref, 0, True, None)
    want = -alpha * pw * oracle.forward_wavelet(((m - ref) / cw).copy(), NX, NY, NZ, 1)
    assert np.allclose(b2, want, rtol=1e-14, atol=1e-16)
    assert cost2 == pytest.approx(np.sum((alpha * pw * (m - ref) / cw) ** 2), rel=1e-12)


@pytest.mark.parametrize("direction", [1, 2, 3])
def test_damping_gradient_rows_are_the_jacobian_of_the_rhs(oracle, direction):
    rng = np.random.default_rng(direction)
    dX, dY, dZ = _grid(rng)
    m, cw, lw = rng.standard_normal(N), rng.uniform(0.5, 2, N), rng.uniform(0.5, 1.5, N)
    beta, pw = 2e-2, 1.3
    C = oracle.SparseMatrix(N, N, 2 * N)
    b = np.zeros(N)
    cost = oracle.damping_gradient_add(C, b, beta, pw, NX, NY, NZ, dX, dY, dZ, 0, N, m, cw, lw, 0, direction)
    C.finalize()
    # the unknown is d(m / cw): rows * (m / cw) = pw beta lw * forward difference = -b  (damping_gradient.F90:177-189)
    assert np.allclose(C.mult_vector(m / cw), -b, rtol=1e-5, atol=1e-6 * np.abs(b).max())     # real(4) entries, cancellation
    vol = m.reshape(NZ, NY, NX)
    step = {1: dX[None, None, :], 2: dY[None, :, None], 3: dZ[:, None, None]}[direction]
    ax = {1: 2, 2: 1, 3: 0}[direction]
    fd = np.zeros_like(vol)
    sl_lo = [slice(None)] * 3; sl_hi = [slice(None)] * 3
    sl_lo[ax] = slice(0, -1); sl_hi[ax] = slice(1, None)
    fd[tuple(sl_lo)] = (vol[tuple(sl_hi)] - vol[tuple(sl_lo)]) / np.broadcast_to(step, vol.shape)[tuple(sl_lo)]
    assert np.allclose(b, (-pw * beta * fd * lw.reshape(vol.shape)).ravel(), rtol=1e-14, atol=1e-18)
    assert cost == pytest.approx(np.sum(fd ** 2), rel=1e-13)
    assert C.nl_nonempty == N - N // {1: NX, 2: NY, 3: NZ}[direction]       # the last layer has no forward neighbour


@pytest.mark.parametrize("der_type", [1, 2])
def test_cross_gradient_rows_are_the_jacobian_of_tau(oracle, der_type):
    """tau = grad m1 x grad m2 is bilinear in (m1, m2): the rows are d tau / d (m1, m2) (times the column weights), so
    tau(m + d) - tau(m) = J d + tau_of_the_increments, exactly."""
    rng = np.random.default_rng(10 + der_type)
    dX, dY, dZ = _grid(rng)
    m1, m2 = rng.standard_normal(N), rng.standard_normal(N)
    d1, d2 = 1e-5 * rng.standard_normal(N), 1e-5 * rng.standard_normal(N)
    one = np.ones(N)
    w = 0.37

    def build(a, b_):
        C = oracle.SparseMatrix(3 * N, 2 * N, 8 * 3 * N)
        rhs = np.zeros(3 * N)
        cost, cg, nnz, nne = oracle.cross_gradient_calculate(C, rhs, NX, NY, NZ, dX, dY, dZ, 0, N, a, b_, one, one, der_type, w)
        C.finalize()
        return C, rhs, cost, cg

    C, rhs, cost, cg = build(m1, m2)
    _, rhs_p, _, _ = build(m1 + d1, m2 + d2)
    _, rhs_dd, _, _ = build(d1, d2)                       # the second-order term of the bilinear form
    lin = C.mult_vector(np.concatenate([d1, d2]))
    # rhs = -w tau  (cross_gradient.F90:323)
    assert np.allclose(-(rhs_p - rhs) - (-rhs_dd), lin, rtol=2e-6, atol=2e-7 * np.abs(lin).max())   # real(4) entries
    tau = (-rhs / w).reshape(N, 3)
    assert np.allclose(cg, np.linalg.norm(tau, axis=1), rtol=1e-14, atol=1e-300)
    assert np.allclose(cost, (tau ** 2).sum(axis=0), rtol=1e-13)
    # cells on a left AND a right boundary are skipped (:262-266), e.g. (i, j, k) = (1, ny, 1)
    assert np.all(tau[(NY - 1) * NX] == 0.0) and np.any(tau[0] != 0.0)
    # identical models have parallel gradients: tau = 0 everywhere
    _, rhs0, cost0, _ = build(m1, m1.copy())
    assert np.abs(rhs0).max() < 1e-18 and np.all(cost0 < 1e-30)


def test_cross_gradient_keep_model_constant_and_slabs(oracle):
    rng = np.random.default_rng(5)
    dX, dY, dZ = _grid(rng)
    m1, m2, cw1, cw2 = rng.standard_normal(N), rng.standard_normal(N), rng.uniform(0.5, 2, N), rng.uniform(0.5, 2, N)
    full = oracle.SparseMatrix(3 * N, 2 * N, 8 * 3 * N)
    b = np.zeros(3 * N)
    _, _, nnz_full, _ = oracle.cross_gradient_calculate(full, b, NX, NY, NZ, dX, dY, dZ, 0, N, m1, m2, cw1, cw2, 1, 1.0)
    full.finalize()
    half = oracle.SparseMatrix(3 * N, 2 * N, 8 * 3 * N)
    b2 = np.zeros(3 * N)
    _, _, nnz_half, _ = oracle.cross_gradient_calculate(half, b2, NX, NY, NZ, dX, dY, dZ, 0, N, m1, m2, cw1, cw2, 1, 1.0, (0, 1))
    half.finalize()
    assert nnz_half == nnz_full // 2 and np.array_equal(b, b2)
    sa, ija, _, _ = half.arrays()
    assert ija.max() <= N                                      # no entries in the second model's columns
    # two column slabs see the two halves of every row (local column indices, cross_gradient.F90:305-325)
    x = np.concatenate([rng.standard_normal(N), rng.standard_normal(N)])
    total = np.zeros(3 * N)
    for nsm, nel in ((0, 70), (70, N - 70)):
        slab = oracle.SparseMatrix(3 * N, 2 * nel, 8 * 3 * N)
        bs = np.zeros(3 * N)
        oracle.cross_gradient_calculate(slab, bs, NX, NY, NZ, dX, dY, dZ, nsm, nel, m1, m2, cw1, cw2, 1, 1.0)
        slab.finalize()
        assert np.array_equal(bs, b)
        total += slab.mult_vector(np.concatenate([x[nsm:nsm + nel], x[N + nsm:N + nsm + nel]]))
    assert np.allclose(total, full.mult_vector(x), rtol=1e-12, atol=1e-15)


def test_admm_projection(oracle):
    rng = np.random.default_rng(7)
    n = 400
    lo = np.sort(rng.uniform(-10, 10, (n, 2)), axis=1)
    xmin, xmax = lo, lo + rng.uniform(0.1, 2.0, (n, 2))
    x = rng.uniform(-15, 15, n)
    z, u = np.zeros(n), np.zeros(n)
    x0 = oracle.admm_iterate(xmin, xmax, x, z, u)
    inside = ((xmin <= z[:, None]) & (z[:, None] <= xmax)).any(axis=1)
    assert inside.all()                                        # z = P_C(x + u) lies in one of the intervals
    was_inside = ((xmin <= x[:, None]) & (x[:, None] <= xmax)).any(axis=1)
    assert np.array_equal(z[was_inside], x[was_inside])        # projection is the identity on C
    d = np.minimum(np.abs(xmin - x[:, None]), np.abs(xmax - x[:, None])).min(axis=1)
    assert np.allclose(np.abs(z - x)[~was_inside], d[~was_inside], rtol=1e-15)   # closest boundary (admm_method.F90:106-122)
    assert np.allclose(u, x - z) and np.allclose(x0, z - u)    # u = u + x - z ; x0 = z - u (:129-131)
