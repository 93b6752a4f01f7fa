"""Config A (Parfile_mansf_slice) on the device through the C ABI vs the oracle.

Parity bar (north-star): LSQR residuals r_k within 1e-6 relative of the reference path. The comparison is
teacher-forced per major iteration (both solvers get the oracle's model / ADMM state) so that the
discontinuous ADMM projection cannot amplify last-bit differences across major iterations; an end-to-end
free run is compared on the data cost."""
import numpy as np
import pytest

import tomofastx_b200 as tfx
from tests import mansf

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pair(oracle):
    cfg = mansf.Config()
    io = mansf.Inversion(mansf.OracleBackend(oracle), cfg, oracle.admm_iterate)
    ig = mansf.Inversion(mansf.TfxBackend(tfx), cfg, oracle.admm_iterate)   # ADMM projection is host logic
    return cfg, io, ig


def test_assembly_matches(pair):
    cfg, io, ig = pair
    assert ig.be.nnz_total == 314368
    assert abs(ig.comp_error - io.comp_error) < 1e-9
    assert np.allclose(ig.d_obs, io.d_obs, rtol=1e-7, atol=1e-9 * np.abs(io.d_obs).max())
    sa_o, ija_o, ijl_o, rp_o = io.S.arrays()
    sa_g, ija_g, ijl_g, rp_g = ig.S.export()
    assert np.array_equal(ija_g, ija_o) and np.array_equal(ijl_g, ijl_o) and np.array_equal(rp_g, rp_o)
    # values: identical except where the f64 value (differing by ~1e-13 rel. through libm) rounds to the
    # neighbouring real(4)
    ulp = np.abs(sa_o) * 2.0 ** -23
    assert np.all(np.abs(sa_g - sa_o) <= ulp)
    assert np.mean(sa_g != sa_o) < 1e-3


def test_teacher_forced_residuals(pair):
    """r_k of 6 major iterations x 100 LSQR iterations.
    strict_order (reference summation order): every r_k within 1e-6 relative of the oracle -- the
    north-star bar. Fast kernels (tree-order sums): LSQR's iterates 13-40 are chaotic under last-bit
    perturbations (the oracle moves by ~1e-3 there when b is scaled by 1+2e-16), so they are held to the
    1e-6 bar on the stable iterations (first 10, last 40) and only bounded (5e-2) in the window where the
    oracle's own perturbation envelope exceeds 1e-6; the solution must agree to 1e-8."""
    cfg, io, ig = pair
    # LSQR parity is measured on the SAME matrix: the oracle's CSR is handed to the device through the
    # reference's own storage (assembly parity is test_assembly_matches' job).
    Sg = tfx.SparseMatrix.from_arrays(cfg.ndata, cfg.ncolumns, *io.S.arrays())
    worst_strict = worst_fast = 0.0
    for major in range(6):
        b = io.build_rhs()                       # oracle state drives both solvers (teacher forcing)
        xo, ho = io.be.solve(cfg, io.S, io.C, b)
        env = np.zeros_like(ho)                  # oracle's own sensitivity to rounding
        for scale in (1.0 + 2.3e-16, 3.0, 1.0 / 3.0):
            xp, hp = io.be.solve(cfg, io.S, io.C, b * scale)
            env = np.maximum(env, np.abs(hp - ho) / ho)
        tfx.set_option("strict_order", 1)
        xs, hs = ig.be.solve(cfg, Sg, ig.C, b)
        tfx.set_option("strict_order", 0)
        xg, hg = ig.be.solve(cfg, Sg, ig.C, b)
        assert len(hs) == len(hg) == len(ho) == cfg.niter
        rel_s = np.abs(hs - ho) / ho
        rel_f = np.abs(hg - ho) / ho
        worst_strict = max(worst_strict, rel_s.max())
        worst_fast = max(worst_fast, rel_f.max())
        assert rel_s.max() < 1e-6, (major, rel_s.max(), rel_s.argmax())
        assert np.allclose(xs, xo, rtol=1e-6, atol=1e-9 * np.abs(xo).max())
        stable = np.r_[0:10, 60:100]
        assert rel_f[stable].max() < 1e-6, (major, rel_f[stable].max())
        assert env[10:60].max() > 1e-6            # the oracle itself is not reproducible there
        assert rel_f[10:60].max() < 5e-2, (major, rel_f.max())
        assert np.abs(xg - xo).max() < 1e-8 * np.abs(xo).max()
        io.histories.append(ho)
        io.apply(xo)
    print("worst relative residual difference: strict_order %.3e, fast kernels %.3e" % (worst_strict, worst_fast))


def test_free_run_costs(oracle):
    cfg = mansf.Config()
    io = mansf.Inversion(mansf.OracleBackend(oracle), cfg, oracle.admm_iterate)
    ig = mansf.Inversion(mansf.TfxBackend(tfx), cfg, oracle.admm_iterate)
    for major in range(10):
        io.step()
        ig.step()
    assert np.allclose(ig.costs, io.costs, rtol=1e-3)
    assert np.abs(ig.m - io.m).max() < 1e-3 * np.abs(io.m).max()


def test_device_resident_major_loop(oracle):
    """The major loop with the model, the ADMM state, matrix_cons and the right-hand side produced and kept on the
    device (mansf.DeviceInversion) against the oracle's loop: the right-hand side of the first major iteration is
    bit-identical, the free run tracks the oracle like the host-orchestrated one does."""
    cfg = mansf.Config()
    io = mansf.Inversion(mansf.OracleBackend(oracle), cfg, oracle.admm_iterate)
    idev = mansf.DeviceInversion(tfx, cfg)
    assert np.allclose(idev.d_obs, io.d_obs, rtol=1e-7, atol=1e-9 * np.abs(io.d_obs).max())
    for major in range(10):
        bo, xo, ho = io.step()
        bd, xd, hd = idev.step()
        if major == 0:
            # m = 0, z = u = 0: the constraint part is exactly the oracle's; the data part differs by the matrices
            assert np.array_equal(bd[cfg.ndata:], bo[cfg.ndata:])
            assert np.allclose(bd[:cfg.ndata], bo[:cfg.ndata], rtol=1e-7, atol=1e-9 * np.abs(bo[:cfg.ndata]).max())
            assert np.allclose(hd[:10], ho[:10], rtol=1e-5)
    assert np.allclose(idev.costs, io.costs, rtol=1e-3)
    m_dev = idev.m.numpy()
    assert np.abs(m_dev - io.m).max() < 1e-3 * np.abs(io.m).max()
    assert len(idev.admm_costs) == 10 and all(np.isfinite(idev.admm_costs))


def test_all_60_major_iterations(oracle):
    """Config A in full: 60 major x 100 LSQR iterations (parfiles/Parfile_mansf_slice.txt:58-59).
    (1) teacher-forced, strict_order: ALL 6000 residuals r_k within 1e-6 relative of the oracle -- the north-star bar on
        the whole Parfile, not on its first major iterations;
    (2) free run on the benchmarked fast kernels: final data cost and model against the oracle's own free run and the
        surveyor's figures (cost ~9e-11, model range [-19.95, 260.0])."""
    cfg = mansf.Config()
    io = mansf.Inversion(mansf.OracleBackend(oracle), cfg, oracle.admm_iterate)
    ig = mansf.Inversion(mansf.TfxBackend(tfx), cfg, oracle.admm_iterate)
    Sg = tfx.SparseMatrix.from_arrays(cfg.ndata, cfg.ncolumns, *io.S.arrays())
    worst = 0.0
    tfx.set_option("strict_order", 1)
    try:
        for major in range(60):
            b = io.build_rhs()
            xo, ho = io.be.solve(cfg, io.S, io.C, b)
            xs, hs = ig.be.solve(cfg, Sg, ig.C, b)
            assert len(hs) == len(ho) == cfg.niter
            rel = np.abs(hs - ho) / ho
            worst = max(worst, rel.max())
            assert rel.max() < 1e-6, (major, rel.max(), rel.argmax())
            assert np.allclose(xs, xo, rtol=1e-6, atol=1e-9 * np.abs(xo).max())
            io.histories.append(ho)
            io.apply(xo)
    finally:
        tfx.set_option("strict_order", 0)
    assert 5e-11 < io.costs[-1] < 2e-10
    print("config A, 60 x 100 iterations, strict_order: worst relative residual difference %.3e" % worst)
    # free run, fast kernels
    for major in range(60):
        ig.step()
    assert ig.costs[-1] < 1e-9 and ig.costs[-1] == pytest.approx(io.costs[-1], rel=0.5)
    assert -19.96 < ig.m.min() < -19.9 and 259.9 < ig.m.max() < 260.0
    assert np.abs(ig.m - io.m).max() < 2e-2 * np.abs(io.m).max()
