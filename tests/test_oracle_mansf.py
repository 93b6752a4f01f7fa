"""Oracle-only run of config A (Parfile_mansf_slice) on the CPU: pins the restatement against the
surveyor's sanity figures (SURVEY.md section 8c / BASELINE.md section 5 -- NOT reference output; the
reference ships no golden outputs for Parfile runs, so end-to-end parity is otherwise unpinned)."""
import numpy as np
import pytest

from tests import mansf


@pytest.fixture(scope="module")
def inv(oracle):
    cfg = mansf.Config()
    return mansf.Inversion(mansf.OracleBackend(oracle), cfg, oracle.admm_iterate)


def test_assembly_figures(inv):
    cfg = inv.cfg
    assert cfg.nel_compressed == 1228                       # int(0.15 * 8192)
    assert inv.S.nel == 314368                              # 256 x 1228, no ties
    assert abs(inv.comp_error - 2.154e-3) < 2e-5            # compression error r
    assert abs(np.linalg.norm(inv.d_obs) - 2.4011e-4) < 2e-8


def test_first_major_iteration_residuals(inv):
    b, x, hist = inv.step()
    assert len(hist) == 100                                 # every solve runs the full 100 iterations
    assert np.allclose(hist[:3], [0.2475, 0.1003, 0.0585], atol=6e-4)
    assert abs(hist[-1] - 8.78e-3) < 2e-4
    assert abs(inv.costs[-1] - 3.03e-4) < 2e-5              # relative data cost after major iteration 1


def test_residual_history_is_chaotic_under_rounding(oracle):
    """The reference algorithm itself does not reproduce its mid-phase residuals under a last-bit change:
    scaling b by (1 + 2.3e-16) -- an exact invariance of LSQR's r_k in real arithmetic -- moves r_k by
    > 1e-5 relative somewhere in iterations 12-40, while early/late iterates and the solution agree.
    This bounds what ANY implementation with a different summation order can match (parity)."""
    cfg = mansf.Config()
    io = mansf.Inversion(mansf.OracleBackend(oracle), cfg, oracle.admm_iterate)
    b = io.build_rhs()
    x0, h0 = io.be.solve(cfg, io.S, io.C, b)
    x1, h1 = io.be.solve(cfg, io.S, io.C, b * (1.0 + 2.3e-16))
    rel = np.abs(h1 - h0) / h0
    assert rel[:10].max() < 1e-12
    assert rel[12:40].max() > 1e-5
    assert rel[60:].max() < 1e-6
    assert np.abs(x1 - x0).max() < 1e-8 * np.abs(x0).max()


def test_all_60_major_iterations(oracle):
    """Parfile_mansf_slice in full (60 major x 100 LSQR iterations, parfiles/Parfile_mansf_slice.txt:58-59): the
    surveyor's end-of-run figures: relative data cost ~9e-11, model inside the ADMM bounds [-20, 260]."""
    cfg = mansf.Config()
    io = mansf.Inversion(mansf.OracleBackend(oracle), cfg, oracle.admm_iterate)
    for _ in range(60):
        io.step()
    assert all(len(h) == 100 for h in io.histories)          # every solve hits niter
    assert 5e-11 < io.costs[-1] < 2e-10
    assert -19.96 < io.m.min() < -19.9 and 259.9 < io.m.max() < 260.0
