"""Joint gravity + magnetic system (BASELINE config E in miniature): one matrix_sensit holding the rows of both
problems (gravity rows first, columns 1..N; magnetic rows after, columns N+1..2N, sensitivity_gravmag.F90:685-686),
built the way the reference builds it -- initialize, one read_sensitivity_kernel per problem, finalize
(problem_joint_gravmag.F90:230-248) -- in memory (re-partitioner) and through the stream files, then a joint
LSQR solve with SOLVE_PROBLEM = (T, T). Oracle: the C restatement fed with add_row / new_row in the same order."""
import numpy as np
import pytest

import tomofastx_b200 as tfx
from tests.synth import make_problem

pytestmark = pytest.mark.gpu


def _rows(sa, ija, ijl, rowptr):
    return {int(rowptr[i]): (ija[ijl[i] - 1:ijl[i + 1] - 1], sa[ijl[i] - 1:ijl[i + 1] - 1]) for i in range(len(rowptr))}


def _problems():
    g = make_problem(nx=10, ny=8, nz=5, ndata=9, compression_type=1, rate=0.25, problem_type=1, problem_weight=1.0)
    m = make_problem(nx=10, ny=8, nz=5, ndata=7, compression_type=1, rate=0.25, problem_type=2, nmodel_components=1,
                     problem_weight=1.0)
    m.par.param_shift = m.N                      # problem 2 of a 1-component joint system
    m.par.ncolumns = g.par.ncolumns = 2 * g.N
    return g, m


def _oracle_joint(oracle, g, m):
    Sg, Sm = g.oracle_matrix(oracle), m.oracle_matrix(oracle)
    S = oracle.SparseMatrix(g.ndata + m.ndata, 2 * g.N, Sg.nel + Sm.nel)
    for part, nd in ((Sg, g.ndata), (Sm, m.ndata)):
        rows = _rows(*part.arrays())
        for r in range(1, nd + 1):
            if r in rows:
                S.add_row(rows[r][1], rows[r][0])
            S.new_row()
    S.finalize()
    return S


@pytest.mark.parametrize("via_files", [False, True])
def test_joint_matrix_and_solve(oracle, tmp_path, via_files):
    g, m = _problems()
    N = g.N
    So = _oracle_joint(oracle, g, m)
    tot = {}
    S = tfx.SparseMatrix(g.ndata + m.ndata, 2 * N, 2 * int(0.25 * N) * (g.ndata + m.ndata))
    for slot, pb in ((1, g), (2, m)):
        rows, nnz_col, cerr, tot[slot] = tfx.sensit_assemble_rows(pb.par, pb.grid, pb.data_xyz, pb.cw, pb.dw)
        if via_files:
            d = str(tmp_path / "SENSIT")
            tfx.write_sensit_file(rows, d)
            tfx.write_sensit_metadata(pb.par, d, 1, 1, cerr, tot[slot], nnz_col)
            tfx.read_sensitivity_kernel_into(S, pb.par, d, pb.dw, 1, slot, [N])
        else:
            tfx.sensit_repartition_into(S, rows, slot, [N])
        assert S.get_current_row_number() == (g.ndata if slot == 1 else g.ndata + m.ndata)
    S.finalize()
    assert S.get_number_elements() == tot[1] + tot[2]
    got, want = _rows(*S.export()), _rows(*So.arrays())
    assert set(got) == set(want)
    bad = sum(len(set(got[r][0]) ^ set(want[r][0])) for r in want)
    assert bad <= 2 * (g.ndata + m.ndata)
    assert min(got[g.ndata + 1][0]) > N and max(got[g.ndata][0]) <= N        # column blocks of the two problems
    with pytest.raises(tfx.TfxError):                                         # the builder is closed after finalize
        tfx.sensit_repartition_into(S, tfx.sensit_assemble_rows(g.par, g.grid, g.data_xyz, g.cw, g.dw)[0], 1, [N])

    # joint solve: damping block on both problems, wavelet domain
    x_true = np.concatenate([g.model_scaled_w(oracle)[:N], m.model_scaled_w(oracle)[N:2 * N]])
    b = np.concatenate([So.mult_vector(x_true), np.zeros(2 * N)])
    alpha = 1e-4
    Co = oracle.SparseMatrix(2 * N, 2 * N, 2 * N)
    Cm = tfx.SparseMatrix(2 * N, 2 * N, 2 * N)
    for p in range(2 * N):
        Co.add(alpha, p + 1); Co.new_row()
        Cm.add(alpha, p + 1); Cm.new_row()
    Co.finalize(); Cm.finalize()
    niter = 25
    x_ref, h_ref, it_ref = oracle.lsqr_solve_sensit(niter, 1e-13, 0.0, 0.0, So, Co, b, N, g.nx, g.ny, g.nz, 1, 1, True,
                                                    solve_problem=(1, 1))
    u = b.copy(); x = np.zeros(2 * N)
    tfx.lsqr_solve_sensit(len(u), 2 * N, niter, 1e-13, 0.0, 0.0, S, Cm, u, x, [1, 1], N, g.nx, g.ny, g.nz, 1, 1, True)
    h, it, fused = tfx.last_history()
    assert it == it_ref and not fused
    n = min(8, len(h_ref))
    assert np.allclose(h[:n], h_ref[:n], rtol=5e-3)          # threshold flips perturb single entries of S
    # against the device matrix itself the solve is exact to rounding: S x reproduces the data part of the fit
    r_dev = np.linalg.norm(S.mult_vector(x) - b[:g.ndata + m.ndata]) / np.linalg.norm(b)
    r_ref = np.linalg.norm(So.mult_vector(x_ref) - b[:g.ndata + m.ndata]) / np.linalg.norm(b)
    assert abs(r_dev - r_ref) <= 5e-3 * max(r_ref, 1e-6) + 1e-6



def test_joint_solve_with_cross_gradient_constraints(oracle):
    """BASELINE config E in miniature: joint gravity + magnetic system with model damping on both problems and the
    cross-gradient coupling (joint_inverse_problem.F90:436-470,529-533,578-608), solved in the physical domain
    (WAVELET_DOMAIN = F: the cross-gradient rows act on the models, the compressed kernels on their wavelet transforms,
    lsqr_solver2.F90:200-207,228-235). matrix_cons and its right-hand side are produced on the device
    (tfx_damping_add x2, tfx_cross_gradient_calculate) and by the oracle's restatements; the same sensitivity matrix is
    given to both solvers, so the residual histories must agree to the LSQR parity bar."""
    g, m = _problems()
    N, nx, ny, nz = g.N, g.nx, g.ny, g.nz
    So = _oracle_joint(oracle, g, m)
    S = tfx.SparseMatrix.from_arrays(g.ndata + m.ndata, 2 * N, *So.arrays())
    rng = np.random.default_rng(21)
    dX, dY, dZ = np.full(nx, 100.0), np.full(ny, 100.0), np.full(nz, 50.0)
    m1 = g.m_true.ravel() + 5.0 * rng.standard_normal(N)           # current models of the two problems
    m2 = m.m_true.ravel()[:N] + 1e-3 * rng.standard_normal(N)
    prior = np.zeros(N)
    ncons = 2 * N + 3 * N
    Co = oracle.SparseMatrix(ncons, 2 * N, 2 * N + 8 * 3 * N)
    Cg = tfx.SparseMatrix(ncons, 2 * N, 2 * N + 8 * 3 * N)
    bo, bg = np.zeros(ncons), np.zeros(ncons)
    alpha = (1e-6, 1e-3)
    for i, (mod, cw) in enumerate(((m1, g.cw), (m2, m.cw))):
        oracle.damping_add(Co, bo, alpha[i], 1.0, 2.0, 1, nx, ny, nz, 0, N, cw, mod, prior, i * N, False)
        tfx.damping_add(Cg, bg, alpha[i], 1.0, 2.0, 1, nx, ny, nz, cw, mod, prior, i * N, False)
    cost_o, cg_o, nnz_o, _ = oracle.cross_gradient_calculate(Co, bo, nx, ny, nz, dX, dY, dZ, 0, N, m1, m2, g.cw, m.cw, 1, 1e-4)
    cost_g, cg_g = tfx.cross_gradient_calculate(Cg, bg, nx, ny, nz, dX, dY, dZ, m1, m2, g.cw, m.cw, 1, 1e-4)
    Co.finalize(); Cg.finalize()
    assert np.array_equal(bg, bo) and np.array_equal(cg_g, cg_o)
    assert all(np.array_equal(a, b_) for a, b_ in zip(Cg.export(), Co.arrays()))
    assert nnz_o > 0 and Cg.get_number_elements() == Co.nel

    data = So.mult_vector(np.concatenate([oracle.forward_wavelet((m1 / g.cw).copy(), nx, ny, nz, 1),
                                          oracle.forward_wavelet((m2 / m.cw).copy(), nx, ny, nz, 1)]))
    b = np.concatenate([0.1 * data, bo])
    niter = 30
    x_ref, h_ref, it_ref = oracle.lsqr_solve_sensit(niter, 1e-13, 0.0, 0.0, So, Co, b, N, nx, ny, nz, 1, 1, False,
                                                    solve_problem=(1, 1))
    u = b.copy(); x = np.zeros(2 * N)
    tfx.lsqr_solve_sensit(len(u), 2 * N, niter, 1e-13, 0.0, 0.0, S, Cg, u, x, [1, 1], N, nx, ny, nz, 1, 1, False)
    h, it, fused = tfx.last_history()
    assert it == it_ref and not fused
    # the solve converges in ~6 iterations and then creeps along the damping floor (r < 1e-6), where r_k is rounding
    # noise: the 1e-6 bar applies on the way down
    descending = h_ref > 1.0e-6
    assert descending.sum() >= 5
    assert np.allclose(h[descending], h_ref[descending], rtol=1e-6), (h[:10], h_ref[:10])
    assert h[-1] < 1.0e-6 and h_ref[-1] < 1.0e-6
    assert np.allclose(x, x_ref, rtol=1e-4, atol=1e-6 * np.abs(x_ref).max())
