"""The two known answers the reference's own unit tests hold for the constraint producers, reproduced on
the oracle -- these PIN orc_damping_add and orc_cross_gradient_calculate to the reference:

 * test_add_damping_identity_matrix   (src/tests/tests_inversion.f90:50-127, assert at :117): the damping block built
   with alpha = problem_weight = column_weight = 1 on a 10 x 72 x 4 grid is the identity: I * (1, 2, ..., N) = b.
 * test_cross_gradient_calculate      (src/tests/tests_inversion.f90:143-253, asserts at :244-246): on a 20 x 20 x 144
   grid of unit cells with model1 = i, model2 = i + 1 the finalized cross-gradient matrix stores exactly 457904
   elements, for derivative type 1 (forward) and 2 (central), on any number of ranks.
"""
import numpy as np
import pytest

from tests.conftest import TOL, comparable

CG_GOLDEN_NNZ = 457904          # tests_inversion.f90:244,246


def damping_identity_case(nbproc):
    nx, ny, nz = 10, 72, 4      # tests_inversion.f90:65-67
    ntot = nx * ny * nz
    assert ntot % nbproc == 0
    return nx, ny, nz, ntot, ntot // nbproc


@pytest.mark.parametrize("nbproc", [1, 2, 3, 8])
def test_add_damping_identity_matrix(oracle, nbproc):
    nx, ny, nz, ntot, nel = damping_identity_case(nbproc)
    b = np.zeros(ntot)
    for rank in range(nbproc):                                            # the ranks of the reference's MPI run
        x = nel * rank + np.arange(1, nel + 1, dtype=np.float64)          # :86-89
        M = oracle.SparseMatrix(ntot, nel, nel)                           # :91-92 (ndata = 0)
        b_rhs = np.zeros(ntot)
        model = np.zeros(ntot)                                            # model%initialize: val = val_prior = 0
        oracle.damping_add(M, b_rhs, 1.0, 1.0, 2.0, 0, nx, ny, nz, nel * rank, nel, np.ones(ntot), model, model, 0, True)
        M.finalize()
        assert M.nl_nonempty == nel
        b += M.mult_vector(x)                                             # :105-107: mult_vector + MPI_Allreduce
    for i in range(ntot):
        assert comparable(b[i], float(i + 1), TOL), i                     # :115-118


def cross_gradient_case():
    nx, ny, nz = 20, 20, 144    # tests_inversion.f90:166-168 ("changing these dimensions will affect the test result")
    n = nx * ny * nz
    i = np.tile(np.arange(1, nx + 1, dtype=np.float64), ny * nz)         # :198-199: model1 = i, model2 = i + 1
    return nx, ny, nz, n, i, i + 1.0


@pytest.mark.parametrize("der_type", [1, 2])
@pytest.mark.parametrize("nbproc", [1, 2, 4])
def test_cross_gradient_calculate_457904(oracle, der_type, nbproc):
    nx, ny, nz, n, m1, m2 = cross_gradient_case()
    nel = n // nbproc
    one = np.ones(n)
    d = np.ones(max(nx, ny, nz))                                          # unit cells: X2 - X1 = 1 (:201-206)
    total = 0
    for rank in range(nbproc):
        M = oracle.SparseMatrix(3 * n, 2 * nel, 8 * 3 * n)
        rhs = np.zeros(3 * n)
        oracle.cross_gradient_calculate(M, rhs, nx, ny, nz, d[:nx], d[:ny], d[:nz], nel * rank, nel, m1, m2, one, one,
                                        der_type, 1.0)
        M.finalize()
        total += len(M.arrays()[0])                                       # get_number_elements (:238), summed (:240)
    assert total == CG_GOLDEN_NNZ
