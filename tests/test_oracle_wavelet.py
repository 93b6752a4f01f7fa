"""Oracle vs the reference's wavelet known answers (src/tests/tests_wavelet_compression.f90)."""
import numpy as np
import pytest

from tests.conftest import TOL, comparable


def test_wavelet_calculate_data(oracle):
    # tests_wavelet_compression.f90:70-135 -- A.x == Haar(A rows).Haar(x) on a 3x4x5 grid.
    nx, ny, nz = 3, 4, 5
    ncol, nrows = nx * ny * nz, 5
    A = np.zeros((nrows, ncol))
    for j in range(1, nrows + 1):
        for i in range(1, ncol + 1):
            A[j - 1, i - 1] = float(2 * i - j) / float(i + j)
    # x(i) = dble(2*j+1) with the stale loop variable j = nrows+1 (:105-107) => x == 13.
    x = np.full(ncol, float(2 * (nrows + 1) + 1))
    b = A @ x
    Aw = np.stack([oracle.forward_wavelet(A[j], nx, ny, nz, 1) for j in range(nrows)])
    xw = oracle.forward_wavelet(x, nx, ny, nz, 1)
    b2 = Aw @ xw
    for j in range(nrows):
        assert comparable(b[j], b2[j], TOL)


def test_wavelet_diagonal_matrix(oracle):
    # tests_wavelet_compression.f90:140-182 -- exact integer golden value 46656.
    nx = ny = nz = 10
    n = nx * ny * nz
    nnz = 0
    for j in range(n):
        a = np.zeros(n)
        a[j] = 1.0
        nnz += int(np.count_nonzero(oracle.forward_wavelet(a, nx, ny, nz, 1)))
    assert nnz == 46656


@pytest.mark.parametrize("wtype", [1, 2])
def test_wavelet_norm_preserving(oracle, wtype):
    # tests_wavelet_compression.f90:187-239
    nx, ny, nz = 10, 11, 12
    x = np.arange(1, nx * ny * nz + 1, dtype=np.float64)
    xw = oracle.forward_wavelet(x, nx, ny, nz, wtype)
    assert comparable(oracle.norm2(x), oracle.norm2(xw), TOL)


@pytest.mark.parametrize("wtype", [1, 2])
def test_wavelet_inverse(oracle, wtype):
    # tests_wavelet_compression.f90:244-326 -- iW(W(e_j)) == e_j, off-diagonals < 1e-15.
    nx, ny, nz = 10, 11, 12
    n = nx * ny * nz
    nnz = 0
    for j in range(n):
        a = np.zeros(n)
        a[j] = 1.0
        r = oracle.inverse_wavelet(oracle.forward_wavelet(a, nx, ny, nz, wtype), nx, ny, nz, wtype)
        nnz += int(np.count_nonzero(r > 1e-15))
        assert comparable(r[j], 1.0, TOL)
        r[j] = 0.0
        assert np.all(np.abs(r) < 1e-15)
    assert nnz == n


@pytest.mark.parametrize("n", [1, 2, 3, 4, 7, 8, 16, 31, 32, 33, 64, 127, 128, 129, 256, 512, 1024, 2048, 4096])
def test_nscale_matches_integer_log2(oracle, n):
    # nscale = int(log(n)/log(2)) (wavelet_transform.F90:85) must equal floor(log2 n) at/near
    # powers of two: transform a delta on an n x 1 x 1 line and count the scales touched.
    import math
    a = np.zeros(n)
    a[0] = 1.0
    w = oracle.forward_wavelet(a, n, 1, 1, 1)
    # coarse coefficient after k scales is 2^(-k/2) with k = floor(log2 n).
    k = n.bit_length() - 1
    assert math.isclose(w[0], 2.0 ** (-k / 2.0), rel_tol=1e-12)
