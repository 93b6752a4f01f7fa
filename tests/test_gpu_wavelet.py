"""GPU wavelet kernels (tomofast-x_b200/csrc/wavelet.cu) vs the oracle and the reference's known answers.
The kernels use the reference's per-element operation order with explicit rounding, so the comparison
with the CPU restatement is BIT-EXACT."""
import numpy as np
import pytest

import tomofastx_b200 as tfx
from tests.conftest import TOL, comparable

pytestmark = pytest.mark.gpu

SHAPES = [(3, 4, 5), (10, 11, 12), (2, 128, 32), (67, 67, 30), (1, 1, 7), (5, 1, 1), (1, 9, 1), (64, 64, 16),
          (33, 17, 65), (256, 3, 2), (2, 2, 2), (1, 1, 1), (130, 70, 9), (20, 300, 33), (600, 40, 18), (16, 16, 1030)]


@pytest.fixture(params=[1, 0], ids=["cols", "generic"])
def kernel_choice(request):
    """Both kernels of wavelet.cu: the column-layout kernel (default where the axis fits its tile) and the generic
    (line, pair) kernel every other shape falls back to."""
    tfx.set_option("wavelet_cols", request.param)
    yield request.param
    tfx.set_option("wavelet_cols", 1)


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("wtype", [1, 2])
def test_forward_inverse_bit_exact(oracle, shape, wtype, kernel_choice):
    n1, n2, n3 = shape
    rng = np.random.default_rng(1234 + n1 * 7 + n2 * 3 + n3)
    x = rng.standard_normal(n1 * n2 * n3)
    want_f = oracle.forward_wavelet(x, n1, n2, n3, wtype)
    got_f = tfx.forward_wavelet(x.copy(), n1, n2, n3, wtype)
    assert np.array_equal(got_f, want_f)
    want_i = oracle.inverse_wavelet(want_f, n1, n2, n3, wtype)
    got_i = tfx.inverse_wavelet(want_f.copy(), n1, n2, n3, wtype)
    assert np.array_equal(got_i, want_i)


def test_named_entry_points(oracle):
    n1, n2, n3 = 10, 11, 12
    x = np.arange(1, n1 * n2 * n3 + 1, dtype=np.float64)
    assert np.array_equal(tfx.Haar3D(x.copy(), n1, n2, n3), oracle.forward_wavelet(x, n1, n2, n3, 1))
    assert np.array_equal(tfx.DaubD43D(x.copy(), n1, n2, n3), oracle.forward_wavelet(x, n1, n2, n3, 2))
    h = oracle.forward_wavelet(x, n1, n2, n3, 1)
    assert np.array_equal(tfx.iHaar3D(h.copy(), n1, n2, n3), oracle.inverse_wavelet(h, n1, n2, n3, 1))
    d = oracle.forward_wavelet(x, n1, n2, n3, 2)
    assert np.array_equal(tfx.iDaubD43D(d.copy(), n1, n2, n3), oracle.inverse_wavelet(d, n1, n2, n3, 2))
    with pytest.raises(tfx.TfxError, match="Unknown wavelet type"):
        tfx.forward_wavelet(x.copy(), n1, n2, n3, 3)


def test_wavelet_diagonal_matrix_46656():
    # tests_wavelet_compression.f90:140-182 through the C ABI (batched as a device-resident volume set).
    import torch
    nx = ny = nz = 10
    n = nx * ny * nz
    eye = torch.eye(n, dtype=torch.float64, device="cuda")
    nnz = 0
    for j in range(n):
        row = eye[j].clone()
        torch.cuda.synchronize()          # libtfx runs on its own non-blocking stream: order torch's work first
        tfx.forward_wavelet(row, nx, ny, nz, 1)
        nnz += int(torch.count_nonzero(row).item())
    assert nnz == 46656


@pytest.mark.parametrize("shape", [(256, 256, 64), (512, 512, 16), (1024, 96, 40), (1024, 1024, 3), (512, 100, 7),
                                   (128, 1000, 5), (40, 24, 200)])
@pytest.mark.parametrize("wtype", [1, 2])
def test_large_volume_kernels_agree_bit_for_bit(shape, wtype):
    """At sizes the oracle does not visit in seconds: column-layout kernel (with and without the L2-blocked slabs, with
    and without the fused Haar axis-1 + low axis-2 scales pass) == generic kernel, bit for bit, forward and inverse."""
    import torch
    n1, n2, n3 = shape
    g = torch.Generator(device="cuda").manual_seed(11)
    x = torch.randn(n1 * n2 * n3, dtype=torch.float64, device="cuda", generator=g)
    outs = []
    try:
        for cols, slab, fuse in ((0, 0, 1), (1, 0, 0), (1, 0, 1), (1, 4, 1)):
            tfx.set_option("wavelet_cols", cols); tfx.set_option("wavelet_slab_mb", slab)
            tfx.set_option("wavelet_fuse12", fuse)
            y = x.clone()
            torch.cuda.synchronize()      # libtfx runs on its own non-blocking stream: order torch's work first
            tfx.forward_wavelet(y, n1, n2, n3, wtype)
            z = y.clone()
            torch.cuda.synchronize()
            tfx.inverse_wavelet(z, n1, n2, n3, wtype)
            outs.append((y, z))
    finally:
        tfx.set_option("wavelet_cols", 1); tfx.set_option("wavelet_slab_mb", 0); tfx.set_option("wavelet_fuse12", 1)
    for y, z in outs[1:]:
        assert torch.equal(y, outs[0][0]) and torch.equal(z, outs[0][1])
    assert float((outs[0][1] - x).abs().max()) < 1e-11


@pytest.mark.parametrize("wtype", [1, 2])
def test_norm_preserving_and_roundtrip_full_size(wtype):
    # size-independent properties at the bench grid (256 x 256 x 64): orthonormality and exact inverse
    import torch
    n1, n2, n3 = 256, 256, 64
    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.randn(n1 * n2 * n3, dtype=torch.float64, device="cuda", generator=g)
    y = x.clone()
    torch.cuda.synchronize()              # libtfx runs on its own non-blocking stream: order torch's work first
    tfx.forward_wavelet(y, n1, n2, n3, wtype)
    assert comparable(float(torch.linalg.norm(x)), float(torch.linalg.norm(y)), 1e-12)
    tfx.inverse_wavelet(y, n1, n2, n3, wtype)
    assert float((y - x).abs().max()) < 1e-11


def test_apply_wavelet_transform_components(oracle):
    # wavelet_utils.F90:37-72 for nbproc = 1: v(nelements, ncomponents, nproblems)
    nx, ny, nz, ncomp = 6, 5, 4, 3
    n = nx * ny * nz
    rng = np.random.default_rng(3)
    v = rng.standard_normal(2 * ncomp * n)
    want = v.copy()
    for k in range(ncomp):          # only problem 2 is active
        o = (1 * ncomp + k) * n
        want[o:o + n] = oracle.forward_wavelet(want[o:o + n], nx, ny, nz, 2)
    got = tfx.apply_wavelet_transform(n, nx, ny, nz, ncomp, v.copy(), True, 2, 2, [0, 1])
    assert np.array_equal(got, want)
