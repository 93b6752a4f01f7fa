"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads without a GPU and exports
every symbol include/tfx.h declares; compute entry points fail loudly (no CPU fallback)."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

import tomofastx_b200 as tfx
from tests.conftest import TOL, comparable

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def libtfx():
    tfx.build()
    return tfx.lib()


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "tfx.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(tfx_[A-Za-z0-9_]+)\s*\(", src)))


def test_header_declares_the_reference_surface():
    names = declared_symbols()
    # one entry per reference procedure on the hot path (sparse_matrix.f90, wavelet_transform.F90,
    # lsqr_solver2.F90, sensitivity_gravmag.F90)
    for needed in ["tfx_sparse_matrix_initialize", "tfx_sparse_matrix_add", "tfx_sparse_matrix_new_row",
                   "tfx_sparse_matrix_finalize", "tfx_sparse_matrix_add_mult_vector",
                   "tfx_sparse_matrix_part_mult_vector", "tfx_sparse_matrix_add_trans_mult_vector",
                   "tfx_forward_wavelet", "tfx_inverse_wavelet", "tfx_Haar3D", "tfx_iDaubD43D",
                   "tfx_lsqr_solve", "tfx_lsqr_solve_sensit", "tfx_calculate_sensit"]:
        assert needed in names


def test_library_exports_every_declared_symbol(libtfx):
    out = subprocess.check_output(["nm", "-D", "--defined-only", tfx.LIB_PATH], text=True)
    exported = set(line.split()[-1] for line in out.splitlines() if line.strip())
    missing = [s for s in declared_symbols() if s not in exported]
    assert not missing, missing
    for s in declared_symbols():
        assert hasattr(libtfx, s)


def test_header_is_plain_c_and_links(libtfx, tmp_path):
    """include/tfx.h is the drop-in boundary of a Fortran / C host: it must compile as C99 (no C++ in the signatures) and
    a C translation unit that takes the address of the new entry points must link against libtfx.so."""
    src = tmp_path / "host.c"
    src.write_text(
        '#include "tfx.h"\n#include <stdio.h>\n'
        "typedef void (*fn)(void);\n"
        "int main(void) {\n"
        "  fn p[] = {(fn)tfx_lsqr_solve_sensit, (fn)tfx_calculate_sensit, (fn)tfx_damping_add,\n"
        "            (fn)tfx_cross_gradient_calculate, (fn)tfx_calculate_depth_weight, (fn)tfx_calculate_data,\n"
        "            (fn)tfx_sensit_repartition_into, (fn)tfx_model_update};\n"
        '  printf("%d %d\\n", tfx_version(), (int)(sizeof(p) / sizeof(p[0])));\n'
        "  return 0;\n}\n")
    exe = tmp_path / "host"
    libdir = os.path.dirname(tfx.LIB_PATH)
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
                           str(src), "-o", str(exe), "-L", libdir, "-ltfx", "-Wl,-rpath," + libdir,
                           "-L/usr/local/cuda/lib64", "-Wl,-rpath,/usr/local/cuda/lib64", "-lcudart"])
    out = subprocess.check_output([str(exe)], text=True).split()
    assert out == ["100", "8"]


def test_library_is_sm100a_only():
    out = subprocess.check_output(["cuobjdump", "-lelf", tfx.LIB_PATH], text=True)
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_builder_semantics_without_gpu(libtfx):
    # The host-side builder mirrors sparse_matrix.f90 and needs no device until finalize().
    m = tfx.SparseMatrix(3, 4, 5)
    m.add(0.0, 1)                      # zero values are not stored (sparse_matrix.f90:219)
    assert m.get_number_elements() == 0
    m.add(1.5, 2)
    m.new_row()
    m.new_row()                        # empty row: counted, not stored
    m.add_row([1.0, 2.0], [1, 4])
    m.new_row()
    assert m.get_current_row_number() == 3 and m.get_number_elements() == 3
    with pytest.raises(tfx.TfxError):  # nnz bound (:222-223)
        for _ in range(5):
            m.add(1.0, 1)


def test_compute_fails_loudly_without_gpu(libtfx):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    s = np.zeros(8)
    with pytest.raises(tfx.TfxError, match="no CUDA device"):
        tfx.forward_wavelet(s, 2, 2, 2, 1)


def test_partition_helpers_match_reference():
    # parallel_tools.f90:46-63 -- remainder goes to the first ranks
    assert [tfx.calculate_nelements_at_cpu(10, r, 3) for r in range(3)] == [4, 3, 3]
    assert [tfx.get_nsmaller(10, r, 3) for r in range(3)] == [0, 4, 7]


def test_normalize_columns_reference_unit_test(libtfx, oracle):
    # tests_sparse_matrix.f90:39-113 on the host-side builder (no device needed before finalize()):
    # half of the columns are zero; the returned norms equal the dense column norms, the scaled columns
    # have unit (or zero) length; bit-identical to the oracle's restatement.
    ncolumns, nrows = 10, 30
    A = np.zeros((nrows, ncolumns))
    m = tfx.SparseMatrix(nrows, ncolumns, ncolumns * nrows)
    mo = oracle.SparseMatrix(nrows, ncolumns, ncolumns * nrows)
    counter = 0
    for j in range(nrows):
        for i in range(ncolumns):
            counter += 1
            A[j, i] = float(counter) if (i + 1) <= ncolumns // 2 else 0.0
            m.add(A[j, i], i + 1); mo.add(A[j, i], i + 1)
        m.new_row(); mo.new_row()
    mo.finalize()
    cn = m.normalize_columns()
    cn_o = mo.normalize_columns()
    assert np.array_equal(cn, cn_o)
    for i in range(ncolumns):
        assert comparable(cn[i], np.linalg.norm(A[:, i]), TOL)
