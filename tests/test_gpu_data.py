"""t_model%calculate_data on the device (csrc/data.cu) against the oracle restatement of model.F90:220-307."""
import numpy as np
import pytest

import tomofastx_b200 as tfx
from tests.synth import make_problem

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", ["grav_none", "grav_haar", "mag3_d4"])
def test_calculate_data_vs_oracle(oracle, case):
    if case == "grav_none":
        pb = make_problem(nx=10, ny=9, nz=5, ndata=12, compression_type=0, problem_weight=0.8)
    elif case == "grav_haar":
        pb = make_problem(nx=12, ny=10, nz=6, ndata=11, compression_type=1, rate=0.25, problem_weight=1.25)
    else:
        pb = make_problem(nx=8, ny=7, nz=4, ndata=6, compression_type=2, rate=0.3, problem_type=2, nmodel_components=3)
    rng = np.random.default_rng(5)
    dw = rng.uniform(0.5, 2.0, (pb.ndata, pb.ndc))
    cw = pb.cw.copy()
    cw[::17] = 0.0                                   # zero column weight -> the cell contributes nothing (model.F90:245-249)
    pb.dw = dw
    S, _, _, _ = tfx.calculate_sensit(pb.par, pb.grid, pb.data_xyz, pb.cw, dw)
    So = pb.oracle_matrix(oracle)
    model = pb.m_true + 0.01 * rng.standard_normal(pb.m_true.shape)
    shift = pb.par.param_shift
    want = oracle.calculate_data(So, model, pb.ndata, pb.ndc, pb.par.problem_weight, cw, dw, pb.par.compression_type,
                                 pb.nx, pb.ny, pb.nz, 1, shift)
    got = tfx.calculate_data(S, model, pb.ndata, pb.ndc, pb.par.problem_weight, cw, dw, pb.par.compression_type,
                             pb.nx, pb.ny, pb.nz, 1, shift)
    # compressed patterns may differ by threshold flips (see test_compressed_assembly_vs_oracle): compare against the
    # oracle loop run on the device's own matrix for the tight bound, and against the oracle matrix loosely
    sa, ija, ijl, rowptr = S.export()
    Sd = oracle.SparseMatrix(pb.ndata * pb.ndc, pb.ncolumns, len(sa))
    k = 0
    for r in range(1, pb.ndata * pb.ndc + 1):
        if k < len(rowptr) and rowptr[k] == r:
            Sd.add_row(sa[ijl[k] - 1:ijl[k + 1] - 1], ija[ijl[k] - 1:ijl[k + 1] - 1]); k += 1
        Sd.new_row()
    Sd.finalize()
    tight = oracle.calculate_data(Sd, model, pb.ndata, pb.ndc, pb.par.problem_weight, cw, dw, pb.par.compression_type,
                                  pb.nx, pb.ny, pb.nz, 1, shift)
    assert np.allclose(got, tight, rtol=1e-12, atol=1e-13 * np.abs(tight).max())
    assert np.allclose(got, want, rtol=1e-5, atol=1e-7 * np.abs(want).max())
    # device-resident model / weights / output
    mb, cb, wb, ob = tfx.Buffer(model.size), tfx.Buffer(cw.size), tfx.Buffer(dw.size), tfx.Buffer(dw.size)
    tfx.copy(mb, np.ascontiguousarray(model), model.size); tfx.copy(cb, cw, cw.size); tfx.copy(wb, np.ascontiguousarray(dw), dw.size)
    tfx.calculate_data(S, mb, pb.ndata, pb.ndc, pb.par.problem_weight, cb, wb, pb.par.compression_type, pb.nx, pb.ny, pb.nz,
                       1, shift, data_calc=ob)
    assert np.array_equal(ob.numpy().reshape(got.shape), got)


def test_calculate_data_zero_problem_weight_aborts():
    pb = make_problem(nx=6, ny=5, nz=4, ndata=5, compression_type=0)
    S, _, _, _ = tfx.calculate_sensit(pb.par, pb.grid, pb.data_xyz, pb.cw, pb.dw)
    with pytest.raises(tfx.TfxError, match="Zero problem weight"):
        tfx.calculate_data(S, pb.m_true, pb.ndata, 1, 0.0, pb.cw, pb.dw, 0, pb.nx, pb.ny, pb.nz)
