"""Config D (Parfile_2body_induced) on the device: the magnetic kernel and the D4-compressed rows of stations 1 and 841
on the reference's real 67 x 67 x 30 padded grid (tests/golden/twobody_induced.npz) against the oracle. The padded grid is
a non-uniform tensor product, i.e. the shared corner / edge path of csrc/assembly.cu with node spacings from 50 m to
several hundred metres."""
import numpy as np
import pytest

import tomofastx_b200 as tfx
from tests.test_oracle_2body import cfg  # noqa: F401  (fixture)

pytestmark = pytest.mark.gpu


def _par(c, ndata, ctype):
    par = tfx.SensitParams()
    par.problem_type = 2
    par.nx, par.ny, par.nz = c["nx"], c["ny"], c["nz"]
    par.ndata, par.ndata_components, par.nmodel_components, par.data_type = ndata, 1, 3, 1
    par.compression_type, par.compression_rate = ctype, c["rate"]
    par.problem_weight = 1.0
    par.mi, par.md, par.theta, par.intensity = c["mi"], c["md"], c["theta"], c["intensity"]
    par.cell0, par.ncells_local, par.param_shift, par.ncolumns = 0, c["N"], 0, 2 * 3 * c["N"]
    return par


def test_config_d_lines_and_compressed_rows(oracle, cfg):  # noqa: F811
    c = cfg
    st = np.array([0, 840])
    xyz = (c["sx"][st].copy(), c["sy"][st].copy(), np.full(2, c["sz"]))
    lines = tfx.sensit_lines(_par(c, 2, 0), c["grid"], xyz)                    # (station, data comp, model comp, cell)
    want = [oracle.magprism(c["grid"], float(xyz[0][i]), float(xyz[1][i]), c["sz"], 3, 1, c["mi"], c["md"], c["theta"],
                            c["intensity"]) for i in range(2)]
    for i in range(2):
        assert np.abs(lines[i] - want[i]).max() / np.abs(want[i]).max() < 1e-11

    # D4-compressed rows (rate 0.3 -> 40 401 entries per component), unit column weight like the surveyor's figures
    N, nel = c["N"], int(c["rate"] * c["N"])
    rows, nnz_col, cerr, tot = tfx.sensit_assemble_rows(_par(c, 2, 2), c["grid"], xyz, np.ones(N), np.ones((2, 1)))
    assert tot == 2 * 3 * nel and nnz_col.sum() == tot
    S = tfx.sensit_repartition(rows, 1, [N])
    sa, ija, ijl, rowptr = S.export()
    assert np.array_equal(rowptr, [1, 2]) and np.array_equal(np.diff(ijl), [3 * nel, 3 * nel])
    errs = []
    flips = 0
    for i in range(2):
        got_cols, got_vals = ija[ijl[i] - 1:ijl[i + 1] - 1], sa[ijl[i] - 1:ijl[i + 1] - 1]
        for k in range(3):
            r = oracle.compress_row(want[i][0, k].copy(), c["nx"], c["ny"], c["nz"], 2, nel)
            seg = slice(k * nel, (k + 1) * nel)
            cols_k = got_cols[seg] - k * N                                      # column = p + (k-1)*nelements (:834)
            assert cols_k.min() >= 1 and cols_k.max() <= N
            diff = set(cols_k) ^ set(r["cols"])
            flips += len(diff)
            if not diff:
                assert np.allclose(got_vals[seg], r["vals"], rtol=3e-6, atol=1e-6 * np.abs(r["vals"]).max())
            errs.append(np.sqrt(r["cost_discarded"] / r["cost_full"]))
    assert flips <= 12                                                          # threshold ties within rounding
    assert cerr == pytest.approx(np.mean(errs), rel=1e-3)                       # comp_error (:346-353)
