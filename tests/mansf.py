"""Config A (parfiles/Parfile_mansf_slice.txt): gravity inversion with global ADMM bounds, Haar
compression 0.15, 60 major x 100 LSQR iterations -- the host orchestration around the hot path.

Mirrors, with defaults resolved:
  problem_joint_gravmag.F90:172-201 (depth weight type 1 * 4e3, assembly), :331-358 (observed data from the
  synthetic model), :413-441 (prior/start = 0), :473-547 (major loop);
  joint_inverse_problem.F90:393-573 (RHS, ADMM block through damping%add, lsqr_solve_sensit,
  inverse wavelet, rescale); damping.F90:97-234; admm_method.F90:70-134; model.F90:220-307.

`be` is a backend object exposing the SAME calls for the oracle (CPU checker) and for the product
(libtfx through the C ABI), so the parity tests drive both with identical host code.
"""
import os

import numpy as np

from tests.synth import depth_weight_type1, regular_grid

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mansf_slice.npz")


class Config:
    def __init__(self):
        z = np.load(GOLDEN)
        self.nx, self.ny, self.nz = int(z["nx"]), int(z["ny"]), int(z["nz"])
        self.N = self.nx * self.ny * self.nz
        self.grid = regular_grid(self.nx, self.ny, self.nz, float(z["dx"]), float(z["dy"]), float(z["dz"]),
                                 x0=float(z["x0"]))
        ys, xs = np.meshgrid(float(z["station_y0"]) + float(z["station_dy"]) * np.arange(128), z["station_x"],
                             indexing="ij")
        self.data_xyz = (xs.ravel().copy(), ys.ravel().copy(), np.full(256, float(z["station_z"])))
        self.ndata = 256
        self.m_true = z["model"].astype(np.float64)
        b = z["admm_bounds"]
        self.xmin = np.tile(b[0::2], (self.N, 1))
        self.xmax = np.tile(b[1::2], (self.N, 1))
        self.compression_type, self.rate = 1, 0.15
        self.nel_compressed = int(self.rate * self.N)            # 1228
        self.problem_weight = 1.0
        self.rho_admm = 1.0e-5
        self.niter, self.rmin = 100, 1.0e-13
        self.cw = depth_weight_type1(self.grid, 2.0, 0.0, 4.0e3)
        self.dw = np.ones(self.ndata)
        self.ncolumns = 2 * self.N


class OracleBackend:
    name = "oracle"

    def __init__(self, orc):
        self.orc = orc

    def assemble(self, cfg):
        orc = self.orc
        S = orc.SparseMatrix(cfg.ndata, cfg.ncolumns, cfg.nel_compressed * cfg.ndata)
        err = 0.0
        for i in range(cfg.ndata):
            line = orc.graviprism_z(cfg.grid, *(float(a[i]) for a in cfg.data_xyz)) * cfg.cw
            r = orc.compress_row(line, cfg.nx, cfg.ny, cfg.nz, cfg.compression_type, cfg.nel_compressed)
            wgt = np.float32(cfg.problem_weight * cfg.dw[i])
            S.add_row((r["vals"] * wgt).astype(np.float32), r["cols"])
            S.new_row()
            err += np.sqrt(r["cost_discarded"] / r["cost_full"])
        S.finalize()
        return S, err / cfg.ndata

    def fwd(self, v, cfg):
        return self.orc.forward_wavelet(v, cfg.nx, cfg.ny, cfg.nz, cfg.compression_type)

    def inv(self, v, cfg):
        return self.orc.inverse_wavelet(v, cfg.nx, cfg.ny, cfg.nz, cfg.compression_type)

    def part_mult(self, S, x, ndata):
        return S.part_mult_vector(x, ndata, 1, 0)

    def cons_matrix(self, cfg, value):
        C = self.orc.SparseMatrix(cfg.N, cfg.ncolumns, cfg.N)
        for i in range(cfg.N):
            C.add(value, i + 1)
            C.new_row()
        C.finalize()
        return C

    def solve(self, cfg, S, C, b):
        x, hist, it = self.orc.lsqr_solve_sensit(cfg.niter, cfg.rmin, 0.0, 0.0, S, C, b, cfg.N, cfg.nx, cfg.ny,
                                                 cfg.nz, 1, cfg.compression_type, True)
        return x, hist


class TfxBackend:
    name = "tfx"

    def __init__(self, tfx):
        self.tfx = tfx

    def assemble(self, cfg):
        tfx = self.tfx
        par = tfx.SensitParams()
        par.problem_type = 1
        par.nx, par.ny, par.nz = cfg.nx, cfg.ny, cfg.nz
        par.ndata, par.ndata_components, par.nmodel_components, par.data_type = cfg.ndata, 1, 1, 1
        par.compression_type, par.compression_rate = cfg.compression_type, cfg.rate
        par.problem_weight = cfg.problem_weight
        par.cell0, par.ncells_local, par.param_shift, par.ncolumns = 0, cfg.N, 0, cfg.ncolumns
        S, nnz_col, cerr, tot = tfx.calculate_sensit(par, cfg.grid, cfg.data_xyz, cfg.cw, cfg.dw.reshape(-1, 1))
        self.nnz_total = tot
        return S, cerr

    def fwd(self, v, cfg):
        return self.tfx.forward_wavelet(np.array(v, dtype=np.float64), cfg.nx, cfg.ny, cfg.nz, cfg.compression_type)

    def inv(self, v, cfg):
        return self.tfx.inverse_wavelet(np.array(v, dtype=np.float64), cfg.nx, cfg.ny, cfg.nz, cfg.compression_type)

    def part_mult(self, S, x, ndata):
        return S.part_mult_vector(x, ndata, 1, 0)

    def cons_matrix(self, cfg, value):
        sa = np.full(cfg.N, value, dtype=np.float32)
        ija = np.arange(1, cfg.N + 1, dtype=np.int32)
        ijl = np.arange(1, cfg.N + 2, dtype=np.int64)
        rowptr = np.arange(1, cfg.N + 1, dtype=np.int32)
        return self.tfx.SparseMatrix.from_arrays(cfg.N, cfg.ncolumns, sa, ija, ijl, rowptr)

    def solve(self, cfg, S, C, b):
        u = b.copy()
        x = np.zeros(cfg.ncolumns)
        self.tfx.lsqr_solve_sensit(len(u), cfg.ncolumns, cfg.niter, cfg.rmin, 0.0, 0.0, S, C, u, x, [1, 0], cfg.N,
                                   cfg.nx, cfg.ny, cfg.nz, 1, cfg.compression_type, True)
        hist, it, fused = self.tfx.last_history()
        return x, hist


def calculate_data(be, cfg, S, m):
    """t_model%calculate_data, model.F90:220-307 (nbproc = 1)."""
    ms = np.where(cfg.cw != 0.0, m / cfg.cw, 0.0)
    ms = be.fwd(ms, cfg)
    d = be.part_mult(S, ms, cfg.ndata)
    return d / cfg.problem_weight / cfg.dw


class Inversion:
    """State of the major loop; step() = one major iteration (jinv%solve + model update + new data)."""

    def __init__(self, be, cfg, admm_iterate):
        self.be, self.cfg = be, cfg
        self.admm_iterate = admm_iterate
        self.S, self.comp_error = be.assemble(cfg)
        self.d_obs = calculate_data(be, cfg, self.S, cfg.m_true)      # observed = forward of the true model
        self.m = np.zeros(cfg.N)                                      # starting model 0, prior 0
        self.d_calc = calculate_data(be, cfg, self.S, self.m)
        self.z = np.zeros(cfg.N)
        self.u_admm = np.zeros(cfg.N)
        self.C = be.cons_matrix(cfg, cfg.rho_admm * cfg.problem_weight * 1.0)   # damping.F90:161-176
        self.costs = [self.cost()]
        self.histories = []

    def cost(self):
        return float(np.linalg.norm(self.d_calc - self.d_obs) / np.linalg.norm(self.d_obs))   # data_gravmag.f90:123

    def build_rhs(self):
        cfg = self.cfg
        res = cfg.dw * (self.d_obs - self.d_calc)                                  # calculate_residuals
        b = np.zeros(cfg.ndata + cfg.N)
        b[:cfg.ndata] = cfg.problem_weight * res                                   # calculate_b_RHS
        x0 = self.admm_iterate(cfg.xmin, cfg.xmax, self.m, self.z, self.u_admm)    # iterate_admm_arrays
        diff = (self.m - x0) / cfg.cw                                              # damping.F90:124-135
        diff = self.be.fwd(diff, cfg)                                              # WAVELET_DOMAIN (:137-149)
        b[cfg.ndata:] = -cfg.rho_admm * cfg.problem_weight * diff * 1.0            # add_RHS (:217-229)
        return b

    def apply(self, x):
        cfg = self.cfg
        delta = self.be.inv(x[:cfg.N].copy(), cfg)                                 # jip.F90:559-567
        delta = delta * cfg.cw                                                     # rescale_model (:570)
        self.m = self.m + delta                                                    # model%update
        self.d_calc = calculate_data(self.be, cfg, self.S, self.m)
        self.costs.append(self.cost())

    def step(self):
        b = self.build_rhs()
        x, hist = self.be.solve(self.cfg, self.S, self.C, b)
        self.histories.append(hist)
        self.apply(x)
        return b, x, hist


class DeviceInversion:
    """The same major loop with every model-sized vector resident in HBM and every step through the device entry
    points: tfx_calculate_data, tfx_admm_iterate_admm_arrays, tfx_damping_add (matrix_cons is
    reset and rebuilt on the device every major iteration, joint_inverse_problem.F90:364-373,497-527),
    tfx_lsqr_solve_sensit, tfx_apply_wavelet_transform, tfx_rescale_model, tfx_model_update. Only the ndata-sized data
    vectors visit the host (residuals / costs, data_gravmag.f90:123-150)."""

    def __init__(self, tfx, cfg):
        self.tfx, self.cfg = tfx, cfg
        be = TfxBackend(tfx)
        self.S, self.comp_error = be.assemble(cfg)
        N, nd = cfg.N, cfg.ndata
        up = lambda a: self._upload(np.ascontiguousarray(a, dtype=np.float64))
        self.cw, self.xmin, self.xmax = up(cfg.cw), up(cfg.xmin), up(cfg.xmax)
        self.m = up(np.zeros(N)); self.z = up(np.zeros(N)); self.u_admm = up(np.zeros(N))
        self.dw2 = cfg.dw.reshape(nd, 1)
        self.d_obs = self.calc(up(cfg.m_true))
        self.d_calc = self.calc(self.m)
        self.C = tfx.SparseMatrix(N, cfg.ncolumns, N)
        self.rhs = tfx.Buffer(nd + N)
        self.x = tfx.Buffer(cfg.ncolumns)
        self.costs = [self.cost()]
        self.histories = []
        self.admm_costs = []

    def _upload(self, a):
        b = self.tfx.Buffer(a.size)
        self.tfx.copy(b, a, a.size)
        return b

    def calc(self, model_buf):
        cfg = self.cfg
        return self.tfx.calculate_data(self.S, model_buf, cfg.ndata, 1, cfg.problem_weight, self.cw, self.dw2,
                                       cfg.compression_type, cfg.nx, cfg.ny, cfg.nz).ravel()

    def cost(self):
        return float(np.linalg.norm(self.d_calc - self.d_obs) / np.linalg.norm(self.d_obs))

    def step(self):
        tfx, cfg = self.tfx, self.cfg
        N, nd = cfg.N, cfg.ndata
        res = cfg.dw * (self.d_obs - self.d_calc)
        tfx.copy(self.rhs, np.ascontiguousarray(cfg.problem_weight * res), nd)         # calculate_b_RHS
        x0 = tfx.admm_iterate_admm_arrays(self.xmin, self.xmax, self.m, self.z, self.u_admm)
        self.C.reset()
        cons = tfx.BufferView(self.rhs, nd, N)
        cost = tfx.damping_add(self.C, cons, cfg.rho_admm, cfg.problem_weight, 2.0, cfg.compression_type, cfg.nx, cfg.ny,
                               cfg.nz, self.cw, self.m, x0, 0, True)
        self.C.finalize()
        self.admm_costs.append(cost)
        b_host = self.rhs.numpy().copy()
        tfx.lsqr_solve_sensit(nd + N, cfg.ncolumns, cfg.niter, cfg.rmin, 0.0, 0.0, self.S, self.C, self.rhs, self.x, [1, 0], N,
                              cfg.nx, cfg.ny, cfg.nz, 1, cfg.compression_type, True)
        hist, it, fused = tfx.last_history()
        self.histories.append(hist)
        x_host = self.x.numpy().copy()
        delta = tfx.BufferView(self.x, 0, N)
        tfx.apply_wavelet_transform(N, cfg.nx, cfg.ny, cfg.nz, 1, delta, False, cfg.compression_type, 1, [1], 0, 1)
        tfx.rescale_model(delta, self.cw)
        tfx.model_update(self.m, delta)
        self.d_calc = self.calc(self.m)
        self.costs.append(self.cost())
        return b_host, x_host, hist
