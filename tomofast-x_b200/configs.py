"""Host orchestration of the BASELINE.json configurations around the hot path (product side: libtfx only).

The reference's driver (src/problem_joint_gravmag.F90:150-547, src/inversion/joint_inverse_problem.F90:393-573) stays
Fortran in a real deployment; these few lines of Python stand in for it so that bench.py and the tests can run the
named configurations end to end through the C ABI:

  * config D -- parfiles/Parfile_2body_induced.txt: magnetic, 3 model components, 67 x 67 x 30 padded grid, 1681
    stations, distance weighting (type 2, power 3, beta 1.5), compression rate 0.3 (BASELINE asks for Daubechies-4:
    forward.matrixCompression.type = 2), start model 1e-3, damping 1e-8 on every component, 2 major x 100 LSQR iterations.
  * config E -- joint gravity + magnetic inversion with the cross-gradient constraint on a shared grid
    (joint_inverse_problem.F90:436-470,529-533,578-608): WAVELET_DOMAIN = .false., so every LSQR iteration transforms
    the model-sized vectors of both problems forward and back (lsqr_solver2.F90:200-207,228-235).

The input fixture of config D (tests/golden/twobody_induced.npz) is DATA generated from the reference's own input files
by tests/golden/make_2body_fixture.py; this module never imports the oracle.
"""
import time

import numpy as np


# ---------------------------------------------------------------------------------------------------------------------
# config D
# ---------------------------------------------------------------------------------------------------------------------
def load_twobody(path, station_stride=1):
    """Grid, stations, synthetic model and field parameters of Parfile_2body_induced from the committed fixture.
    station_stride > 1 keeps every n-th station of the 41 x 41 lattice in both directions (parity tests)."""
    z = np.load(path)
    nx, ny, nz = int(z["nx"]), int(z["ny"]), int(z["nz"])
    k, j, i = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    i, j, k = i.ravel(), j.ravel(), k.ravel()
    xn, yn, zn = z["xn"], z["yn"], z["zn"]
    grid = [np.ascontiguousarray(a) for a in (xn[i], xn[i + 1], yn[j], yn[j + 1], zn[k], zn[k + 1])]
    ns = int(z["nstations_side"])
    sel = np.arange(0, ns, station_stride)
    sy, sx = np.meshgrid(float(z["station_y0"]) + float(z["station_dy"]) * sel,
                         float(z["station_x0"]) + float(z["station_dx"]) * sel, indexing="ij")
    N = nx * ny * nz
    model = np.tile(np.asarray(z["model_background"], dtype=np.float64).reshape(-1, 1), (1, N))     # (ncomp, N)
    model[:, z["model_cells"]] = np.asarray(z["model_values"], dtype=np.float64).T
    dwt = z["depth_weighting"]
    return dict(nx=nx, ny=ny, nz=nz, N=N, grid=grid, data_xyz=(sx.ravel().copy(), sy.ravel().copy(),
                                                                np.full(sx.size, float(z["station_z"]))),
                ndata=int(sx.size), ncomp=int(model.shape[0]), m_true=model, mi=float(z["inclination"]),
                md=float(z["declination"]), theta=float(z["xaxis_declination"]), intensity=float(z["intensity_nT"]),
                rate=float(z["compression_rate"]), dw_type=int(dwt[0]), dw_power=float(dwt[1]), dw_beta=float(dwt[2]),
                start_value=1.0e-3, alpha=1.0e-8, problem_weight=1.0, nmajor=2, niter=100, rmin=1.0e-13)


def config_d_params(tfx, c, compression_type, cell0, ncl):
    par = tfx.SensitParams()
    par.problem_type = 2
    par.nx, par.ny, par.nz = c["nx"], c["ny"], c["nz"]
    par.ndata, par.ndata_components, par.nmodel_components, par.data_type = c["ndata"], 1, c["ncomp"], 1
    par.compression_type, par.compression_rate = compression_type, c["rate"]
    par.problem_weight = c["problem_weight"]
    par.mi, par.md, par.theta, par.intensity = c["mi"], c["md"], c["theta"], c["intensity"]
    par.cell0, par.ncells_local = cell0, ncl
    par.param_shift = c["ncomp"] * ncl                      # problem 2 of the joint column space (:685-686)
    par.ncolumns = 2 * c["ncomp"] * ncl
    return par


def run_config_d(tfx, c, compression_type=2, rank=0, world=1, S=None, column_weight=None, sync=None):
    """Parfile_2body_induced end to end on the device(s). Returns a dict with the timings, the per-iteration residual
    histories of every major iteration, the costs and this rank's slab of the final model.

    S / column_weight: a prebuilt sensitivity matrix and full column weight (tests hand the same matrix to the checker)."""
    N, nd, ncomp, pw = c["N"], c["ndata"], c["ncomp"], c["problem_weight"]
    nx, ny, nz = c["nx"], c["ny"], c["nz"]
    out = {}
    wall = time.perf_counter
    barrier = sync if sync is not None else (lambda: None)

    # (III) depth weight: distance weighting, every rank the full array (the row pipeline weighs whole kernel lines)
    tfx.synchronize(); barrier(); t0 = wall()
    if column_weight is None:
        column_weight = tfx.calculate_depth_weight(c["dw_type"], c["grid"], c["data_xyz"], c["dw_power"], c["dw_beta"], 0.0)
        column_weight = 1.0 * column_weight                 # column_weight_multiplier(2) = 1 (parameters_init.f90:346)
    tfx.synchronize(); out["depth_weight_s"] = wall() - t0

    dw = np.ones((nd, 1))                                   # data%weight = 1 (data_gravmag.f90:95)
    barrier(); t0 = wall()
    if S is not None:
        ncl, cell0 = N, 0
        out["nnz"] = int(S.get_number_elements())
    elif world == 1:
        ncl, cell0 = N, 0
        S, _, cerr, tot = tfx.calculate_sensit(config_d_params(tfx, c, compression_type, 0, N), c["grid"], c["data_xyz"],
                                               column_weight, dw)
        out["nnz"], out["compression_error"] = int(tot), cerr
    else:
        par = config_d_params(tfx, c, compression_type, 0, N)
        rows, nnz_col, cerr, tot = tfx.sensit_assemble_rows(par, c["grid"], c["data_xyz"], column_weight, dw, rank, world)
        nnz_at, nel_at = tfx.get_load_balancing_nelements(nnz_col, world)
        S = tfx.sensit_repartition(rows, 2, nel_at, rank, world)
        ncl, cell0 = int(nel_at[rank]), int(nel_at[:rank].sum())
        out["nnz"], out["compression_error"], out["column_slabs"] = int(tot), cerr, [int(v) for v in nel_at]
        del rows
    tfx.synchronize(); barrier(); out["assemble_s"] = wall() - t0
    sl = slice(cell0, cell0 + ncl)
    cw = np.ascontiguousarray(column_weight[sl])
    shift = ncomp * ncl
    ncol = 2 * ncomp * ncl
    nlines = nd + ncomp * N

    def calc(model):                                        # t_model%calculate_data (model.F90:220-307)
        return tfx.calculate_data(S, model, nd, 1, pw, cw, dw, compression_type, nx, ny, nz, 1, shift, rank, world).ravel()

    m_true = np.ascontiguousarray(c["m_true"][:, sl])
    m = np.full((ncomp, ncl), c["start_value"])             # starting model in ALL components (model_IO.F90:63-65)
    prior = np.zeros((ncomp, ncl))
    d_obs = calc(m_true)                                    # useSyntheticModelForDataValues = 1
    d_calc = calc(m)
    cost = lambda: float(np.linalg.norm(d_calc - d_obs) / np.linalg.norm(d_obs))
    out["costs"] = [cost()]
    out["histories"], out["rhs_norms"] = [], []
    C = tfx.SparseMatrix(ncomp * N, ncol, ncomp * ncl)
    x = np.zeros(ncol)
    loop_ms = 0.0
    iters = 0
    barrier(); t0 = wall()
    for _ in range(c["nmajor"]):
        b = np.zeros(nlines)
        b[:nd] = pw * (dw.ravel() * (d_obs - d_calc))       # calculate_residuals + calculate_b_RHS
        C.reset()
        cons = b[nd:]
        for k in range(ncomp):                              # joint_inverse_problem.F90:449-463
            tfx.damping_add(C, cons, c["alpha"], pw, 2.0, compression_type, nx, ny, nz, cw, m[k], prior[k], shift + k * ncl,
                            True, None, rank, world)
        C.finalize()
        out["rhs_norms"].append(float(np.linalg.norm(b)))
        tfx.lsqr_solve_sensit(nlines, ncol, c["niter"], c["rmin"], 0.0, 0.0, S, C, b, x, [0, 1], ncl, nx, ny, nz, ncomp,
                              compression_type, True, myrank=rank, nbproc=world)
        h, it, _ = tfx.last_history()
        out["histories"].append(h.copy())
        ms, _, _ = tfx.last_timing()
        loop_ms += ms
        iters += it
        if compression_type > 0:                            # delta_model back from the wavelet domain (:559-567)
            tfx.apply_wavelet_transform(ncl, nx, ny, nz, ncomp, x, False, compression_type, 2, [0, 1], rank, world)
        delta = x[shift:shift + ncomp * ncl].reshape(ncomp, ncl) * cw          # rescale_model (:569-571)
        m = m + delta                                       # model%update
        d_calc = calc(m)
        out["costs"].append(cost())
    tfx.synchronize(); barrier()
    out["inversion_s"] = wall() - t0
    out["iters"], out["loop_ms"] = int(iters), loop_ms
    out["model"], out["cell0"], out["ncl"] = m, cell0, ncl
    out["d_obs"], out["d_calc"] = d_obs, d_calc
    out["S"], out["column_weight"] = S, column_weight
    return out


# ---------------------------------------------------------------------------------------------------------------------
# config E
# ---------------------------------------------------------------------------------------------------------------------
def _block_model(nx, ny, nz, value, ripple):
    """One block in the centre (SURVEY 8d true model) plus a smooth ripple: every cell has non-zero derivatives, so the
    cross-gradient rows are fully populated (no dropped zeros) -- the heaviest constraint block the path can meet."""
    m = np.zeros((nz, ny, nx))
    sl = lambda n: slice(max(0, n // 2 - max(1, n // 8)), n // 2 + max(1, n // 8))
    m[sl(nz), sl(ny), sl(nx)] = value
    x = np.sin(np.arange(nx) * (2.0 * np.pi / 37.0))
    y = np.cos(np.arange(ny) * (2.0 * np.pi / 53.0))
    zz = np.sin(np.arange(nz) * (2.0 * np.pi / 29.0) + 0.3)
    m += ripple * (x[None, None, :] + y[None, :, None] + zz[:, None, None])
    return m.ravel()


def run_config_e(tfx, nx, ny, nz, nd1, nd2, rate=0.002, niter=20, warmup=3, rank=0, world=1, sync=None,
                 alpha=(1.0e-7, 1.0e-4), cross_grad_weight=1.0e-3, compression_type=1):
    """Joint gravity (problem 1, nd1 stations) + magnetic TMI (problem 2, nd2 stations) system on one shared grid with
    model damping on both problems and the cross-gradient coupling, solved in the PHYSICAL domain (WAVELET_DOMAIN = F:
    the constraint rows act on the models, the compressed kernels on their wavelet transforms). Column slabs balanced on
    the summed nnz counts of both kernels (the reference adds the two sensit_nnz files before
    get_load_balancing_nelements). All model- and row-sized vectors stay in HBM."""
    from .synth import depth_weight_type1, regular_grid, station_lattice
    barrier = sync if sync is not None else (lambda: None)
    wall = time.perf_counter
    N = nx * ny * nz
    out = {"nx": nx, "ny": ny, "nz": nz, "ndata": [nd1, nd2]}
    grid = regular_grid(nx, ny, nz)
    lx, ly = 100.0 * nx, 100.0 * ny
    xyz = (station_lattice(nd1, lx, ly, z=-0.1), station_lattice(nd2, lx, ly, z=-5.0))
    cw_full = (depth_weight_type1(grid, 2.0, 0.0, 4.0e3), depth_weight_type1(grid, 3.0, 0.0, 1.0))
    nds = (nd1, nd2)

    def params(i):
        par = tfx.SensitParams()
        par.problem_type = i + 1
        par.nx, par.ny, par.nz = nx, ny, nz
        par.ndata, par.ndata_components, par.nmodel_components, par.data_type = nds[i], 1, 1, 1
        par.compression_type, par.compression_rate = compression_type, rate
        par.problem_weight = 1.0
        par.mi, par.md, par.theta, par.intensity = -60.0, 2.0, 0.0, 55000.0
        par.cell0, par.ncells_local, par.param_shift, par.ncolumns = 0, N, 0, 2 * N
        return par

    # ---- (III) both kernels: rows sharded by station, then ONE nnz-balanced column partition for the joint matrix
    tfx.synchronize(); barrier(); t0 = wall()
    grid = tfx.grid_pin(grid)      # both problems share the grid: one upload
    rows, nnz_col, tot = [], np.zeros(N, dtype=np.int64), []
    for i in range(2):
        r, nc, cerr, t = tfx.sensit_assemble_rows(params(i), grid, xyz[i], cw_full[i], np.ones((nds[i], 1)), rank, world)
        rows.append(r); nnz_col += nc; tot.append(int(t))
    tfx.grid_unpin()
    tfx.synchronize(); barrier(); out["assemble_rows_s"] = wall() - t0
    _, nel_at = tfx.get_load_balancing_nelements(np.minimum(nnz_col, 2**31 - 1).astype(np.int32), world)
    ncl, cell0 = int(nel_at[rank]), int(nel_at[:rank].sum())
    ncol = 2 * ncl
    S = tfx.SparseMatrix(nd1 + nd2, ncol, sum(tot))
    for i in range(2):
        tfx.sensit_repartition_into(S, rows[i], i + 1, nel_at, rank, world)
    S.finalize()
    del rows
    tfx.synchronize(); barrier(); out["assemble_s"] = wall() - t0
    out["nnz"], out["column_slabs"] = int(sum(tot)), [int(v) for v in nel_at]
    del grid

    # ---- current models (full copies: the cross-gradient stencil reaches across slab boundaries), data, right-hand side
    sl = slice(cell0, cell0 + ncl)
    m_full = (_block_model(nx, ny, nz, 250.0, 5.0), _block_model(nx, ny, nz, 0.05, 1.0e-3))
    cw = tuple(np.ascontiguousarray(c[sl]) for c in cw_full)
    nd = nd1 + nd2
    ncons = 2 * N + 3 * N
    nlines = nd + ncons
    rhs = tfx.zero(tfx.Buffer(nlines))
    b_data = np.zeros(nd)
    line0 = (1, nd1 + 1)
    for i in range(2):
        dcalc = tfx.calculate_data(S, np.ascontiguousarray(m_full[i][sl]), nds[i], 1, 1.0, cw[i], np.ones((nds[i], 1)),
                                   compression_type, nx, ny, nz, line0[i], i * ncl, rank, world).ravel()
        b_data[line0[i] - 1:line0[i] - 1 + nds[i]] = 0.1 * dcalc          # residual of a model 10 % off
    tfx.copy(rhs, b_data, nd)
    mdev = []
    for i in range(2):
        b = tfx.Buffer(N); tfx.copy(b, m_full[i], N); mdev.append(b)
    prior = tfx.zero(tfx.Buffer(ncl))
    dX, dY, dZ = np.full(nx, 100.0), np.full(ny, 100.0), np.full(nz, 50.0)

    tfx.synchronize(); barrier(); t0 = wall()
    C = tfx.SparseMatrix(ncons, ncol, 2 * ncl + 3 * 12 * ncl)
    cons = tfx.BufferView(rhs, nd, ncons)
    for i in range(2):                                                   # joint_inverse_problem.F90:449-463
        tfx.damping_add(C, cons, alpha[i], 1.0, 2.0, compression_type, nx, ny, nz, cw[i], tfx.BufferView(mdev[i], cell0, ncl),
                        prior, i * ncl, False, None, rank, world)
    cost, _ = tfx.cross_gradient_calculate(C, cons, nx, ny, nz, dX, dY, dZ, mdev[0], mdev[1], cw[0], cw[1], 1,
                                           cross_grad_weight, (0, 0), rank, world, want_cross_grad=False)
    C.finalize()
    tfx.synchronize(); barrier(); out["constraints_s"] = wall() - t0
    out["constraint_rows"], out["constraint_nnz_local"] = ncons, int(C.get_number_elements())
    out["cross_gradient_cost"] = [float(v) for v in cost]
    del mdev

    # ---- the solve: warm-up, then the timed iterations (device-resident u and x)
    x = tfx.Buffer(ncol)
    u = tfx.Buffer(nlines)
    res = {}
    for name, n in (("warmup", warmup), ("timed", niter)):
        tfx.copy(u, rhs, nlines)
        barrier()
        tfx.lsqr_solve_sensit(nlines, ncol, n, 1.0e-13, 0.0, 0.0, S, C, u, x, [1, 1], ncl, nx, ny, nz, 1, compression_type,
                              False, myrank=rank, nbproc=world)
        ms, _, _ = tfx.last_timing()
        h, it, _ = tfx.last_history()
        res[name] = (ms, it, h.copy())
    out["loop_ms"], out["iters"], out["history"] = res["timed"]
    # ---- where the time goes: the two products with S and one distributed wavelet transform, timed on their own
    xs, us, qs, ts = tfx.zero(tfx.Buffer(ncol)), tfx.zero(tfx.Buffer(nd)), tfx.Buffer(nd), tfx.Buffer(ncol)
    barrier(); out["S_fwd_ms"] = S.time_product(0, xs, qs, 5)
    barrier(); out["S_trans_ms"] = S.time_product(1, us, ts, 5)
    vol = tfx.zero(tfx.Buffer(ncl))
    tfx.apply_wavelet_transform(ncl, nx, ny, nz, 1, vol, True, compression_type, 1, [1], rank, world)
    tfx.synchronize(); barrier(); t0 = wall()
    for _ in range(4):
        tfx.apply_wavelet_transform(ncl, nx, ny, nz, 1, vol, True, compression_type, 1, [1], rank, world)
    tfx.synchronize(); barrier(); out["wavelet_slab_ms"] = (wall() - t0) / 4 * 1e3
    mode = int(tfx.lib().tfx_wavelet_last_distributed())
    out["wavelet_distributed"] = bool(mode)
    out["wavelet_exchange"] = {0: "all-gather (fallback)", 1: "ncclSend/ncclRecv", 2: "peer memory (cudaIpc)"}.get(mode, str(mode))
    out["ncl"], out["cell0"], out["nlines"], out["ncolumns_local"] = ncl, cell0, nlines, ncol
    return out
