"""Synthetic gravity / magnetic problems of the shape SURVEY.md section 8(d) prescribes.

Grid: regular boxes 100 x 100 x 50 m, i fastest then j then k (src/inversion/model_IO.F90:184-193);
stations on a lattice over the grid footprint at z = -0.1 (like data/gravmag/mansf_slice/data_grid.txt),
never on a cell face; depth weighting type 1 (src/forward/gravmag/weights_gravmag.f90:71-79,170-195) with
the default gravity column-weight multiplier 4e3 (src/parameters_init.f90:345); true model = one block.

The helpers that need the CPU oracle take it as an argument: this module never imports it, so it is
shared by the tests (checker = oracle) and by bench.py / smoke() (product path only).
"""
import numpy as np



def regular_grid(nx, ny, nz, dx=100.0, dy=100.0, dz=50.0, x0=0.0, y0=0.0):
    i = np.arange(nx, dtype=np.float64)
    j = np.arange(ny, dtype=np.float64)
    k = np.arange(nz, dtype=np.float64)
    K, J, I = np.meshgrid(k, j, i, indexing="ij")          # flattened C order == i fastest
    X1 = (x0 + dx * I).ravel(); X2 = X1 + dx
    Y1 = (y0 + dy * J).ravel(); Y2 = Y1 + dy
    Z1 = (dz * K).ravel(); Z2 = Z1 + dz
    return [np.ascontiguousarray(a) for a in (X1, X2, Y1, Y2, Z1, Z2)]


def station_lattice(ndata, lx, ly, z=-0.1, x0=0.0, y0=0.0):
    nsx = int(np.ceil(np.sqrt(ndata)))
    nsy = int(np.ceil(ndata / nsx))
    ix = np.arange(nsx, dtype=np.float64)
    iy = np.arange(nsy, dtype=np.float64)
    # irrational-ish offsets keep every station off the cell faces (cf. gravity_field.f90:176-181)
    xs = x0 + (ix + 0.5) * lx / nsx + 0.3719
    ys = y0 + (iy + 0.5) * ly / nsy + 0.2113
    YY, XX = np.meshgrid(ys, xs, indexing="ij")
    x = XX.ravel()[:ndata].copy()
    y = YY.ravel()[:ndata].copy()
    return x, y, np.full(ndata, z)


def depth_weight_type1(grid, power, Z0=0.0, multiplier=1.0):
    """calculate_depth_weight type 1 + volume scaling + normalisation + inversion
    (weights_gravmag.f90:71-79,170-195), then the column-weight multiplier
    (src/problem_joint_gravmag.F90:178)."""
    X1, X2, Y1, Y2, Z1, Z2 = grid
    depth = 0.5 * (Z1 + Z2)
    w = (depth + Z0) ** (-power / 2.0)
    w = w * np.sqrt(np.abs((X2 - X1) * (Y2 - Y1) * (Z2 - Z1)))
    w = w / w.max()
    return multiplier * (1.0 / w)


class Problem:
    pass


def make_problem(nx, ny, nz, ndata, compression_type=0, rate=0.1, problem_type=1, nmodel_components=1,
                 ndata_components=1, problem_weight=1.0, seed=None):
    pb = Problem()
    pb.nx, pb.ny, pb.nz, pb.ndata = nx, ny, nz, ndata
    pb.N = nx * ny * nz
    pb.grid = regular_grid(nx, ny, nz)
    pb.data_xyz = station_lattice(ndata, 100.0 * nx, 100.0 * ny, z=-0.1 if problem_type == 1 else -5.0)
    if problem_type == 1:
        pb.cw = depth_weight_type1(pb.grid, 2.0, 0.0, 4.0e3)
    else:
        pb.cw = depth_weight_type1(pb.grid, 3.0, 0.0, 1.0)
    pb.dw = np.ones((ndata, ndata_components))              # data%weight defaults to 1 (data_gravmag.f90:95)
    pb.ncomp = nmodel_components
    pb.ndc = ndata_components
    pb.ncolumns = 2 * nmodel_components * pb.N              # joint_inverse_problem.F90:213-214
    from . import SensitParams
    par = SensitParams()
    par.problem_type = problem_type
    par.nx, par.ny, par.nz = nx, ny, nz
    par.ndata = ndata
    par.ndata_components = ndata_components
    par.nmodel_components = nmodel_components
    par.data_type = 1
    par.compression_type = compression_type
    par.compression_rate = rate
    par.problem_weight = problem_weight
    par.mi, par.md, par.theta, par.intensity = -60.0, 2.0, 0.0, 55000.0   # Parfile_2body_induced
    par.cell0, par.ncells_local = 0, pb.N
    par.param_shift = 0 if problem_type == 1 else nmodel_components * pb.N   # sensitivity_gravmag.F90:685-686
    par.ncolumns = pb.ncolumns
    pb.par = par
    pb.nel_compressed = int(rate * pb.N) if compression_type > 0 else pb.N

    # true model: one block in the centre (250 kg/m3 gravity, 0.05 SI magnetic)
    m = np.zeros((nmodel_components, nz, ny, nx))
    sl = lambda n: slice(max(0, n // 2 - max(1, n // 8)), n // 2 + max(1, n // 8))
    m[:, sl(nz), sl(ny), sl(nx)] = 250.0 if problem_type == 1 else 0.05
    pb.m_true = m.reshape(nmodel_components, pb.N)

    def oracle_matrix(orc, pb=pb):
        """CPU restatement of calculate_and_write_sensit + read_sensitivity_kernel (nbproc = 1)."""
        p = pb.par
        nl = p.ndata * p.ndata_components
        S = orc.SparseMatrix(nl, p.ncolumns, pb.nel_compressed * nl * p.nmodel_components)
        for i in range(p.ndata):
            xd, yd, zd = (float(a[i]) for a in pb.data_xyz)
            if p.problem_type == 1:
                lines = orc.graviprism_z(pb.grid, xd, yd, zd).reshape(1, 1, pb.N)
            else:
                lines = orc.magprism(pb.grid, xd, yd, zd, p.nmodel_components, p.ndata_components,
                                     p.mi, p.md, p.theta, p.intensity)
            for d in range(p.ndata_components):
                for k in range(p.nmodel_components):
                    line = lines[d, k] * pb.cw                                   # apply_column_weight
                    r = orc.compress_row(line, p.nx, p.ny, p.nz, p.compression_type, pb.nel_compressed)
                    wgt = np.float32(p.problem_weight * pb.dw[i, d])             # combined_weight, real(4)
                    cols = r["cols"] + (p.param_shift + k * pb.N)
                    S.add_row((r["vals"] * wgt).astype(np.float32), cols)
                S.new_row()
        S.finalize()
        return S

    def model_scaled_w(orc, pb=pb):
        """m/column_weight, wavelet transformed when compression is on (model.F90:243-283), laid out in
        the solver's column space."""
        x = np.zeros(pb.ncolumns)
        for k in range(pb.ncomp):
            v = pb.m_true[k] / pb.cw
            if pb.par.compression_type > 0:
                v = orc.forward_wavelet(v, pb.nx, pb.ny, pb.nz, pb.par.compression_type)
            o = pb.par.param_shift + k * pb.N
            x[o:o + pb.N] = v
        return x

    def rhs(S_orc, orc, pb=pb):
        """Right-hand side of the first major iteration: S.W(m_true/cw) (observed data, zero start). `orc` is the
        checker module handed in by the caller (tests / smoke); the package itself never imports it."""
        return S_orc.mult_vector(model_scaled_w(orc))

    pb.oracle_matrix = oracle_matrix
    pb.model_scaled_w = model_scaled_w
    pb.rhs = rhs
    return pb
