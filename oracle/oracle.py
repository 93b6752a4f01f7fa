"""ctypes loader for the CPU ORACLE (test infrastructure, NOT product code).

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference)
may import this module. The product package (tomofast-x_b200/) never does.

Each wrapper mirrors one reference routine; see oracle/tfx_oracle.h for the
file:line citations and the parity-pinning statement.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libtfx_oracle.so")


def build(force=False):
    """Compile oracle/tfx_oracle.c -> libtfx_oracle.so (gcc, a second or two)."""
    src = os.path.join(_HERE, "tfx_oracle.c")
    if (not force and os.path.exists(_LIB_PATH)
            and os.path.getmtime(_LIB_PATH) >= os.path.getmtime(src)
            and os.path.getmtime(_LIB_PATH) >= os.path.getmtime(os.path.join(_HERE, "tfx_oracle.h"))):
        return _LIB_PATH
    subprocess.check_call(["make", "-C", _HERE, "-B", "libtfx_oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None
_d = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_f = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_i = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


class _CsrStruct(C.Structure):
    _fields_ = [("nnz", C.c_int64), ("nel", C.c_int64), ("nel_last", C.c_int64),
                ("nl", C.c_int32), ("nl_nonempty", C.c_int32), ("nl_nonempty_allocated", C.c_int32),
                ("nl_current", C.c_int32), ("nl_current_all", C.c_int32), ("ncolumns", C.c_int32),
                ("sa", C.POINTER(C.c_float)), ("ija", C.POINTER(C.c_int32)),
                ("ijl", C.POINTER(C.c_int64)), ("rowptr", C.POINTER(C.c_int32)),
                ("finalized", C.c_int32)]


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(_LIB_PATH)
    P = C.POINTER(_CsrStruct)
    L.orc_csr_new.restype = P
    L.orc_csr_new.argtypes = [C.c_int32, C.c_int32, C.c_int64, C.c_int32]
    L.orc_csr_free.argtypes = [P]
    L.orc_csr_reset.argtypes = [P]
    L.orc_csr_add.argtypes = [P, C.c_double, C.c_int32]
    L.orc_csr_add_row.argtypes = [P, C.c_int32, _f, _i]
    L.orc_csr_new_row.argtypes = [P]
    L.orc_csr_add_empty_rows.argtypes = [P, C.c_int32]
    L.orc_csr_finalize.argtypes = [P]
    for name in ("orc_csr_mult_vector", "orc_csr_add_mult_vector", "orc_csr_trans_mult_vector",
                 "orc_csr_add_trans_mult_vector"):
        getattr(L, name).argtypes = [P, _d, _d]
        getattr(L, name).restype = None
    L.orc_csr_part_mult_vector.argtypes = [P, C.c_int32, _d, C.c_int32, _d, C.c_int32, C.c_int32]
    L.orc_csr_normalize_columns.argtypes = [P, _d]
    for name in ("orc_haar3d", "orc_ihaar3d", "orc_daubd43d", "orc_idaubd43d"):
        getattr(L, name).argtypes = [_d, C.c_int, C.c_int, C.c_int]
        getattr(L, name).restype = None
    L.orc_forward_wavelet.argtypes = [_d, C.c_int, C.c_int, C.c_int, C.c_int]
    L.orc_inverse_wavelet.argtypes = [_d, C.c_int, C.c_int, C.c_int, C.c_int]
    L.orc_lsqr_solve.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_double, C.c_double, P, _d, _d,
                                 C.c_void_p, C.POINTER(C.c_int32)]
    L.orc_lsqr_solve_sensit.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_double, C.c_double, C.c_double,
                                        P, P, _d, _d, _i, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                        C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.POINTER(C.c_int32)]
    L.orc_graviprism_z.argtypes = [C.c_int32, _d, _d, _d, _d, _d, _d, C.c_double, C.c_double, C.c_double, _d]
    L.orc_gradiprism_zz.argtypes = [C.c_int32, _d, _d, _d, _d, _d, _d, C.c_double, C.c_double, C.c_double, _d]
    L.orc_gradiprism_zz.restype = None
    L.orc_gradiprism_full.argtypes = [C.c_int32, _d, _d, _d, _d, _d, _d, C.c_double, C.c_double, C.c_double, _d]
    L.orc_gradiprism_full.restype = C.c_int32
    L.orc_magprism.argtypes = [C.c_int32, C.c_int32, C.c_int32, _d, _d, _d, _d, _d, _d,
                               C.c_double, C.c_double, C.c_double,
                               C.c_double, C.c_double, C.c_double, C.c_double, _d]
    L.orc_compress_row.argtypes = [_d, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _i, _f,
                                   C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.orc_compress_row.restype = C.c_int32
    L.orc_depth_weight.argtypes = [C.c_int32, C.c_int32, _d, _d, _d, _d, _d, _d, C.c_int32, _d, _d, _d,
                                   C.c_double, C.c_double, C.c_double, _d]
    L.orc_admm_iterate.argtypes = [C.c_int32, C.c_int32, _d, _d, _d, _d, _d, _d]
    L.orc_admm_iterate.restype = None
    L.orc_damping_add.argtypes = [P, _d, C.c_double, C.c_double, C.c_double, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                  C.c_int32, C.c_int32, _d, _d, _d, C.c_int32, C.c_int32, C.c_void_p,
                                  C.POINTER(C.c_double)]
    L.orc_damping_gradient_add.argtypes = [P, _d, C.c_double, C.c_double, C.c_int32, C.c_int32, C.c_int32, _d, _d, _d,
                                           C.c_int32, C.c_int32, _d, _d, _d, C.c_int32, C.c_int32,
                                           C.POINTER(C.c_double)]
    L.orc_cross_gradient_calculate.argtypes = [P, _d, C.c_int32, C.c_int32, C.c_int32, _d, _d, _d, C.c_int32, C.c_int32,
                                               _d, _d, _d, _d, C.c_int32, C.c_double, _i, _d, C.c_void_p,
                                               C.POINTER(C.c_int64), C.POINTER(C.c_int32)]
    L.orc_norm2.argtypes = [C.c_int64, _d]
    L.orc_norm2.restype = C.c_double
    _lib = L
    return L


class SparseMatrix:
    """Mirror of t_sparse_matrix (src/inversion/sparse_matrix.f90:31-98)."""

    def __init__(self, nl, ncolumns, nnz, nl_empty=0):
        self._p = lib().orc_csr_new(nl, ncolumns, nnz, nl_empty)
        if not self._p:
            raise ValueError("Wrong sizes in sparse_matrix_allocate_arrays!")

    def __del__(self):
        try:
            if self._p:
                lib().orc_csr_free(self._p)
                self._p = None
        except Exception:
            pass

    nl = property(lambda s: s._p.contents.nl)
    ncolumns = property(lambda s: s._p.contents.ncolumns)
    nel = property(lambda s: s._p.contents.nel)
    nl_nonempty = property(lambda s: s._p.contents.nl_nonempty)

    current_row = property(lambda s: s._p.contents.nl_current_all)   # get_current_row_number (:458-463)

    def reset(self):
        lib().orc_csr_reset(self._p)

    def add(self, value, column):
        if lib().orc_csr_add(self._p, float(value), int(column)) != 0:
            raise RuntimeError("Error in total number of elements in sparse_matrix_add!")

    def add_row(self, values, columns):
        values = np.ascontiguousarray(values, dtype=np.float32)
        columns = np.ascontiguousarray(columns, dtype=np.int32)
        if lib().orc_csr_add_row(self._p, len(values), values, columns) != 0:
            raise RuntimeError("Error in total number of elements in sparse_matrix_add_row!")

    def new_row(self):
        if lib().orc_csr_new_row(self._p) != 0:
            raise RuntimeError("Error in number of rows in sparse_matrix_new_row!")

    def add_empty_rows(self, nrows):
        lib().orc_csr_add_empty_rows(self._p, int(nrows))

    def finalize(self):
        rc = lib().orc_csr_finalize(self._p)
        if rc != 0:
            raise RuntimeError("sparse_matrix_finalize failed, code %d" % rc)

    def arrays(self):
        """(sa f32, ija i32 1-based, ijl i64 1-based, rowptr i32 1-based) copies."""
        c = self._p.contents
        nel, nne = int(c.nel), int(c.nl_nonempty)
        sa = np.ctypeslib.as_array(c.sa, shape=(max(nel, 1),))[:nel].copy()
        ija = np.ctypeslib.as_array(c.ija, shape=(max(nel, 1),))[:nel].copy()
        ijl = np.ctypeslib.as_array(c.ijl, shape=(nne + 1,)).copy()
        rowptr = np.ctypeslib.as_array(c.rowptr, shape=(max(nne, 1),))[:nne].copy()
        return sa, ija, ijl, rowptr

    def mult_vector(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        b = np.zeros(self.nl)
        lib().orc_csr_mult_vector(self._p, x, b)
        return b

    def add_mult_vector(self, x, b):
        lib().orc_csr_add_mult_vector(self._p, np.ascontiguousarray(x, dtype=np.float64), b)

    def part_mult_vector(self, x, ndata, line_start, param_shift):
        x = np.ascontiguousarray(x, dtype=np.float64)
        b = np.zeros(ndata)
        if lib().orc_csr_part_mult_vector(self._p, len(x), x, ndata, b, line_start, param_shift) != 0:
            raise RuntimeError("Wrong line index in sparse_matrix_part_mult_vector!")
        return b

    def trans_mult_vector(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        b = np.zeros(self.ncolumns)
        lib().orc_csr_trans_mult_vector(self._p, x, b)
        return b

    def add_trans_mult_vector(self, x, b):
        lib().orc_csr_add_trans_mult_vector(self._p, np.ascontiguousarray(x, dtype=np.float64), b)

    def normalize_columns(self):
        cn = np.zeros(self.ncolumns)
        lib().orc_csr_normalize_columns(self._p, cn)
        return cn


def _vol(s, n1, n2, n3):
    s = np.ascontiguousarray(s, dtype=np.float64).reshape(-1)
    assert s.size == n1 * n2 * n3
    return s


def forward_wavelet(s, n1, n2, n3, wavelet_type):
    """In-place on a copy; s is the flattened Fortran-order volume s(n1,n2,n3)."""
    s = _vol(s, n1, n2, n3).copy()
    if lib().orc_forward_wavelet(s, n1, n2, n3, wavelet_type) != 0:
        raise ValueError("Unknown wavelet type!")
    return s


def inverse_wavelet(s, n1, n2, n3, wavelet_type):
    s = _vol(s, n1, n2, n3).copy()
    if lib().orc_inverse_wavelet(s, n1, n2, n3, wavelet_type) != 0:
        raise ValueError("Unknown wavelet type!")
    return s


def lsqr_solve(niter, rmin, gamma, matrix, b):
    """lsqr_solve (lsqr_solver2.F90:321). Returns (x, r_history, iters); b is not modified."""
    u = np.array(b, dtype=np.float64)
    x = np.zeros(matrix.ncolumns)
    hist = np.zeros(max(niter, 1))
    it = C.c_int32(0)
    rc = lib().orc_lsqr_solve(matrix.nl, matrix.ncolumns, niter, rmin, gamma, matrix._p, u, x,
                              hist.ctypes.data, C.byref(it))
    if rc != 0:
        raise RuntimeError("lsqr_solve failed, code %d" % rc)
    return x, hist[:it.value].copy(), it.value


def lsqr_solve_sensit(niter, rmin, gamma, target_misfit, S, Cm, b, nelements, nx, ny, nz, ncomponents,
                      compression_type, wavelet_domain, solve_problem=(1, 0)):
    """lsqr_solve_sensit (lsqr_solver2.F90:47). Returns (x, r_history, iters)."""
    u = np.array(b, dtype=np.float64)
    nlines, ncolumns = S.nl + Cm.nl, S.ncolumns
    assert u.size == nlines
    x = np.zeros(ncolumns)
    hist = np.zeros(max(niter, 1))
    it = C.c_int32(0)
    sp = np.array(solve_problem, dtype=np.int32)
    rc = lib().orc_lsqr_solve_sensit(nlines, ncolumns, niter, rmin, gamma, target_misfit, S._p, Cm._p, u, x,
                                     sp, nelements, nx, ny, nz, ncomponents, compression_type,
                                     int(bool(wavelet_domain)), hist.ctypes.data, C.byref(it))
    if rc != 0:
        raise RuntimeError("lsqr_solve_sensit failed, code %d" % rc)
    return x, hist[:it.value].copy(), it.value


def _grid6(grid):
    return [np.ascontiguousarray(g, dtype=np.float64) for g in grid]


def graviprism_z(grid, xd, yd, zd):
    X1, X2, Y1, Y2, Z1, Z2 = _grid6(grid)
    out = np.zeros(X1.size)
    rc = lib().orc_graviprism_z(X1.size, X1, X2, Y1, Y2, Z1, Z2, xd, yd, zd, out)
    if rc != 0:
        raise RuntimeError("Data coordinate coincides with model grid boundary. Adjust the model grid!")
    return out


def gradiprism_zz(grid, xd, yd, zd):
    X1, X2, Y1, Y2, Z1, Z2 = _grid6(grid)
    out = np.zeros(X1.size)
    lib().orc_gradiprism_zz(X1.size, X1, X2, Y1, Y2, Z1, Z2, xd, yd, zd, out)
    return out


def gradiprism_full(grid, xd, yd, zd):
    """gradiprism_full (gravity_field.f90:207-309): numpy shape (6, n), components XX, YY, ZZ, XY, YZ, ZX -- the order of
    sensit_line_full(:, 1, 1..6) at sensitivity_gravmag.F90:207-209."""
    X1, X2, Y1, Y2, Z1, Z2 = _grid6(grid)
    out = np.zeros(6 * X1.size)
    rc = lib().orc_gradiprism_full(X1.size, X1, X2, Y1, Y2, Z1, Z2, xd, yd, zd, out)
    if rc == 3:
        raise RuntimeError("Zero denominator in gradiprism_full! Adjust the model grid.")
    if rc == 4:
        raise RuntimeError("Bad log argument in gradiprism_full! Adjust the model grid.")
    return out.reshape(6, X1.size)


def magprism(grid, xd, yd, zd, nmodel_comp, ndata_comp, mi, md, theta, intensity):
    """Returns sensit_line with numpy shape (ndata_comp, nmodel_comp, n) == Fortran (n, k, d)."""
    X1, X2, Y1, Y2, Z1, Z2 = _grid6(grid)
    out = np.zeros(X1.size * nmodel_comp * ndata_comp)
    rc = lib().orc_magprism(X1.size, nmodel_comp, ndata_comp, X1, X2, Y1, Y2, Z1, Z2, xd, yd, zd,
                            mi, md, theta, intensity, out)
    if rc != 0:
        raise RuntimeError("magprism failed, code %d" % rc)
    return out.reshape(ndata_comp, nmodel_comp, X1.size)


def compress_row(line, nx, ny, nz, compression_type, nel_compressed):
    """Row pipeline of sensitivity_gravmag.F90:230-295 on a column-weighted line.
    Returns dict(cols (1-based), vals, threshold, cost_full, cost_discarded, line_w)."""
    work = np.array(line, dtype=np.float64).reshape(-1)
    N = nx * ny * nz
    cols = np.zeros(N, dtype=np.int32)
    vals = np.zeros(N, dtype=np.float32)
    thr, cf, cd = C.c_double(0), C.c_double(0), C.c_double(0)
    nel = lib().orc_compress_row(work, nx, ny, nz, compression_type, nel_compressed, cols, vals,
                                 C.byref(thr), C.byref(cf), C.byref(cd))
    return dict(cols=cols[:nel].copy(), vals=vals[:nel].copy(), threshold=thr.value,
                cost_full=cf.value, cost_discarded=cd.value, line_w=work)


def depth_weight(wtype, grid, xd, yd, zd, power, beta=1.0, Z0=0.0):
    X1, X2, Y1, Y2, Z1, Z2 = _grid6(grid)
    xd, yd, zd = (np.ascontiguousarray(a, dtype=np.float64) for a in (xd, yd, zd))
    cw = np.zeros(X1.size)
    rc = lib().orc_depth_weight(wtype, X1.size, X1, X2, Y1, Y2, Z1, Z2, xd.size, xd, yd, zd,
                                power, beta, Z0, cw)
    if rc != 0:
        raise RuntimeError("depth weight failed, code %d" % rc)
    return cw


def admm_iterate(xmin, xmax, x, z, u):
    """admm_method_iterate_admm_arrays; xmin/xmax shape (n, nlithos); z, u updated in place."""
    xmin = np.ascontiguousarray(xmin, dtype=np.float64)
    xmax = np.ascontiguousarray(xmax, dtype=np.float64)
    n, nl = xmin.shape
    x0 = np.zeros(n)
    lib().orc_admm_iterate(n, nl, xmin.reshape(-1), xmax.reshape(-1),
                           np.ascontiguousarray(x, dtype=np.float64), z, u, x0)
    return x0


def norm2(x):
    x = np.ascontiguousarray(x, dtype=np.float64).reshape(-1)
    return lib().orc_norm2(x.size, x)


def calculate_data(S, model_val, ndata, ndata_components, problem_weight, column_weight, data_weight,
                   compression_type, nx, ny, nz, line_start=1, param_shift=0):
    """t_model%calculate_data, serial version (src/inversion/model.F90:220-307). model_val: (ncomponents, nelements);
    returns data_calc (ndata, ndata_components)."""
    model_val = np.atleast_2d(np.asarray(model_val, dtype=np.float64))
    cw = np.asarray(column_weight, dtype=np.float64)
    scaled = np.zeros_like(model_val)
    for k in range(model_val.shape[0]):                                  # :243-251
        nz_ = cw != 0.0
        scaled[k, nz_] = model_val[k, nz_] / cw[nz_]
        if compression_type > 0:                                         # :278-283
            scaled[k] = forward_wavelet(scaled[k].copy(), nx, ny, nz, compression_type)
    d = S.part_mult_vector(scaled.ravel(), ndata * ndata_components, line_start, param_shift)   # :288
    if problem_weight == 0.0:
        raise RuntimeError("Zero problem weight in model_calculate_data!")
    d = d / problem_weight                                               # :297
    d = d / np.asarray(data_weight, dtype=np.float64).ravel()            # :304
    return d.reshape(ndata, ndata_components)


def _dd(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def damping_add(matrix, b_RHS, alpha, problem_weight, norm_power, compression_type, nx, ny, nz, nsmaller, nelements,
                column_weight, model, model_ref, param_shift, wavelet_domain, local_weight=None):
    """damping_add (damping.F90:97-201) on full-grid arrays, the rank's slab = (nsmaller, nelements). b_RHS (the
    constraint part of the right-hand side) is filled in place; returns the damping cost."""
    cost = C.c_double(0.0)
    lw = None if local_weight is None else _dd(local_weight)
    rc = lib().orc_damping_add(matrix._p, b_RHS, alpha, problem_weight, norm_power, compression_type, nx, ny, nz,
                               nsmaller, nelements, _dd(column_weight), _dd(model), _dd(model_ref), param_shift,
                               int(bool(wavelet_domain)), None if lw is None else lw.ctypes.data, C.byref(cost))
    if rc != 0:
        raise RuntimeError("Sanity check failed in damping_add!" if rc == -1 else "sparse matrix overflow in damping_add")
    return cost.value


def damping_gradient_add(matrix, b_RHS, beta, problem_weight, nx, ny, nz, dX, dY, dZ, nsmaller, nelements, val_full,
                         column_weight, local_weight, param_shift, direction):
    """damping_gradient_add (damping_gradient.F90:93-203); returns the cost."""
    cost = C.c_double(0.0)
    rc = lib().orc_damping_gradient_add(matrix._p, b_RHS, beta, problem_weight, nx, ny, nz, _dd(dX), _dd(dY), _dd(dZ),
                                        nsmaller, nelements, _dd(val_full), _dd(column_weight), _dd(local_weight),
                                        param_shift, direction, C.byref(cost))
    if rc != 0:
        raise RuntimeError("Wrong direction in damping_gradient_add!" if rc == -1 else "sparse matrix overflow")
    return cost.value


def cross_gradient_calculate(matrix, b_RHS, nx, ny, nz, dX, dY, dZ, nsmaller, nparams_loc, model1, model2,
                             column_weight1, column_weight2, der_type, glob_weight, keep_model_constant=(0, 0)):
    """cross_gradient_calculate with add = .true. (cross_gradient.F90:220-391). Returns (cost[3], cross_grad(N), nnz,
    nl_nonempty)."""
    cost = np.zeros(3)
    cg = np.zeros(nx * ny * nz)
    nnz, nne = C.c_int64(0), C.c_int32(0)
    keep = np.ascontiguousarray(keep_model_constant, dtype=np.int32)
    rc = lib().orc_cross_gradient_calculate(matrix._p, b_RHS, nx, ny, nz, _dd(dX), _dd(dY), _dd(dZ), nsmaller,
                                            nparams_loc, _dd(model1), _dd(model2), _dd(column_weight1),
                                            _dd(column_weight2), der_type, glob_weight, keep, cost, cg.ctypes.data,
                                            C.byref(nnz), C.byref(nne))
    if rc != 0:
        raise RuntimeError("Unsupported derivative type!" if rc == -1 else "sparse matrix overflow")
    return cost, cg, nnz.value, nne.value
