"""tomofast-x_b200 -- Python host mirror of the reference's module interfaces over libtfx (C ABI).

The compute path is hand-written sm_100a CUDA inside libtfx.so (tomofast-x_b200/csrc); this module
only binds include/tfx.h with ctypes and mirrors the reference's names and argument meaning:

    reference (Fortran)                         here
    ------------------------------------------  ----------------------------------------------
    type(t_sparse_matrix) + bound procedures    SparseMatrix (sparse_matrix.f90:31-405)
    forward_wavelet / inverse_wavelet           forward_wavelet / inverse_wavelet (wavelet_transform.F90:37-70)
    lsqr_solve / lsqr_solve_sensit              lsqr_solve / lsqr_solve_sensit (lsqr_solver2.F90:47,321)
    calculate_and_write_sensit + read_...       calculate_sensit (sensitivity_gravmag.F90:82,648)

There is NO CPU fallback: every compute call fails loudly (TfxError) when libtfx.so or a CUDA device
is missing. The directory name contains a hyphen; import it as `tomofastx_b200` (alias module at
the repository root).
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libtfx.so")
_lib = None


class TfxError(RuntimeError):
    """Raised where the reference would call exit_MPI (src/utils/mpi_tools.F90:30-54)."""


def build(force=False, verbose=False):
    import importlib.util
    spec = importlib.util.spec_from_file_location("_tfx_build", os.path.join(_HERE, "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.build(force=force, verbose=verbose)


class SensitParams(C.Structure):
    """struct tfx_sensit_params (include/tfx.h)."""
    _fields_ = [("problem_type", C.c_int32), ("nx", C.c_int32), ("ny", C.c_int32), ("nz", C.c_int32),
                ("ndata", C.c_int32), ("ndata_components", C.c_int32), ("nmodel_components", C.c_int32),
                ("data_type", C.c_int32), ("compression_type", C.c_int32), ("compression_rate", C.c_double),
                ("problem_weight", C.c_double), ("mi", C.c_double), ("md", C.c_double), ("theta", C.c_double),
                ("intensity", C.c_double), ("cell0", C.c_int32), ("ncells_local", C.c_int32),
                ("param_shift", C.c_int32), ("ncolumns", C.c_int32)]


def lib():
    """Loads libtfx.so; raises TfxError if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise TfxError("libtfx.so is missing: run `python tomofast-x_b200/build.py` (there is no CPU fallback)")
    L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    vp, i32, i64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_double
    L.tfx_last_error.restype = C.c_char_p
    L.tfx_launch_count.restype = C.c_uint64
    L.tfx_init.argtypes = [C.c_int]
    L.tfx_set_option.argtypes = [C.c_char_p, C.c_int]
    L.tfx_comm_unique_id.argtypes = [C.c_char_p]
    L.tfx_comm_init.argtypes = [C.c_int, C.c_int, C.c_char_p]
    L.tfx_comm_allreduce_sum.argtypes = [vp, i64]
    L.tfx_sparse_matrix_initialize.argtypes = [C.POINTER(vp), i32, i32, i64, i32, i32]
    L.tfx_sparse_matrix_destroy.argtypes = [vp]
    L.tfx_sparse_matrix_reset.argtypes = [vp]
    L.tfx_sparse_matrix_finalize.argtypes = [vp, i32]
    L.tfx_sparse_matrix_add.argtypes = [vp, dbl, i32, i32]
    L.tfx_sparse_matrix_add_row.argtypes = [vp, i32, vp, vp, i32]
    L.tfx_sparse_matrix_new_row.argtypes = [vp, i32]
    L.tfx_sparse_matrix_add_empty_rows.argtypes = [vp, i32, i32]
    for n in ("mult_vector", "add_mult_vector", "trans_mult_vector", "add_trans_mult_vector"):
        getattr(L, "tfx_sparse_matrix_" + n).argtypes = [vp, vp, vp]
    L.tfx_sparse_matrix_part_mult_vector.argtypes = [vp, i32, vp, i32, vp, i32, i32, i32]
    for n in ("get_total_row_number", "get_current_row_number", "get_ncolumns"):
        getattr(L, "tfx_sparse_matrix_" + n).argtypes = [vp]
        getattr(L, "tfx_sparse_matrix_" + n).restype = i32
    for n in ("get_number_elements", "get_nnz"):
        getattr(L, "tfx_sparse_matrix_" + n).argtypes = [vp]
        getattr(L, "tfx_sparse_matrix_" + n).restype = i64
    L.tfx_sparse_matrix_normalize_columns.argtypes = [vp, vp]
    L.tfx_sparse_matrix_time_product.argtypes = [vp, C.c_int, vp, vp, C.c_int, C.POINTER(dbl)]
    L.tfx_sparse_matrix_device_bytes.argtypes = [vp]
    L.tfx_sparse_matrix_device_bytes.restype = i64
    L.tfx_sparse_matrix_drop_csr.argtypes = [vp]
    L.tfx_timer_stop.argtypes = [C.POINTER(dbl)]
    L.tfx_sparse_matrix_from_arrays.argtypes = [C.POINTER(vp), i32, i32, i32, i64, vp, vp, vp, vp]
    L.tfx_sparse_matrix_storage_kind.argtypes = [vp]
    L.tfx_sparse_matrix_export.argtypes = [vp, C.POINTER(i64), C.POINTER(i32), vp, vp, vp, vp]
    for n in ("tfx_forward_wavelet", "tfx_inverse_wavelet"):
        getattr(L, n).argtypes = [vp, i32, i32, i32, i32]
    for n in ("tfx_Haar3D", "tfx_iHaar3D", "tfx_DaubD43D", "tfx_iDaubD43D"):
        getattr(L, n).argtypes = [vp, i32, i32, i32]
    L.tfx_apply_wavelet_transform.argtypes = [i32, i32, i32, i32, i32, vp, i32, i32, i32, vp, i32, i32]
    L.tfx_lsqr_solve.argtypes = [i32, i32, i32, dbl, dbl, vp, vp, vp, i32]
    L.tfx_lsqr_solve_sensit.argtypes = [i32, i32, i32, dbl, dbl, dbl, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32,
                                        i32, i32, C.POINTER(dbl), i32, i32]
    L.tfx_lsqr_last_history.argtypes = [vp, i32, C.POINTER(i32), C.POINTER(i32)]
    L.tfx_lsqr_last_timing.argtypes = [C.POINTER(dbl), C.POINTER(dbl), C.POINTER(i32)]
    L.tfx_calculate_sensit.argtypes = [C.POINTER(vp), C.POINTER(SensitParams)] + [vp] * 11 + [vp, C.POINTER(dbl),
                                                                                              C.POINTER(i64)]
    L.tfx_sensit_assemble_rows.argtypes = [C.POINTER(vp), C.POINTER(SensitParams)] + [vp] * 11 + [i32, i32, vp,
                                                                                                   C.POINTER(dbl), C.POINTER(i64)]
    L.tfx_sensit_rows_info.argtypes = [vp, C.POINTER(i32), C.POINTER(i32), C.POINTER(i64)]
    L.tfx_sensit_rows_destroy.argtypes = [vp]
    L.tfx_get_load_balancing_nelements.argtypes = [i32, vp, i32, vp, vp]
    L.tfx_sensit_repartition.argtypes = [C.POINTER(vp), vp, i32, vp, i32, i32]
    L.tfx_sensit_lines.argtypes = [C.POINTER(SensitParams)] + [vp] * 6 + [i32] + [vp] * 4
    L.tfx_debug_math.argtypes = [i64, vp, vp, vp]
    L.tfx_grid_pin.argtypes = [i32] + [vp] * 6
    _lib = L
    return L


def _check(rc):
    if rc != 0:
        raise TfxError(lib().tfx_last_error().decode("utf-8", "replace") + " (code %d)" % rc)


def init(device=-1):
    _check(lib().tfx_init(int(device)))


def launch_count():
    return int(lib().tfx_launch_count())


def set_option(name, value):
    _check(lib().tfx_set_option(name.encode(), int(value)))


def timer_start():
    _check(lib().tfx_timer_start())


def timer_stop():
    """Elapsed device time (ms) on the library stream since timer_start()."""
    ms = C.c_double(0.0)
    _check(lib().tfx_timer_stop(C.byref(ms)))
    return ms.value


def synchronize():
    _check(lib().tfx_device_synchronize())


class Buffer:
    """A float64 vector in device memory (kind='device') or pinned host memory (kind='pinned')."""

    def __init__(self, n, kind="device"):
        self.n, self.kind = int(n), kind
        self.ptr = C.c_void_p()
        L = lib()
        fn = L.tfx_device_alloc if kind == "device" else L.tfx_host_alloc
        fn.argtypes = [C.POINTER(C.c_void_p), C.c_int64]
        _check(fn(C.byref(self.ptr), self.n * 8))

    def data_ptr(self):
        return self.ptr.value

    def numpy(self):
        """View of a pinned buffer / copy of a device buffer."""
        if self.kind == "pinned":
            return np.ctypeslib.as_array((C.c_double * self.n).from_address(self.ptr.value))
        out = np.empty(self.n)
        copy(out, self, self.n)
        return out

    def free(self):
        if self.ptr:
            L = lib()
            fn = L.tfx_device_free if self.kind == "device" else L.tfx_host_free
            fn.argtypes = [C.c_void_p]
            fn(self.ptr)
            self.ptr = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def copy(dst, src, n):
    """cudaMemcpyDefault of n float64 between numpy arrays / Buffers / device tensors."""
    L = lib()
    L.tfx_memcpy.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
    _check(L.tfx_memcpy(_ptr(dst), _ptr(src), int(n) * 8))


def zero(buf, n=None, offset=0):
    """Zero-fills n float64 of a device Buffer (default: all of it) starting at element `offset`."""
    L = lib()
    L.tfx_device_memset.argtypes = [C.c_void_p, C.c_int, C.c_int64]
    n = buf.n - offset if n is None else n
    _check(L.tfx_device_memset(_ptr(buf) + 8 * int(offset), 0, int(n) * 8))
    return buf


def device_mem_info():
    L = lib()
    f, t = C.c_int64(0), C.c_int64(0)
    L.tfx_device_mem_info.argtypes = [C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    _check(L.tfx_device_mem_info(C.byref(f), C.byref(t)))
    return f.value, t.value


def _ptr(a):
    """Host numpy array, torch CUDA tensor or raw int device pointer -> void*."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    if isinstance(a, int):
        return a
    if hasattr(a, "data_ptr"):
        return a.data_ptr()
    raise TypeError("unsupported vector type %r" % type(a))


def _f64(a):
    if isinstance(a, np.ndarray):
        return np.ascontiguousarray(a, dtype=np.float64)
    if hasattr(a, "data_ptr"):
        return a
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


class SparseMatrix:
    """t_sparse_matrix (src/inversion/sparse_matrix.f90:31-98); indices are 1-based like the reference's."""

    def __init__(self, nl=None, ncolumns=None, nnz=None, myrank=0, nl_empty=0, _handle=None):
        self._h = C.c_void_p()
        if _handle is not None:
            self._h = _handle
            return
        _check(lib().tfx_sparse_matrix_initialize(C.byref(self._h), nl, ncolumns, nnz, myrank, nl_empty))

    @classmethod
    def from_arrays(cls, nl, ncolumns, sa, ija, ijl, rowptr):
        sa = np.ascontiguousarray(sa, dtype=np.float32)
        ija = np.ascontiguousarray(ija, dtype=np.int32)
        ijl = np.ascontiguousarray(ijl, dtype=np.int64)
        rowptr = np.ascontiguousarray(rowptr, dtype=np.int32)
        h = C.c_void_p()
        _check(lib().tfx_sparse_matrix_from_arrays(C.byref(h), nl, ncolumns, len(rowptr), len(sa), sa.ctypes.data,
                                                   ija.ctypes.data, ijl.ctypes.data, rowptr.ctypes.data))
        return cls(_handle=h)

    def __del__(self):
        try:
            if self._h:
                lib().tfx_sparse_matrix_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass

    def reset(self):
        _check(lib().tfx_sparse_matrix_reset(self._h))

    def finalize(self, myrank=0):
        _check(lib().tfx_sparse_matrix_finalize(self._h, myrank))

    def add(self, value, column, myrank=0):
        _check(lib().tfx_sparse_matrix_add(self._h, float(value), int(column), myrank))

    def add_row(self, values, columns, myrank=0):
        values = np.ascontiguousarray(values, dtype=np.float32)
        columns = np.ascontiguousarray(columns, dtype=np.int32)
        _check(lib().tfx_sparse_matrix_add_row(self._h, len(values), values.ctypes.data, columns.ctypes.data, myrank))

    def new_row(self, myrank=0):
        _check(lib().tfx_sparse_matrix_new_row(self._h, myrank))

    def add_empty_rows(self, nrows, myrank=0):
        _check(lib().tfx_sparse_matrix_add_empty_rows(self._h, int(nrows), myrank))

    def normalize_columns(self):
        """normalize_columns (sparse_matrix.f90:414-443): scales the stored values, returns the column norms."""
        cn = np.zeros(self.get_ncolumns())
        _check(lib().tfx_sparse_matrix_normalize_columns(self._h, cn.ctypes.data))
        return cn

    def get_total_row_number(self):
        return lib().tfx_sparse_matrix_get_total_row_number(self._h)

    def get_current_row_number(self):
        return lib().tfx_sparse_matrix_get_current_row_number(self._h)

    def get_ncolumns(self):
        return lib().tfx_sparse_matrix_get_ncolumns(self._h)

    def get_number_elements(self):
        return lib().tfx_sparse_matrix_get_number_elements(self._h)

    def get_nnz(self):
        return lib().tfx_sparse_matrix_get_nnz(self._h)

    def time_product(self, transposed, x, b, reps=10):
        """Mean device milliseconds of one product on device-resident vectors (CUDA events)."""
        ms = C.c_double(0.0)
        _check(lib().tfx_sparse_matrix_time_product(self._h, int(transposed), _ptr(x), _ptr(b), int(reps), C.byref(ms)))
        return ms.value

    def device_bytes(self):
        return int(lib().tfx_sparse_matrix_device_bytes(self._h))

    def drop_csr(self):
        _check(lib().tfx_sparse_matrix_drop_csr(self._h))

    def storage_kind(self):
        return lib().tfx_sparse_matrix_storage_kind(self._h)

    def _prod(self, fn, x, b, nin, nout):
        x = _f64(x)
        if b is None:
            b = np.zeros(nout)
        _check(fn(self._h, _ptr(x), _ptr(b)))
        return b

    def mult_vector(self, x, b=None):
        return self._prod(lib().tfx_sparse_matrix_mult_vector, x, b, self.get_ncolumns(), self.get_total_row_number())

    def add_mult_vector(self, x, b):
        return self._prod(lib().tfx_sparse_matrix_add_mult_vector, x, b, 0, 0)

    def trans_mult_vector(self, x, b=None):
        return self._prod(lib().tfx_sparse_matrix_trans_mult_vector, x, b, self.get_total_row_number(), self.get_ncolumns())

    def add_trans_mult_vector(self, x, b):
        return self._prod(lib().tfx_sparse_matrix_add_trans_mult_vector, x, b, 0, 0)

    def part_mult_vector(self, x, ndata, line_start, param_shift, b=None, myrank=0):
        x = _f64(x)
        nel = x.size if isinstance(x, np.ndarray) else x.numel()
        if b is None:
            b = np.zeros(ndata)
        _check(lib().tfx_sparse_matrix_part_mult_vector(self._h, nel, _ptr(x), ndata, _ptr(b), line_start, param_shift,
                                                        myrank))
        return b

    def export(self):
        """(sa, ija, ijl, rowptr) in the reference's storage (1-based)."""
        nel, nne = C.c_int64(0), C.c_int32(0)
        _check(lib().tfx_sparse_matrix_export(self._h, C.byref(nel), C.byref(nne), None, None, None, None))
        sa = np.zeros(max(nel.value, 1), dtype=np.float32)
        ija = np.zeros(max(nel.value, 1), dtype=np.int32)
        ijl = np.zeros(nne.value + 1, dtype=np.int64)
        rowptr = np.zeros(max(nne.value, 1), dtype=np.int32)
        _check(lib().tfx_sparse_matrix_export(self._h, C.byref(nel), C.byref(nne), sa.ctypes.data, ija.ctypes.data,
                                              ijl.ctypes.data, rowptr.ctypes.data))
        return sa[:nel.value], ija[:nel.value], ijl, rowptr[:nne.value]


def _wavelet(fn, s, n1, n2, n3, *extra):
    if isinstance(s, np.ndarray):
        assert s.dtype == np.float64 and s.flags["C_CONTIGUOUS"] and s.size == n1 * n2 * n3
    _check(fn(_ptr(s), n1, n2, n3, *extra))
    return s


def forward_wavelet(s, n1, n2, n3, wavelet_type):
    """In place on s (flattened Fortran-order volume s(n1,n2,n3)); host array or device tensor."""
    return _wavelet(lib().tfx_forward_wavelet, s, n1, n2, n3, wavelet_type)


def inverse_wavelet(s, n1, n2, n3, wavelet_type):
    return _wavelet(lib().tfx_inverse_wavelet, s, n1, n2, n3, wavelet_type)


def Haar3D(s, n1, n2, n3):
    return _wavelet(lib().tfx_Haar3D, s, n1, n2, n3)


def iHaar3D(s, n1, n2, n3):
    return _wavelet(lib().tfx_iHaar3D, s, n1, n2, n3)


def DaubD43D(s, n1, n2, n3):
    return _wavelet(lib().tfx_DaubD43D, s, n1, n2, n3)


def iDaubD43D(s, n1, n2, n3):
    return _wavelet(lib().tfx_iDaubD43D, s, n1, n2, n3)


def apply_wavelet_transform(nelements, nx, ny, nz, ncomponents, v, fwd, compression_type, nproblems, solve_problem,
                            myrank=0, nbproc=1):
    sp = np.ascontiguousarray(solve_problem, dtype=np.int32)
    _check(lib().tfx_apply_wavelet_transform(nelements, nx, ny, nz, ncomponents, _ptr(v), int(bool(fwd)),
                                             compression_type, nproblems, sp.ctypes.data, myrank, nbproc))
    return v


def last_history():
    """(r_history, iters, fused) of the last solve."""
    it, fused = C.c_int32(0), C.c_int32(0)
    _check(lib().tfx_lsqr_last_history(None, 0, C.byref(it), C.byref(fused)))
    h = np.zeros(max(it.value, 1))
    _check(lib().tfx_lsqr_last_history(h.ctypes.data, it.value, C.byref(it), C.byref(fused)))
    return h[:it.value], it.value, bool(fused.value)


def last_iterations():
    """(executed, reported): loop bodies executed and the `iter - 1` the reference prints."""
    a, b = C.c_int32(0), C.c_int32(0)
    _check(lib().tfx_lsqr_last_iterations(C.byref(a), C.byref(b)))
    return a.value, b.value


def last_timing():
    """(loop_ms, sweep_ms, nsweeps) of the last solve, CUDA-event timed on the library stream."""
    a, b, n = C.c_double(0), C.c_double(0), C.c_int32(0)
    _check(lib().tfx_lsqr_last_timing(C.byref(a), C.byref(b), C.byref(n)))
    return a.value, b.value, n.value


def lsqr_solve(nlines, nelements, niter, rmin, gamma, matrix, u, x, myrank=0):
    """lsqr_solve (lsqr_solver2.F90:321): u (rhs) is overwritten, x receives the solution."""
    _check(lib().tfx_lsqr_solve(nlines, nelements, niter, rmin, gamma, matrix._h, _ptr(u), _ptr(x), myrank))
    return x


def lsqr_solve_sensit(nlines, ncolumns, niter, rmin, gamma, target_misfit, matrix_sensit, matrix_cons, u, x,
                      SOLVE_PROBLEM, nelements, nx, ny, nz, ncomponents, compression_type, WAVELET_DOMAIN,
                      myrank=0, nbproc=1):
    """lsqr_solve_sensit (lsqr_solver2.F90:47). Returns `memory` like the reference's out argument."""
    sp = np.ascontiguousarray([int(bool(s)) for s in SOLVE_PROBLEM], dtype=np.int32)
    mem = C.c_double(0.0)
    _check(lib().tfx_lsqr_solve_sensit(nlines, ncolumns, niter, rmin, gamma, target_misfit, matrix_sensit._h,
                                       matrix_cons._h if matrix_cons is not None else None, _ptr(u), _ptr(x),
                                       sp.ctypes.data, nelements, nx, ny, nz, ncomponents, compression_type,
                                       int(bool(WAVELET_DOMAIN)), C.byref(mem), myrank, nbproc))
    return mem.value


def _grid_ptrs(grid):
    arrs = [np.ascontiguousarray(g, dtype=np.float64) for g in grid]
    return arrs, [a.ctypes.data for a in arrs]


def calculate_sensit(par, grid, data_xyz, column_weight_full, data_weight):
    """calculate_and_write_sensit + read_sensitivity_kernel on the device (no disk round trip).
    Returns (SparseMatrix, sensit_nnz, comp_error, nnz_total)."""
    arrs, gp = _grid_ptrs(grid)
    dx, dy, dz = (np.ascontiguousarray(a, dtype=np.float64) for a in data_xyz)
    cw = np.ascontiguousarray(column_weight_full, dtype=np.float64)
    dw = np.ascontiguousarray(data_weight, dtype=np.float64)
    n = par.nx * par.ny * par.nz
    assert cw.size == n and dw.size == par.ndata * par.ndata_components and dx.size == par.ndata
    nnz_col = np.zeros(n, dtype=np.int32)
    cerr, tot = C.c_double(0.0), C.c_int64(0)
    h = C.c_void_p()
    _check(lib().tfx_calculate_sensit(C.byref(h), C.byref(par), *gp, dx.ctypes.data, dy.ctypes.data, dz.ctypes.data,
                                      cw.ctypes.data, dw.ctypes.data, nnz_col.ctypes.data, C.byref(cerr),
                                      C.byref(tot)))
    return SparseMatrix(_handle=h), nnz_col, cerr.value, tot.value


class SensitRows:
    """Row-sharded kernel of one problem resident on the device: stands in for the per-rank stream file
    sensit_<type>_<nbproc>_<rank> of calculate_and_write_sensit (sensitivity_gravmag.F90:143-318)."""

    def __init__(self, handle):
        self._h = handle

    def info(self):
        """(data0, ndata_loc, nnz_local): 0-based first station, number of stations, entries held."""
        a, b, n = C.c_int32(0), C.c_int32(0), C.c_int64(0)
        _check(lib().tfx_sensit_rows_info(self._h, C.byref(a), C.byref(b), C.byref(n)))
        return a.value, b.value, n.value

    def __del__(self):
        try:
            if self._h:
                lib().tfx_sensit_rows_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass


def sensit_assemble_rows(par, grid, data_xyz, column_weight_full, data_weight, myrank=0, nbproc=1):
    """Stage 1 of the multi-GPU assembly. Returns (SensitRows, sensit_nnz, comp_error, nnz_total), the last
    three reduced over the ranks."""
    arrs, gp = _grid_ptrs(grid)
    dx, dy, dz = (np.ascontiguousarray(a, dtype=np.float64) for a in data_xyz)
    cw = np.ascontiguousarray(column_weight_full, dtype=np.float64)
    dw = np.ascontiguousarray(data_weight, dtype=np.float64)
    n = par.nx * par.ny * par.nz
    assert cw.size == n and dw.size == par.ndata * par.ndata_components and dx.size == par.ndata
    nnz_col = np.zeros(n, dtype=np.int32)
    cerr, tot = C.c_double(0.0), C.c_int64(0)
    h = C.c_void_p()
    _check(lib().tfx_sensit_assemble_rows(C.byref(h), C.byref(par), *gp, dx.ctypes.data, dy.ctypes.data,
                                          dz.ctypes.data, cw.ctypes.data, dw.ctypes.data, myrank, nbproc,
                                          nnz_col.ctypes.data, C.byref(cerr), C.byref(tot)))
    return SensitRows(h), nnz_col, cerr.value, tot.value


def get_load_balancing_nelements(sensit_nnz, nbproc):
    """get_load_balancing_nelements (sensitivity_gravmag.F90:470-524): (nnz_at_cpu, nelements_at_cpu)."""
    nnz = np.ascontiguousarray(sensit_nnz, dtype=np.int32)
    a = np.zeros(nbproc, dtype=np.int64)
    b = np.zeros(nbproc, dtype=np.int32)
    _check(lib().tfx_get_load_balancing_nelements(nnz.size, nnz.ctypes.data, nbproc, a.ctypes.data, b.ctypes.data))
    return a, b


def sensit_repartition(rows, problem_slot, nelements_at_cpu, myrank=0, nbproc=1):
    """Stage 3: this rank's column-slab SparseMatrix (local column indices); consumes `rows`."""
    nel = np.ascontiguousarray(nelements_at_cpu, dtype=np.int32)
    assert nel.size == nbproc
    h = C.c_void_p()
    _check(lib().tfx_sensit_repartition(C.byref(h), rows._h, problem_slot, nel.ctypes.data, myrank, nbproc))
    return SparseMatrix(_handle=h)


def sensit_lines(par, grid, data_xyz):
    """Raw kernel lines, numpy shape (ndata, ndata_components, nmodel_components, ncells)."""
    arrs, gp = _grid_ptrs(grid)
    dx, dy, dz = (np.ascontiguousarray(a, dtype=np.float64) for a in data_xyz)
    n = par.nx * par.ny * par.nz
    out = np.zeros((dx.size, par.ndata_components, par.nmodel_components, n))
    _check(lib().tfx_sensit_lines(C.byref(par), *gp, dx.size, dx.ctypes.data, dy.ctypes.data, dz.ctypes.data,
                                  out.ctypes.data))
    return out


def grid_pin(grid):
    """Keeps a device copy of `grid` for the following assembly calls that pass the same arrays (tfx_grid_pin)."""
    arrs, gp = _grid_ptrs(grid)
    _check(lib().tfx_grid_pin(arrs[0].size, *gp))
    return arrs


def grid_unpin():
    _check(lib().tfx_grid_unpin())


def debug_math(y, x):
    """(tfx_log(x), log(x), tfx_atan2(y, x), atan2(y, x)) evaluated on the device (csrc/mathx.cuh vs the CUDA library)."""
    y = np.ascontiguousarray(y, dtype=np.float64)
    x = np.ascontiguousarray(x, dtype=np.float64)
    assert y.shape == x.shape and x.ndim == 1
    out = np.zeros((4, x.size))
    _check(lib().tfx_debug_math(x.size, y.ctypes.data, x.ctypes.data, out.ctypes.data))
    return out


# ---- communicator (one rank per GPU) ----------------------------------------------------------------
def comm_unique_id():
    buf = C.create_string_buffer(128)
    _check(lib().tfx_comm_unique_id(buf))
    return buf.raw


def comm_init(nranks, rank, uid):
    _check(lib().tfx_comm_init(nranks, rank, uid))


def comm_finalize():
    _check(lib().tfx_comm_finalize())


def comm_allreduce_sum(buf, count):
    _check(lib().tfx_comm_allreduce_sum(_ptr(buf), int(count)))


# ---- partitioning helpers (src/utils/parallel_tools.f90:46-86) ----------------------------------------
def calculate_nelements_at_cpu(nelements_total, myrank, nbproc):
    """Even split with the remainder on the first ranks (parallel_tools.f90:46-63)."""
    n = nelements_total // nbproc
    if myrank + 1 <= nelements_total - n * nbproc:
        n += 1
    return n


def get_nsmaller(nelements_total, myrank, nbproc):
    """Number of elements on ranks below myrank for the even split (parallel_tools.f90:68-86)."""
    return sum(calculate_nelements_at_cpu(nelements_total, r, nbproc) for r in range(myrank))


# ---- the reference's on-disk sensitivity formats (csrc/sensit_io.cu) ------------------------------------
def _sigs_io():
    L = lib()
    vp, i32, i64, dbl, cp = C.c_void_p, C.c_int32, C.c_int64, C.c_double, C.c_char_p
    L.tfx_create_sensit_directory.argtypes = [cp]
    L.tfx_write_sensit_file.argtypes = [vp, cp]
    L.tfx_write_sensit_metadata.argtypes = [C.POINTER(SensitParams), cp, i32, i32, dbl, i64, vp]
    L.tfx_read_sensitivity_metadata.argtypes = [C.POINTER(SensitParams), cp, i32, C.POINTER(i32), C.POINTER(dbl),
                                                C.POINTER(i64)]
    L.tfx_read_sensit_nnz.argtypes = [C.POINTER(SensitParams), cp, vp]
    L.tfx_write_depth_weight.argtypes = [C.POINTER(SensitParams), cp, vp]
    L.tfx_read_depth_weight.argtypes = [C.POINTER(SensitParams), cp, vp]
    L.tfx_read_sensitivity_kernel.argtypes = [C.POINTER(vp), C.POINTER(SensitParams), cp, vp, i32, i32, i32, i32, vp,
                                              C.POINTER(i64)]
    return L


def write_sensit_file(rows, path):
    """The rank's stream file sensit_<grav|magn>_<nbproc>_<rank> (sensitivity_gravmag.F90:143-318)."""
    L = _sigs_io()
    _check(L.tfx_create_sensit_directory(path.encode()))
    _check(L.tfx_write_sensit_file(rows._h, path.encode()))


def write_sensit_metadata(par, path, nbproc, depth_weighting_type, comp_error, nnz_total, sensit_nnz=None):
    L = _sigs_io()
    _check(L.tfx_create_sensit_directory(path.encode()))
    nnz = None if sensit_nnz is None else np.ascontiguousarray(sensit_nnz, dtype=np.int32)
    _check(L.tfx_write_sensit_metadata(C.byref(par), path.encode(), nbproc, depth_weighting_type, float(comp_error),
                                       int(nnz_total), None if nnz is None else nnz.ctypes.data))


def read_sensitivity_metadata(par, path, depth_weighting_type):
    """read_sensitivity_metadata (:974-1037): (nbproc_sensit, comp_error, nnz_total)."""
    L = _sigs_io()
    nb, ce, nt = C.c_int32(0), C.c_double(0.0), C.c_int64(0)
    _check(L.tfx_read_sensitivity_metadata(C.byref(par), path.encode(), depth_weighting_type, C.byref(nb), C.byref(ce),
                                           C.byref(nt)))
    return nb.value, ce.value, nt.value


def read_sensit_nnz(par, path):
    L = _sigs_io()
    out = np.zeros(par.nx * par.ny * par.nz, dtype=np.int32)
    _check(L.tfx_read_sensit_nnz(C.byref(par), path.encode(), out.ctypes.data))
    return out


def write_depth_weight(par, path, column_weight_full):
    L = _sigs_io()
    cw = np.ascontiguousarray(column_weight_full, dtype=np.float64)
    assert cw.size == par.nx * par.ny * par.nz
    _check(L.tfx_write_depth_weight(C.byref(par), path.encode(), cw.ctypes.data))


def read_depth_weight(par, path):
    L = _sigs_io()
    out = np.zeros(par.nx * par.ny * par.nz)
    _check(L.tfx_read_depth_weight(C.byref(par), path.encode(), out.ctypes.data))
    return out


def read_sensitivity_kernel(par, path, data_weight, depth_weighting_type, problem_slot, nelements_at_cpu, myrank=0,
                            nbproc=1):
    """read_sensitivity_kernel (:648-883): this rank's column-slab SparseMatrix from the stream files."""
    L = _sigs_io()
    dw = np.ascontiguousarray(data_weight, dtype=np.float64)
    nel = np.ascontiguousarray(nelements_at_cpu, dtype=np.int32)
    assert nel.size == nbproc and dw.size == par.ndata * par.ndata_components
    h, nloc = C.c_void_p(), C.c_int64(0)
    _check(L.tfx_read_sensitivity_kernel(C.byref(h), C.byref(par), path.encode(), dw.ctypes.data, depth_weighting_type,
                                         problem_slot, myrank, nbproc, nel.ctypes.data, C.byref(nloc)))
    return SparseMatrix(_handle=h)


def sensit_repartition_into(matrix, rows, problem_slot, nelements_at_cpu, myrank=0, nbproc=1):
    """Appends this rank's slab of `rows` to a SparseMatrix under construction (the reference's calling convention:
    read_sensitivity_kernel once per problem, then finalize; problem_joint_gravmag.F90:241-248)."""
    L = lib()
    L.tfx_sensit_repartition_into.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int32]
    nel = np.ascontiguousarray(nelements_at_cpu, dtype=np.int32)
    assert nel.size == nbproc
    _check(L.tfx_sensit_repartition_into(matrix._h, rows._h, problem_slot, nel.ctypes.data, myrank, nbproc))


def read_sensitivity_kernel_into(matrix, par, path, data_weight, depth_weighting_type, problem_slot, nelements_at_cpu,
                                 myrank=0, nbproc=1):
    L = _sigs_io()
    L.tfx_read_sensitivity_kernel_into.argtypes = [C.c_void_p, C.POINTER(SensitParams), C.c_char_p, C.c_void_p, C.c_int32,
                                                   C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.POINTER(C.c_int64)]
    dw = np.ascontiguousarray(data_weight, dtype=np.float64)
    nel = np.ascontiguousarray(nelements_at_cpu, dtype=np.int32)
    nloc = C.c_int64(0)
    _check(L.tfx_read_sensitivity_kernel_into(matrix._h, C.byref(par), path.encode(), dw.ctypes.data, depth_weighting_type,
                                              problem_slot, myrank, nbproc, nel.ctypes.data, C.byref(nloc)))
    return nloc.value


def sensit_rows_apply_weights(rows, problem_weight, data_weight):
    """combined_weight = real(problem_weight * data_weight, 4) applied in real(4) (sensitivity_gravmag.F90:837-843)."""
    L = lib()
    L.tfx_sensit_rows_apply_weights.argtypes = [C.c_void_p, C.c_double, C.c_void_p]
    dw = np.ascontiguousarray(data_weight, dtype=np.float64)
    _check(L.tfx_sensit_rows_apply_weights(rows._h, float(problem_weight), dw.ctypes.data))


def calculate_data(matrix_sensit, model_val, ndata, ndata_components, problem_weight, column_weight, data_weight,
                   compression_type, nx, ny, nz, line_start=1, param_shift=0, myrank=0, nbproc=1, data_calc=None):
    """t_model%calculate_data (model.F90:220-307). model_val: (ncomponents, nelements) C-ordered == Fortran
    val(nelements, ncomponents). Returns data_calc with shape (ndata, ndata_components)."""
    L = lib()
    vp, i32, dbl = C.c_void_p, C.c_int32, C.c_double
    L.tfx_calculate_data.argtypes = [vp, i32, i32, vp, i32, i32, dbl, vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, i32]
    m = _f64(model_val)
    cw = _f64(column_weight)
    dw = _f64(data_weight)
    nelements = cw.size if isinstance(cw, np.ndarray) else cw.n
    ncomp = (m.size if isinstance(m, np.ndarray) else m.n) // nelements
    if data_calc is None:
        data_calc = np.zeros((ndata, ndata_components))
    _check(L.tfx_calculate_data(matrix_sensit._h, nelements, ncomp, _ptr(m), ndata, ndata_components, float(problem_weight),
                                _ptr(cw), _ptr(dw), _ptr(data_calc), compression_type, nx, ny, nz, line_start, param_shift,
                                myrank, nbproc))
    return data_calc


def calculate_depth_weight(depth_weighting_type, grid, data_xyz, power, beta=1.0, Z0=0.0, nsmaller=0, nelements=None,
                           myrank=0, nbproc=1, column_weight=None):
    """calculate_depth_weight (weights_gravmag.f90:46-199) for the cells nsmaller+1 .. nsmaller+nelements of the
    full grid (6 arrays X1, X2, Y1, Y2, Z1, Z2). Returns column_weight(nelements)."""
    L = lib()
    vp, i32, dbl = C.c_void_p, C.c_int32, C.c_double
    L.tfx_calculate_depth_weight.argtypes = [i32, dbl, dbl, dbl, i32] + [vp] * 6 + [i32, vp, vp, vp, i32, i32, vp, i32, i32]
    g = [_f64(a) for a in grid]
    xyz = [_f64(a) for a in data_xyz]
    ntot = g[0].size if isinstance(g[0], np.ndarray) else g[0].n
    ndata = xyz[0].size if isinstance(xyz[0], np.ndarray) else xyz[0].n
    if nelements is None:
        nelements = ntot - nsmaller
    if column_weight is None:
        column_weight = np.zeros(nelements)
    _check(L.tfx_calculate_depth_weight(depth_weighting_type, float(power), float(beta), float(Z0), ntot,
                                        *[_ptr(a) for a in g], ndata, *[_ptr(a) for a in xyz], nsmaller, nelements,
                                        _ptr(column_weight), myrank, nbproc))
    return column_weight


# ---- constraint-matrix producers (csrc/cons.cu) ------------------------------------------------------
def _opt(a):
    return None if a is None else _f64(a)


def damping_add(matrix, b_RHS, alpha, problem_weight, norm_power, compression_type, nx, ny, nz, column_weight, model,
                model_ref, param_shift, wavelet_domain, local_weight=None, myrank=0, nbproc=1):
    """t_damping%add (damping.F90:97-201). Slab arrays (nelements); b_RHS (host array or Buffer) is the constraint
    part of the right-hand side and is updated in place. Returns the damping cost."""
    L = lib()
    vp, i32, dbl = C.c_void_p, C.c_int32, C.c_double
    L.tfx_damping_add.argtypes = [vp, i32, vp, dbl, dbl, dbl, i32, i32, i32, i32, i32, vp, vp, vp, i32, i32, vp, i32, i32,
                                  C.POINTER(dbl)]
    cw, m, ref, lw = _f64(column_weight), _f64(model), _f64(model_ref), _opt(local_weight)
    nelements = cw.size if isinstance(cw, np.ndarray) else cw.n
    nrows = b_RHS.size if isinstance(b_RHS, np.ndarray) else b_RHS.n
    cost = dbl(0.0)
    _check(L.tfx_damping_add(matrix._h, nrows, _ptr(b_RHS), float(alpha), float(problem_weight), float(norm_power),
                             compression_type, nx, ny, nz, nelements, _ptr(cw), _ptr(m), _ptr(ref), param_shift,
                             int(bool(wavelet_domain)), _ptr(lw), myrank, nbproc, C.byref(cost)))
    return cost.value


def damping_gradient_add(matrix, b_RHS, beta, problem_weight, nx, ny, nz, dX, dY, dZ, val_full, column_weight,
                         local_weight, param_shift, direction, myrank=0, nbproc=1):
    """t_damping_gradient%add (damping_gradient.F90:93-203). Returns the cost."""
    L = lib()
    vp, i32, dbl = C.c_void_p, C.c_int32, C.c_double
    L.tfx_damping_gradient_add.argtypes = [vp, i32, vp, dbl, dbl, i32, i32, i32, vp, vp, vp, i32, vp, vp, vp, i32, i32,
                                           i32, i32, C.POINTER(dbl)]
    dX, dY, dZ, vf, cw, lw = (_f64(a) for a in (dX, dY, dZ, val_full, column_weight, local_weight))
    nelements = cw.size if isinstance(cw, np.ndarray) else cw.n
    nrows = b_RHS.size if isinstance(b_RHS, np.ndarray) else b_RHS.n
    cost = dbl(0.0)
    _check(L.tfx_damping_gradient_add(matrix._h, nrows, _ptr(b_RHS), float(beta), float(problem_weight), nx, ny, nz,
                                      _ptr(dX), _ptr(dY), _ptr(dZ), nelements, _ptr(vf), _ptr(cw), _ptr(lw), param_shift,
                                      direction, myrank, nbproc, C.byref(cost)))
    return cost.value


def cross_gradient_calculate(matrix, b_RHS, nx, ny, nz, dX, dY, dZ, model1, model2, column_weight1, column_weight2,
                             der_type, glob_weight, keep_model_constant=(0, 0), myrank=0, nbproc=1, want_cross_grad=True):
    """t_cross_gradient%calculate with add = .true. (cross_gradient.F90:220-391). Returns (cost[3], cross_grad or None)."""
    L = lib()
    vp, i32, dbl = C.c_void_p, C.c_int32, C.c_double
    L.tfx_cross_gradient_calculate.argtypes = [vp, i32, vp, i32, i32, i32, vp, vp, vp, i32, vp, vp, vp, vp, i32, dbl, vp,
                                               i32, i32, vp, vp]
    dX, dY, dZ, m1, m2, w1, w2 = (_f64(a) for a in (dX, dY, dZ, model1, model2, column_weight1, column_weight2))
    nloc = w1.size if isinstance(w1, np.ndarray) else w1.n
    nrows = b_RHS.size if isinstance(b_RHS, np.ndarray) else b_RHS.n
    keep = np.ascontiguousarray(keep_model_constant, dtype=np.int32)
    cost = np.zeros(3)
    cg = np.zeros(nx * ny * nz) if want_cross_grad else None
    _check(L.tfx_cross_gradient_calculate(matrix._h, nrows, _ptr(b_RHS), nx, ny, nz, _ptr(dX), _ptr(dY), _ptr(dZ), nloc,
                                          _ptr(m1), _ptr(m2), _ptr(w1), _ptr(w2), der_type, float(glob_weight),
                                          keep.ctypes.data, myrank, nbproc, cost.ctypes.data, _ptr(cg)))
    return cost, cg


def admm_iterate_admm_arrays(xmin, xmax, x, z, u):
    """t_admm_method%iterate_admm_arrays (admm_method.F90:70-134); xmin/xmax shape (nelements, nlithos) C-ordered ==
    Fortran (nlithos, nelements); z and u are updated in place; returns x0."""
    L = lib()
    vp, i32 = C.c_void_p, C.c_int32
    L.tfx_admm_iterate_admm_arrays.argtypes = [i32, i32, vp, vp, vp, vp, vp, vp]
    xmin, xmax, x = _f64(xmin), _f64(xmax), _f64(x)
    n = x.size if isinstance(x, np.ndarray) else x.n
    nlithos = (xmin.size if isinstance(xmin, np.ndarray) else xmin.n) // n
    x0 = np.zeros(n) if isinstance(x, np.ndarray) else Buffer(n)
    _check(L.tfx_admm_iterate_admm_arrays(n, nlithos, _ptr(xmin), _ptr(xmax), _ptr(x), _ptr(z), _ptr(u), _ptr(x0)))
    return x0


class BufferView:
    """n float64 elements of a Buffer starting at element `offset` (device pointer arithmetic only)."""

    def __init__(self, buf, offset, n):
        assert 0 <= offset and offset + n <= buf.n
        self.buf, self.offset, self.n = buf, int(offset), int(n)

    def data_ptr(self):
        return self.buf.data_ptr() + 8 * self.offset


def rescale_model(model, weight, ncomponents=1):
    """rescale_model (model.F90:312-324): model(nelements, ncomponents) *= weight, in place."""
    L = lib()
    L.tfx_rescale_model.argtypes = [C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
    w = _f64(weight)
    nelements = w.size if isinstance(w, np.ndarray) else w.n
    _check(L.tfx_rescale_model(nelements, ncomponents, _ptr(model), _ptr(w)))
    return model


def model_update(val, delta_model, ncomponents=1):
    """t_model%update (model.F90:194-200): val += delta_model, in place."""
    L = lib()
    L.tfx_model_update.argtypes = [C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
    n = val.size if isinstance(val, np.ndarray) else val.n
    _check(L.tfx_model_update(n // ncomponents, ncomponents, _ptr(val), _ptr(delta_model)))
    return val
