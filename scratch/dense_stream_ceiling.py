"""Diagnostic: streaming ceiling of the dense sweep's access pattern (TMA ring without the products) next to the real
kernel, on a block of config B's shape. Usage (GPU box): python scratch/dense_stream_ceiling.py [ncols]"""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tomofastx_b200 as tfx
from tests.synth import depth_weight_type1, regular_grid, station_lattice

nx, ny, nz, ndata = 256, 256, int(sys.argv[1]) if len(sys.argv) > 1 else 16, 10000
N = nx * ny * nz
tfx.init(0)
grid = regular_grid(nx, ny, nz)
xyz = station_lattice(ndata, 100.0 * nx, 100.0 * ny, z=-0.1)
cw = depth_weight_type1(grid, 2.0, 0.0, 4.0e3)
par = tfx.SensitParams()
par.problem_type = 1
par.nx, par.ny, par.nz = nx, ny, nz
par.ndata, par.ndata_components, par.nmodel_components, par.data_type = ndata, 1, 1, 1
par.compression_type, par.compression_rate = 0, 1.0
par.problem_weight = 1.0
par.cell0, par.ncells_local, par.param_shift, par.ncolumns = 0, N, 0, 2 * N
S, _, _, _ = tfx.calculate_sensit(par, grid, xyz, cw, np.ones((ndata, 1)))
C = tfx.SparseMatrix.from_arrays(N, 2 * N, np.full(N, 1e-11, dtype=np.float32), np.arange(1, N + 1, dtype=np.int32),
                                 np.arange(1, N + 2, dtype=np.int64), np.arange(1, N + 1, dtype=np.int32))
b = np.zeros(ndata + N); b[:ndata] = np.random.default_rng(0).standard_normal(ndata)
u, x = tfx.Buffer(ndata + N), tfx.Buffer(2 * N)
tfx.set_option("profile_sweeps", 1)
bytes_ = 4.0 * ((ndata + 3) // 4 * 4) * N
for name, opts in (("512x4_f2f0", {"dense_f2f_rows": 0}), ("512x4_f2f2", {}), ("512x4_f2f_all", {"dense_f2f_rows": 99}),
                   ("stream_only", {"dense_stream_only": 1}), ("1024x2_f2f0", {"dense_vec4": 0, "dense_f2f_rows": 0})):
    if len(sys.argv) > 2 and name not in sys.argv[2].split(","):
        continue
    tfx.set_option("dense_stream_only", 0); tfx.set_option("dense_vec4", 1); tfx.set_option("dense_f2f_rows", 2)
    for k, v in opts.items():
        tfx.set_option(k, v)
    for it in (3, 10):
        tfx.copy(u, b, ndata + N)
        tfx.lsqr_solve_sensit(ndata + N, 2 * N, it, 1e-300, 0.0, 0.0, S, C, u, x, [1, 0], N, nx, ny, nz, 1, 0, True)
    loop_ms, sweep_ms, ns = tfx.last_timing()
    h, it, fused = tfx.last_history()
    print("%-22s sweep %.3f ms  %.0f GB/s  r_last %.12e" % (name, sweep_ms / ns, bytes_ / (sweep_ms / ns) / 1e6, h[-1]), flush=True)
