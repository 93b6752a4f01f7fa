"""Compressed products on the bench's Haar-5% matrix (256x256xNZ cells, ND stations): timing of S x / S^T u through
tfx_sparse_matrix_time_product. Usage (GPU box): python scratch/t16_products.py [nz] [ndata] [reps]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tomofastx_b200 as tfx
from tests.synth import depth_weight_type1, regular_grid, station_lattice

nx, ny = 256, 256
nz = int(sys.argv[1]) if len(sys.argv) > 1 else 64
nd = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
N = nx * ny * nz
tfx.init(0)
for kv in sys.argv[4:]:
    k, v = kv.split("="); tfx.set_option(k, int(v))
grid = regular_grid(nx, ny, nz)
xyz = station_lattice(nd, 100.0 * nx, 100.0 * ny, z=-0.1)
cw = depth_weight_type1(grid, 2.0, 0.0, 4.0e3)
par = tfx.SensitParams()
par.problem_type = 1
par.nx, par.ny, par.nz = nx, ny, nz
par.ndata, par.ndata_components, par.nmodel_components, par.data_type = nd, 1, 1, 1
par.compression_type, par.compression_rate = 1, 0.05
par.problem_weight = 1.0
par.cell0, par.ncells_local, par.param_shift, par.ncolumns = 0, N, 0, 2 * N
t0 = time.perf_counter()
S, _, cerr, nnz = tfx.calculate_sensit(par, grid, xyz, cw, np.ones((nd, 1)))
tfx.synchronize()
print("assembled nnz=%d in %.2f s, kind=%d" % (nnz, time.perf_counter() - t0, S.storage_kind()), flush=True)
rng = np.random.default_rng(1235)
x = tfx.Buffer(2 * N); u = tfx.Buffer(nd); q = tfx.Buffer(nd); t = tfx.Buffer(2 * N)
xl = np.zeros(2 * N); xl[:N] = rng.uniform(-1, 1, N)
tfx.copy(x, xl, 2 * N); tfx.copy(u, rng.uniform(-1, 1, nd), nd)
for name, tr, xi, yo in (("forward", 0, x, q), ("transposed", 1, u, t)):
    ms = S.time_product(tr, xi, yo, reps)
    print("%-10s %.4f ms  moved %.0f GB/s (6 B/nnz)  ref-accounting %.0f GB/s (8 B/nnz)" %
          (name, ms, 6.0 * nnz / ms / 1e6, 8.0 * nnz / ms / 1e6), flush=True)
lhs = float(np.dot(q.numpy(), u.numpy())); rhs = float(np.dot(x.numpy(), t.numpy()))
print("adjoint rel err %.3e" % (abs(lhs - rhs) / max(abs(lhs), abs(rhs))))
