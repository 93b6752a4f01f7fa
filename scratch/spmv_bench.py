"""Compressed-path microbench on a real wavelet-compressed gravity matrix (scratch)."""
import sys, time, numpy as np
sys.path.insert(0, '/root/repo')
import tomofastx_b200 as tfx
from tests.synth import depth_weight_type1, regular_grid, station_lattice
nx, ny, nz, nd = [int(a) for a in sys.argv[1:5]]
rate = float(sys.argv[5])
use_t16 = int(sys.argv[6]) if len(sys.argv) > 6 else 1
tfx.init(0)
tfx.set_option("t16_min_nnz", 0 if use_t16 else 2**31 - 1)
N = nx * ny * nz
grid = regular_grid(nx, ny, nz)
xyz = station_lattice(nd, 100.0 * nx, 100.0 * ny, z=-0.1)
cw = depth_weight_type1(grid, 2.0, 0.0, 4e3)
par = tfx.SensitParams()
par.problem_type = 1
par.nx, par.ny, par.nz = nx, ny, nz
par.ndata, par.ndata_components, par.nmodel_components, par.data_type = nd, 1, 1, 1
par.compression_type, par.compression_rate = 1, rate
par.problem_weight = 1.0
par.cell0, par.ncells_local, par.param_shift, par.ncolumns = 0, N, 0, 2 * N
t0 = time.perf_counter()
S, nnzcol, cerr, tot = tfx.calculate_sensit(par, grid, xyz, cw, np.ones((nd, 1)))
tfx.synchronize()
print("assembly %.2f s, nnz %.4g, comp_error %.3e, kind %d, device bytes %.3f GB" %
      (time.perf_counter() - t0, tot, cerr, S.storage_kind(), S.device_bytes() / 1e9), flush=True)
x = tfx.Buffer(2 * N); u = tfx.Buffer(nd); q = tfx.Buffer(nd); t = tfx.Buffer(2 * N)
rng = np.random.default_rng(1)
tfx.copy(x, rng.standard_normal(2 * N), 2 * N); tfx.copy(u, rng.standard_normal(nd), nd)
for name, tr, xi, yo in (("fwd  S x ", 0, x, q), ("trans S^T u", 1, u, t)):
    ms = S.time_product(tr, xi, yo, 20)
    print("%s: %.3f ms  -> %.0f GB/s at 8 B/nnz (reference accounting), %.0f GB/s at 6 B/nnz (bytes moved)" %
          (name, ms, 8.0 * tot / ms / 1e6, 6.0 * tot / ms / 1e6), flush=True)
