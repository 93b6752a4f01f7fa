"""Compressed row pipeline on a few stations (for an ncu launch list):
    python scratch/assembly_breakdown.py [nz] [ndata] [problem_type] [option=value | nx=.. | ny=.. | wt=..] ..."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tomofastx_b200 as tfx
from tomofastx_b200.synth import depth_weight_type1, regular_grid, station_lattice
nx, ny, wt = 256, 256, 1
for kv in list(sys.argv[4:]):
    k, v = kv.split('=')
    if k in ('nx', 'ny', 'wt'):
        globals()[k] = int(v); sys.argv.remove(kv)
nz = int(sys.argv[1]) if len(sys.argv) > 1 else 64
nd = int(sys.argv[2]) if len(sys.argv) > 2 else 64
N = nx * ny * nz
tfx.init(0)
grid = regular_grid(nx, ny, nz)
xyz = station_lattice(nd, 100.0 * nx, 100.0 * ny, z=-0.1)
cw = depth_weight_type1(grid, 2.0, 0.0, 4.0e3)
par = tfx.SensitParams()
ptype = int(sys.argv[3]) if len(sys.argv) > 3 else 1
for kv in sys.argv[4:]:
    k, v = kv.split("="); tfx.set_option(k, int(v))
par.problem_type = ptype
par.nx, par.ny, par.nz = nx, ny, nz
par.ndata, par.ndata_components, par.nmodel_components, par.data_type = nd, 1, 1, 1
par.mi, par.md, par.theta, par.intensity = 60.0, 10.0, 0.0, 50000.0
par.compression_type, par.compression_rate = wt, 0.05
par.problem_weight = 1.0
par.cell0, par.ncells_local, par.param_shift, par.ncolumns = 0, N, 0, 2 * N
for rep in range(2):
    tfx.synchronize(); t0 = time.perf_counter()
    rows, nnz_col, cerr, tot = tfx.sensit_assemble_rows(par, grid, xyz, cw, np.ones((nd, 1)))
    tfx.synchronize(); dt = time.perf_counter() - t0
    print("rows=%d  %.3f ms/row  %.2e cell evaluations/s" % (nd, 1e3 * dt / nd, nd * N / dt), flush=True)
    del rows
