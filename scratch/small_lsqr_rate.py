"""LSQR iteration rate on the small configs (A: mansf_slice shape, D: 2body shape): launch-bound regime."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tomofastx_b200 as tfx
from tests.synth import make_problem
tfx.init(0)
for name, kw, niter in (("A-like 2x128x32, 256 data, Haar 0.15", dict(nx=2, ny=128, nz=32, ndata=256, compression_type=1, rate=0.15), 400),
                        ("D-like 67x67x30, 1681 data x3 comp, D4 0.3", dict(nx=67, ny=67, nz=30, ndata=1681, compression_type=2, rate=0.3,
                                                                           problem_type=2, nmodel_components=3), 100)):
    pb = make_problem(**kw)
    S, _, _, nnz = tfx.calculate_sensit(pb.par, pb.grid, pb.data_xyz, pb.cw, pb.dw)
    ncol = pb.ncolumns; N = pb.N; ncomp = kw.get("nmodel_components", 1)
    nl = S.get_total_row_number()
    nc = N * ncomp
    C = tfx.SparseMatrix.from_arrays(nc, ncol, np.full(nc, 1e-7, dtype=np.float32), np.arange(1, nc + 1, dtype=np.int32),
                                     np.arange(1, nc + 2, dtype=np.int64), np.arange(1, nc + 1, dtype=np.int32))
    b = np.zeros(nl + nc); b[:nl] = np.random.default_rng(0).standard_normal(nl)
    u, x = tfx.Buffer(nl + nc), tfx.Buffer(ncol)
    for graph in (0, 1):
      tfx.set_option("lsqr_graph", graph)
      for it in (5, niter):
        tfx.copy(u, b, nl + nc)
        l0 = tfx.launch_count()
        t0 = time.perf_counter()
        tfx.lsqr_solve_sensit(nl + nc, ncol, it, 1e-300, 0.0, 0.0, S, C, u, x, [1, 0], N, pb.nx, pb.ny, pb.nz, ncomp, kw["compression_type"], True)
        wall = time.perf_counter() - t0
      loop_ms, _, _ = tfx.last_timing()
      h, iters, fused = tfx.last_history()
      print("%-44s graph=%d nnz=%d kind=%d iters=%d  %.1f us/it (device loop)  %.1f us/it (wall)  launches/it=%.1f  r_last=%.15e" %
            (name, graph, nnz, S.storage_kind(), iters, 1e3 * loop_ms / iters, 1e6 * wall / iters, (tfx.launch_count() - l0) / iters, h[-1]), flush=True)
