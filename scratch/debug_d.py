import sys, os
import numpy as np
sys.path.insert(0, ".")
import tomofastx_b200 as tfx
from tomofastx_b200 import configs
from oracle import oracle
from tests.test_gpu_config_d import oracle_config_d, oracle_from_export, GOLDEN
tfx.init(0)
c = configs.load_twobody(GOLDEN, station_stride=4)
got = configs.run_config_d(tfx, c, compression_type=2)
N, nd, ncomp = c["N"], c["ndata"], c["ncomp"]
So = oracle_from_export(oracle, nd, 2 * 3 * N, got["S"].export())
want = oracle_config_d(oracle, c, So, got["column_weight"], 2)
b = want["rhs"][0]; ho = want["histories"][0]
Cg = tfx.SparseMatrix(ncomp * N, 2 * ncomp * N, ncomp * N)
bg = np.zeros_like(b); bg[:nd] = b[:nd]
m0 = np.full((ncomp, N), c["start_value"]); prior = np.zeros((ncomp, N))
for k in range(ncomp):
    tfx.damping_add(Cg, bg[nd:], c["alpha"], 1.0, 2.0, 2, c["nx"], c["ny"], c["nz"], got["column_weight"], m0[k], prior[k], ncomp * N + k * N, True)
Cg.finalize()
hist = {}
for name, opts in (("strict", {"strict_order": 1}), ("fast", {}), ("fast_nograph", {"lsqr_graph": 0})):
    for k, v in opts.items(): tfx.set_option(k, v)
    u = b.copy(); x = np.zeros(2 * ncomp * N)
    tfx.lsqr_solve_sensit(len(u), x.size, 100, 1e-13, 0.0, 0.0, got["S"], Cg, u, x, [0, 1], N, c["nx"], c["ny"], c["nz"], ncomp, 2, True)
    hist[name] = tfx.last_history()[0].copy()
    tfx.set_option("strict_order", 0); tfx.set_option("lsqr_graph", 1)
np.set_printoptions(linewidth=200, precision=6)
print("k   oracle        strict        fast          fast_nograph")
for k in range(26):
    print(k, ho[k], hist["strict"][k], hist["fast"][k], hist["fast_nograph"][k])
print("|b_d|", np.linalg.norm(b[:nd]), "|b_c|", np.linalg.norm(b[nd:]))
solve_o = lambda rhs: oracle.lsqr_solve_sensit(100, 1e-13, 0.0, 0.0, So, want["C0"], rhs, N, c["nx"], c["ny"], c["nz"], ncomp, 2, True, solve_problem=(0, 1))[1]
rng = np.random.default_rng(0)
for eps in (1e-16, 1e-15, 4e-14, 1e-12):
    env = np.zeros_like(ho)
    for t in range(3):
        env = np.maximum(env, np.abs(solve_o(b * (1.0 + eps * rng.standard_normal(b.size))) - ho) / ho)
    print("eps", eps, "env[0:26]", env[:26])
relf = np.abs(hist["fast"] - ho) / ho
print("rel_f", relf[:26])
