import sys, numpy as np
sys.path.insert(0, '/root/repo')
import tomofastx_b200 as tfx
from oracle import oracle as orc
from tests import mansf
cfg = mansf.Config()
io = mansf.Inversion(mansf.OracleBackend(orc), cfg, orc.admm_iterate)
ig = mansf.Inversion(mansf.TfxBackend(tfx), cfg, orc.admm_iterate)
b = io.build_rhs()
xo, ho = io.be.solve(cfg, io.S, io.C, b)
xg, hg = ig.be.solve(cfg, ig.S, ig.C, b)
rel = np.abs(hg-ho)/ho
np.set_printoptions(linewidth=200, precision=3)
print(rel)
print(ho[:12]); print(hg[:12])
print(np.abs(xg-xo).max(), np.abs(xo).max())
# compare matrices
so = io.S.arrays(); sg = ig.S.export()
print('nel', len(so[0]), len(sg[0]), 'cols equal', np.array_equal(so[1], sg[1]), 'max val diff', np.abs(so[0]-sg[0]).max(), np.abs(so[0]).max())
