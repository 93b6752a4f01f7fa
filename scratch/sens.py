import sys, numpy as np
sys.path.insert(0, '/root/repo')
from oracle import oracle as orc
from tests import mansf
cfg = mansf.Config()
io = mansf.Inversion(mansf.OracleBackend(orc), cfg, orc.admm_iterate)
b = io.build_rhs()
xo, ho = io.be.solve(cfg, io.S, io.C, b)
np.set_printoptions(linewidth=200, precision=2)
for eps in (1.0 + 2.3e-16, 3.0, 1.0/3.0):
    xp, hp = io.be.solve(cfg, io.S, io.C, b*eps)
    rel = np.abs(hp-ho)/ho
    print('scale', eps, 'max rel', rel.max(), 'at', rel.argmax(), 'x diff', np.abs(xp/eps-xo).max()/np.abs(xo).max())
    print(rel[:40])
