import sys, os
import numpy as np
sys.path.insert(0, ".")
import tomofastx_b200 as tfx
from tomofastx_b200 import configs
from oracle import oracle
from tests.test_gpu_config_d import oracle_config_d, oracle_from_export, GOLDEN
tfx.init(0)
c = configs.load_twobody(GOLDEN, station_stride=4)
for graph in (1, 0):
    tfx.set_option("lsqr_graph", graph)
    got = configs.run_config_d(tfx, c, compression_type=2)
    if graph == 1:
        So = oracle_from_export(oracle, c["ndata"], 2 * 3 * c["N"], got["S"].export())
        want = oracle_config_d(oracle, c, So, got["column_weight"], 2)
    h, ho = got["histories"][0], want["histories"][0]
    rel = np.abs(h - ho) / ho
    print("graph", graph, "argmax", np.argmax(rel), rel.max())
    print(" h ", h[30:42]); print(" ho", ho[30:42])
    print(" rel", rel[30:42])
