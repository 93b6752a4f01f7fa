"""One forward Haar + one forward D4 transform of a 512x512x128 volume (for ncu)."""
import sys
import numpy as np
sys.path.insert(0, ".")
import tomofastx_b200 as tfx
tfx.init(0)
nx, ny, nz = 512, 512, 128
N = nx * ny * nz
vol = tfx.Buffer(N)
tfx.copy(vol, np.random.default_rng(0).uniform(-1, 1, N), N)
tfx.set_option("wavelet_slab_mb", int(sys.argv[1]) if len(sys.argv) > 1 else 0)
for wtype in (1, 2):
    tfx.forward_wavelet(vol, nx, ny, nz, wtype)
    tfx.inverse_wavelet(vol, nx, ny, nz, wtype)
tfx.synchronize()
