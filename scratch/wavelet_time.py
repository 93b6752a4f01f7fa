"""Timing of the 3-D wavelet transforms on a device-resident volume. Usage: python scratch/wavelet_time.py nx ny nz"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tomofastx_b200 as tfx
nx, ny, nz = (int(v) for v in sys.argv[1:4]) if len(sys.argv) > 3 else (256, 256, 64)
N = nx * ny * nz
tfx.init(0)
vol = tfx.Buffer(N)
tfx.copy(vol, np.random.default_rng(0).uniform(-1, 1, N), N)
for wname, wtype in (("haar", 1), ("d4", 2)):
    for _ in range(2):
        tfx.forward_wavelet(vol, nx, ny, nz, wtype); tfx.inverse_wavelet(vol, nx, ny, nz, wtype)
    tfx.timer_start()
    for _ in range(10):
        tfx.forward_wavelet(vol, nx, ny, nz, wtype); tfx.inverse_wavelet(vol, nx, ny, nz, wtype)
    ms = tfx.timer_stop() / 20.0
    print("%dx%dx%d %-5s %.4f ms per transform  %.0f GB/s (16 B/element)" % (nx, ny, nz, wname, ms, 16.0 * N / ms / 1e6))
