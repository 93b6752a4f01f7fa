"""Times the 3-D wavelet transforms on device-resident volumes (CUDA events on the library stream)."""
import sys
import numpy as np
sys.path.insert(0, ".")
import tomofastx_b200 as tfx

tfx.init(0)
rng = np.random.default_rng(0)
for (nx, ny, nz) in ((256, 256, 64), (512, 512, 128), (1024, 1024, 128)):
    N = nx * ny * nz
    vol = tfx.Buffer(N)
    tfx.copy(vol, rng.uniform(-1, 1, N), N)
    for slab, tile, fuse in ((0, 0, 0), (0, 0, 1), (0, 32, 1), (0, 128, 1)):
        tfx.set_option("wavelet_slab_mb", slab)
        tfx.set_option("wavelet_tile_kb", tile)
        tfx.set_option("wavelet_fuse12", fuse)
        for wname, wtype in (("haar", 1), ("d4", 2)):
            if wtype == 2 and (fuse == 0 or tile != 0):
                continue
            for _ in range(2):
                tfx.forward_wavelet(vol, nx, ny, nz, wtype); tfx.inverse_wavelet(vol, nx, ny, nz, wtype)
            tfx.timer_start()
            for _ in range(5):
                tfx.forward_wavelet(vol, nx, ny, nz, wtype); tfx.inverse_wavelet(vol, nx, ny, nz, wtype)
            ms = tfx.timer_stop() / 10.0
            # forward and inverse separately
            tfx.timer_start()
            for _ in range(5):
                tfx.forward_wavelet(vol, nx, ny, nz, wtype)
            msf = tfx.timer_stop() / 5.0
            tfx.timer_start()
            for _ in range(5):
                tfx.inverse_wavelet(vol, nx, ny, nz, wtype)
            msi = tfx.timer_stop() / 5.0
            print("fuse12=%d fwd %.4f inv %.4f | " % (fuse, msf, msi), end="")
            print("%dx%dx%d slab=%2d MB tile=%d KB %-4s %.4f ms/transform  %.0f GB/s at 16 B/elem (%.3f of 6456)" %
                  (nx, ny, nz, slab, tile, wname, ms, 16.0 * N / ms / 1e6, 16.0 * N / ms / 1e6 / 6456), flush=True)
    del vol
