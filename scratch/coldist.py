import sys, time, numpy as np
sys.path.insert(0, '/root/repo')
from oracle import oracle as orc
from tests.synth import regular_grid, station_lattice, depth_weight_type1
nx, ny, nz, nd, rate = [int(a) for a in sys.argv[1:5]] + [float(sys.argv[5])]
N = nx*ny*nz
grid = regular_grid(nx, ny, nz)
xs, ys, zs = station_lattice(nd, 100.0*nx, 100.0*ny)
cw = depth_weight_type1(grid, 2.0, 0.0, 4e3)
nel = int(rate*N)
cnt = np.zeros(N, dtype=np.int64)
t0 = time.time()
rows = []
for i in range(nd):
    line = orc.graviprism_z(grid, float(xs[i]), float(ys[i]), float(zs[i])) * cw
    r = orc.compress_row(line, nx, ny, nz, 1, nel)
    cnt[r["cols"]-1] += 1
    rows.append(r["cols"].astype(np.int32)-1)
print("time", time.time()-t0, "nnz", cnt.sum(), "nel", nel)
np.save("/tmp/coldist_%d_%d_%d_%d.npy" % (nx,ny,nz,nd), cnt)
np.save("/tmp/rows_%d_%d_%d_%d.npy" % (nx,ny,nz,nd), np.concatenate(rows))
used = cnt[cnt>0]
print("columns used: %.3f" % (len(used)/N))
tot = cnt.sum()
for d in (1.0, 0.9, 0.67, 0.5, 0.25, 0.1, 0.05, 0.02, 0.01, 0.005):
    sel = cnt >= d*nd
    print("density >= %.3f: cols %.4f of N, nnz share %.3f" % (d, sel.sum()/N, cnt[sel].sum()/tot))
for L in (1,2,4,8,16,32,64,128):
    sel = (cnt>0)&(cnt<=L)
    print("len <= %d: cols %.4f of N, nnz share %.4f" % (L, sel.sum()/N, cnt[sel].sum()/tot))
