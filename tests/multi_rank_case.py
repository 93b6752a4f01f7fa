"""One rank of a column-split LSQR solve (launched by torchrun from tests/test_gpu_multi.py and
tests/test_dist_model.py).

    python -m torch.distributed.run --nproc-per-node P tests/multi_rank_case.py --backend nccl|model

The decomposition is the reference's (lsqr_solver2.F90:16 "parallelized by model parameters"): rank r owns
the column slab [nsmaller, nsmaller + nelements) (parallel_tools.f90:46-86), every rank holds all data rows,
the products S_loc * v_loc are summed over ranks (MPI_Allreduce at lsqr_solver2.F90:214) and |v|^2 is a
scalar all-reduce (:514).

 --backend nccl  : the product path -- libtfx on one GPU per rank, NCCL inside the library.
 --backend model : host model of the same decomposition on CPU (numpy products, torch.distributed gloo
                   all-reduce) -- exercises the partition helpers, the slab builders and the summation
                   structure without a GPU.
Rank 0 gathers the slabs and compares with the single-rank oracle solve of the full problem.
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def full_problem(kind, seed=7):
    rng = np.random.default_rng(seed)
    nx, ny, nz, ndata = 7, 6, 5, 40
    N = nx * ny * nz
    if kind == "dense" or kind == "dense_xgrad":
        nel = N
        cols = [np.arange(N, dtype=np.int32) for _ in range(ndata)]
    else:
        nel = N // 3
        cols = [np.sort(rng.choice(N, size=nel, replace=False)).astype(np.int32) for _ in range(ndata)]
    vals = [rng.standard_normal(nel).astype(np.float32) for _ in range(ndata)]
    alpha = (0.3 + 0.1 * (np.arange(N) % 4)).astype(np.float32)        # damping block alpha_p * I (damping.F90:158-179)
    # constraint block as rows (columns, values) in the full column space
    crow = [(np.array([p], dtype=np.int32), alpha[p:p + 1]) for p in range(N)]
    if kind.endswith("xgrad"):
        # gradient-like rows that couple a cell with its +x / +y / +z neighbours (cross_gradient.F90:220-391 shape):
        # next to a slab boundary their entries live on two ranks ("shared" rows); every 7th row is left without
        # entries at all (zero derivatives are dropped, sparse_matrix.f90:219) but keeps a right-hand side
        for p in range(N):
            nb = [q for q in (p + 1, p + nx, p + nx * ny) if q < N]
            if p % 7 == 3 or not nb:
                crow.append((np.zeros(0, dtype=np.int32), np.zeros(0, dtype=np.float32)))
                continue
            c = np.array([p] + nb, dtype=np.int32)
            v = (0.2 * rng.standard_normal(len(c))).astype(np.float32)
            crow.append((c, v))
    ncons = len(crow)
    b = np.concatenate([rng.standard_normal(ndata), 0.01 * rng.standard_normal(ncons)])
    return dict(nx=nx, ny=ny, nz=nz, ndata=ndata, N=N, cols=cols, vals=vals, alpha=alpha, b=b, crow=crow, ncons=ncons)


def cons_slab_arrays(pb, cell0, ncl):
    """CSR arrays (1-based) of the constraint block restricted to the column slab; rows without local entries are not
    stored (new_row(), sparse_matrix.f90:266-274)."""
    sa, ija, ijl, rowptr = [], [], [1], []
    for i, (c, v) in enumerate(pb["crow"]):
        sel = (c >= cell0) & (c < cell0 + ncl)
        if sel.any():
            sa.append(v[sel]); ija.append(c[sel] - cell0 + 1)
            ijl.append(ijl[-1] + int(sel.sum())); rowptr.append(i + 1)
    return (np.concatenate(sa).astype(np.float32), np.concatenate(ija).astype(np.int32), np.array(ijl, dtype=np.int64),
            np.array(rowptr, dtype=np.int32))


def slab_arrays(pb, cell0, ncl):
    """CSR arrays (reference storage, 1-based) of S restricted to the column slab, local column indices."""
    sa, ija, ijl, rowptr = [], [], [1], []
    for i in range(pb["ndata"]):
        c, v = pb["cols"][i], pb["vals"][i]
        sel = (c >= cell0) & (c < cell0 + ncl)
        if sel.any():
            sa.append(v[sel]); ija.append(c[sel] - cell0 + 1)
            ijl.append(ijl[-1] + int(sel.sum())); rowptr.append(i + 1)
    return (np.concatenate(sa), np.concatenate(ija).astype(np.int32), np.array(ijl, dtype=np.int64),
            np.array(rowptr, dtype=np.int32))


def oracle_reference(pb, niter):
    from oracle import oracle as orc
    N, ndata = pb["N"], pb["ndata"]
    S = orc.SparseMatrix(ndata, 2 * N, sum(len(c) for c in pb["cols"]))
    for c, v in zip(pb["cols"], pb["vals"]):
        S.add_row(v, c + 1); S.new_row()
    S.finalize()
    Cm = orc.SparseMatrix(pb["ncons"], 2 * N, sum(len(c) for c, _ in pb["crow"]))
    for c, v in pb["crow"]:
        if len(c):
            Cm.add_row(v, c + 1)
        Cm.new_row()
    Cm.finalize()
    x, h, it = orc.lsqr_solve_sensit(niter, 1e-13, 0.0, 0.0, S, Cm, pb["b"], N, pb["nx"], pb["ny"], pb["nz"], 1, 0, True)
    return x[:N], h, it


def solve_nccl(pb, kind, rank, world, td, niter):
    import tomofastx_b200 as tfx
    tfx.init(int(os.environ.get("LOCAL_RANK", "0")))
    box = [tfx.comm_unique_id() if rank == 0 else None]
    td.broadcast_object_list(box, src=0)
    tfx.comm_init(world, rank, box[0])
    N, ndata = pb["N"], pb["ndata"]
    ncl = tfx.calculate_nelements_at_cpu(N, rank, world)
    cell0 = tfx.get_nsmaller(N, rank, world)
    ncol = 2 * ncl
    sa, ija, ijl, rowptr = slab_arrays(pb, cell0, ncl)
    dense = kind.startswith("dense")
    tfx.set_option("dense_detect", 1 if dense else 0)
    S = tfx.SparseMatrix.from_arrays(ndata, ncol, sa, ija, ijl, rowptr)
    tfx.set_option("dense_detect", 1)
    assert S.storage_kind() == (1 if dense else 0)
    Cm = tfx.SparseMatrix.from_arrays(pb["ncons"], ncol, *cons_slab_arrays(pb, cell0, ncl))
    u = pb["b"].copy(); x = np.zeros(ncol)
    tfx.lsqr_solve_sensit(len(u), ncol, niter, 1e-13, 0.0, 0.0, S, Cm, u, x, [1, 0], ncl, pb["nx"], pb["ny"], pb["nz"],
                          1, 0, True, myrank=rank, nbproc=world)
    h, it, fused = tfx.last_history()
    assert fused == dense
    # a collective through the C ABI on a host buffer, too
    chk = np.array([rank + 1.0, 1.0]); tfx.comm_allreduce_sum(chk, 2)
    assert chk[0] == world * (world + 1) / 2 and chk[1] == world
    tfx.comm_finalize()
    return x[:ncl], h, it, cell0, ncl


def solve_model(pb, kind, rank, world, td, niter):
    """Host model of lsqr.cu's SPLIT path for one rank (numpy products; gloo all-reduces)."""
    import torch
    import tomofastx_b200 as tfx                                      # partition helpers only (no GPU call)
    N, ndata = pb["N"], pb["ndata"]
    ncl = tfx.calculate_nelements_at_cpu(N, rank, world)
    cell0 = tfx.get_nsmaller(N, rank, world)
    sa, ija, ijl, rowptr = slab_arrays(pb, cell0, ncl)
    A = np.zeros((ndata, ncl))
    for s, row in enumerate(rowptr):
        k0, k1 = ijl[s] - 1, ijl[s + 1] - 1
        A[row - 1, ija[k0:k1] - 1] = sa[k0:k1].astype(np.float64)
    ncons = pb["ncons"]
    Cl = np.zeros((ncons, ncl))
    for i, (c, v) in enumerate(pb["crow"]):
        sel = (c >= cell0) & (c < cell0 + ncl)
        Cl[i, c[sel] - cell0] = v[sel].astype(np.float64)

    def allreduce(a):
        t = torch.from_numpy(np.ascontiguousarray(a)); td.all_reduce(t); return t.numpy()

    nlines = ndata + ncons
    u = pb["b"].copy(); x = np.zeros(ncl); hist = []
    beta = np.linalg.norm(u); u /= beta; b1 = beta
    v = A.T @ u[:ndata] + Cl.T @ u[ndata:]
    alpha = np.sqrt(allreduce(np.array([v @ v]))[0]); v /= alpha
    w = v.copy(); rhobar, phibar = alpha, beta
    it = 0
    for it in range(1, niter + 1):
        q = np.zeros(nlines)
        q[:ndata] = A @ v
        q[ndata:] = Cl @ v
        u = -alpha * u + allreduce(q)                                 # MPI_Allreduce(u), lsqr_solver2.F90:214
        beta = np.linalg.norm(u); u /= beta
        v = -beta * v + A.T @ u[:ndata] + Cl.T @ u[ndata:]
        alpha = np.sqrt(allreduce(np.array([v @ v]))[0]); v /= alpha  # normalize(), :514
        rho = np.hypot(rhobar, beta); c, s = rhobar / rho, beta / rho
        theta = s * alpha; rhobar = -c * alpha; phi = c * phibar; phibar = s * phibar
        x += (phi / rho) * w; w = -(theta / rho) * w + v
        hist.append(phibar / b1)
        if hist[-1] <= 1e-13:
            break
    return x, np.array(hist), it, cell0, ncl


def solve_model_owned(pb, kind, rank, world, td, niter):
    """Host model of csrc/lsqr.cu's row-ownership scheme (build_plan + the split path with deferred normalisation):
    a constraint row whose entries live on one rank is kept there only; rows with entries on several ranks ("shared")
    travel with the data rows; rows without entries anywhere stay on rank 0. One all-reduce of
    [q_data, q_shared, owned |u|^2] per iteration instead of the reference's whole u (lsqr_solver2.F90:214)."""
    import torch
    import tomofastx_b200 as tfx
    N, ndata, ncons = pb["N"], pb["ndata"], pb["ncons"]
    ncl = tfx.calculate_nelements_at_cpu(N, rank, world)
    cell0 = tfx.get_nsmaller(N, rank, world)
    sa, ija, ijl, rowptr = slab_arrays(pb, cell0, ncl)
    A = np.zeros((ndata, ncl))
    for s, row in enumerate(rowptr):
        k0, k1 = ijl[s] - 1, ijl[s + 1] - 1
        A[row - 1, ija[k0:k1] - 1] = sa[k0:k1].astype(np.float64)
    csa, cija, cijl, crowptr = cons_slab_arrays(pb, cell0, ncl)
    Cl = np.zeros((ncons, ncl))
    for s, row in enumerate(crowptr):
        k0, k1 = cijl[s] - 1, cijl[s + 1] - 1
        Cl[row - 1, cija[k0:k1] - 1] = csa[k0:k1].astype(np.float64)

    def allreduce(a):
        t = torch.from_numpy(np.ascontiguousarray(a)); td.all_reduce(t); return t.numpy()

    # ---- plan (build_plan): per-row count of ranks holding entries
    mine = np.zeros(ncons); mine[crowptr - 1] = 1
    cnt = allreduce(mine.copy())                  # the all-reduce works in place
    shared = np.flatnonzero(cnt >= 2)
    owned = np.flatnonzero((cnt == 1) & (mine == 1))
    if rank == 0:
        owned = np.concatenate([owned, np.flatnonzero(cnt == 0)])
    got = allreduce(np.array([float(len(owned))]))[0]
    assert got + len(shared) == ncons, "every constraint row is owned exactly once or shared"
    nq = ndata + len(shared)
    # only this rank's window of b is read (tfx_lsqr_solve_sensit with a host u)
    b = pb["b"]
    ud = b[:ndata].copy()
    uc = np.full(ncons, np.nan)                   # rows of other ranks are never touched
    uc[owned] = b[ndata + owned]; uc[shared] = b[ndata + shared]
    own2 = allreduce(np.array([np.sum(uc[owned] ** 2)]))[0]
    beta = np.sqrt(ud @ ud + np.sum(uc[shared] ** 2) + own2)
    su = 1.0 / beta; b1 = beta
    mineC = np.concatenate([owned, shared]).astype(int)
    v = su * (A.T @ ud + Cl[mineC].T @ uc[mineC])
    alpha = np.sqrt(allreduce(np.array([v @ v]))[0]); sv = 1.0 / alpha
    w = sv * v; x = np.zeros(ncl); hist = []
    rhobar, phibar = alpha, beta
    cu, cq = -alpha * su, sv
    it = 0
    for it in range(1, niter + 1):
        q = np.zeros(nq + 1)
        q[:ndata] = A @ v
        q[ndata:nq] = Cl[shared] @ v
        uc[owned] = cu * uc[owned] + cq * (Cl[owned] @ v)
        q[nq] = np.sum(uc[owned] ** 2)
        q = allreduce(q)
        ud = cu * ud + cq * q[:ndata]
        uc[shared] = cu * uc[shared] + cq * q[ndata:nq]
        beta = np.sqrt(ud @ ud + np.sum(uc[shared] ** 2) + q[nq]); su = 1.0 / beta
        v = (-beta * sv) * v + su * (A.T @ ud + Cl[mineC].T @ uc[mineC])
        alpha = np.sqrt(allreduce(np.array([v @ v]))[0]); sv = 1.0 / alpha
        rho = np.hypot(rhobar, beta); c, s_ = rhobar / rho, beta / rho
        theta = s_ * alpha; rhobar = -c * alpha; phi = c * phibar; phibar = s_ * phibar
        x += (phi / rho) * w; w = -(theta / rho) * w + sv * v
        cu, cq = -alpha * su, sv
        hist.append(phibar / b1)
        if hist[-1] <= 1e-13:
            break
    return x, np.array(hist), it, cell0, ncl


def dist_sensit_case(rank, world, td, niter):
    """Multi-GPU assembly (csrc/sensit_dist.cu): rows sharded by data, all-to-all to nnz-balanced column slabs,
    then a column-split wavelet-domain LSQR solve on the result -- against the single-rank oracle."""
    import tomofastx_b200 as tfx
    from oracle import oracle as orc
    from oracle import partition as orp
    from tests.synth import make_problem
    pb = make_problem(nx=12, ny=10, nz=6, ndata=13, compression_type=1, rate=0.2)
    N, ndata = pb.N, pb.ndata
    tfx.init(int(os.environ.get("LOCAL_RANK", "0")))
    box = [tfx.comm_unique_id() if rank == 0 else None]
    td.broadcast_object_list(box, src=0)
    tfx.comm_init(world, rank, box[0])
    rows, nnz_col, cerr, tot = tfx.sensit_assemble_rows(pb.par, pb.grid, pb.data_xyz, pb.cw, pb.dw, rank, world)
    d0, nloc, nnz_loc = rows.info()
    assert nloc == orp.calculate_nelements_at_cpu(ndata, rank, world)
    nnz_at, nel_at = tfx.get_load_balancing_nelements(nnz_col, world)
    S = tfx.sensit_repartition(rows, 1, nel_at, rank, world)
    ncl = int(nel_at[rank]); cell0 = int(nel_at[:rank].sum()); ncol = 2 * ncl
    assert S.get_number_elements() == nnz_at[rank] and S.get_ncolumns() == ncol

    # oracle: full matrix, slab by the reference's rule
    So = pb.oracle_matrix(orc)
    sa, ija, ijl, rowptr = So.arrays()
    full = {int(rowptr[i]): (ija[ijl[i] - 1:ijl[i + 1] - 1], sa[ijl[i] - 1:ijl[i + 1] - 1]) for i in range(len(rowptr))}
    want = orp.column_slab(full, N, 1, nel_at, rank, 1)
    gsa, gija, gijl, growptr = S.export()
    got = {int(growptr[i]): (gija[gijl[i] - 1:gijl[i + 1] - 1], gsa[gijl[i] - 1:gijl[i + 1] - 1]) for i in range(len(growptr))}
    bad = sum(len(set(got.get(k, ([], []))[0]) ^ set(want.get(k, ([], []))[0])) for k in set(got) | set(want))
    assert bad <= 2 * ndata, bad
    assert abs(tot - So.nel) <= 2 * ndata and abs(cerr) < 1.0
    for k in set(got) & set(want):
        if np.array_equal(got[k][0], want[k][0]):
            assert np.allclose(got[k][1], want[k][1], rtol=3e-6, atol=1e-6 * np.abs(want[k][1]).max())

    # column-split solve on the distributed matrix (damping block alpha*I, wavelet domain)
    alpha = np.float32(1e-3)
    Cm = tfx.SparseMatrix.from_arrays(N, ncol, np.full(ncl, alpha, dtype=np.float32), np.arange(1, ncl + 1, dtype=np.int32),
                                      np.arange(1, ncl + 2, dtype=np.int64),
                                      np.arange(cell0 + 1, cell0 + ncl + 1, dtype=np.int32))
    b = np.concatenate([pb.rhs(So, orc), np.zeros(N)])
    u = b.copy(); x = np.zeros(ncol)
    tfx.lsqr_solve_sensit(len(u), ncol, niter, 1e-13, 0.0, 0.0, S, Cm, u, x, [1, 0], ncl, pb.nx, pb.ny, pb.nz, 1, 1, True,
                          myrank=rank, nbproc=world)
    h, it, fused = tfx.last_history()

    # ---- distributed wavelet (wavelet_utils.F90:37-72) and calculate_data (model.F90:220-307) on the slabs
    rng = np.random.default_rng(11)
    vol = rng.standard_normal(N)
    mine = vol[cell0:cell0 + ncl].copy()
    want_w = orc.forward_wavelet(vol.copy(), pb.nx, pb.ny, pb.nz, 1)
    want_i = orc.inverse_wavelet(want_w.copy(), pb.nx, pb.ny, pb.nz, 1)
    # the layout changes through peer memory (cudaIpc, default), through NCCL send/recv, and the all-gather fallback;
    # twice each (the second transform re-uses the mapped buffers: ordering between consecutive transforms)
    for p2p, dist in ((1, 1), (1, 1), (0, 1), (0, 0), (1, 1)):
        tfx.set_option("wavelet_p2p", p2p); tfx.set_option("wavelet_dist", dist)
        mine = vol[cell0:cell0 + ncl].copy()
        tfx.apply_wavelet_transform(ncl, pb.nx, pb.ny, pb.nz, 1, mine, True, 1, 1, [1], rank, world)
        assert np.array_equal(mine, want_w[cell0:cell0 + ncl]), "distributed Haar must be bit-identical to the serial one"
        tfx.apply_wavelet_transform(ncl, pb.nx, pb.ny, pb.nz, 1, mine, False, 1, 1, [1], rank, world)
        assert np.array_equal(mine, want_i[cell0:cell0 + ncl])
    tfx.set_option("wavelet_p2p", 1); tfx.set_option("wavelet_dist", 1)
    dwt = np.linspace(0.5, 1.5, ndata).reshape(ndata, 1)
    d_got = tfx.calculate_data(S, pb.m_true[:, cell0:cell0 + ncl], ndata, 1, 1.0, pb.cw[cell0:cell0 + ncl], dwt, 1,
                               pb.nx, pb.ny, pb.nz, 1, 0, rank, world)
    d_want = orc.calculate_data(So, pb.m_true, ndata, 1, 1.0, pb.cw, dwt, 1, pb.nx, pb.ny, pb.nz, 1, 0)
    assert np.allclose(d_got, d_want, rtol=1e-5, atol=1e-7 * np.abs(d_want).max())

    # ---- the same slab assembled in two station batches as row blocks (option sensit_row_blocks): same solve
    import copy
    tfx.set_option("sensit_row_blocks", 1)
    Sb = tfx.SparseMatrix(ndata, ncol, int(nnz_at[rank]) + 64)
    xs_, ys_, zs_ = pb.data_xyz
    for d0b, nb in ((0, 6), (6, ndata - 6)):
        parb = copy.copy(pb.par); parb.ndata = nb
        slb = slice(d0b, d0b + nb)
        rows_b, _, _, _ = tfx.sensit_assemble_rows(parb, pb.grid, (xs_[slb], ys_[slb], zs_[slb]), pb.cw, pb.dw[slb], rank, world)
        tfx.sensit_repartition_into(Sb, rows_b, 1, nel_at, rank, world)
    Sb.finalize()
    tfx.set_option("sensit_row_blocks", 0)
    assert Sb.get_number_elements() == S.get_number_elements()
    ub = b.copy(); xb = np.zeros(ncol)
    tfx.lsqr_solve_sensit(len(ub), ncol, niter, 1e-13, 0.0, 0.0, Sb, Cm, ub, xb, [1, 0], ncl, pb.nx, pb.ny, pb.nz, 1, 1, True,
                          myrank=rank, nbproc=world)
    hb, itb, _ = tfx.last_history()
    nh = min(10, len(h))
    assert itb == it and np.allclose(hb[:nh], h[:nh], rtol=1e-9), (hb[:nh], h[:nh])

    # ---- constraint producers and depth weight on the slabs (csrc/cons.cu, csrc/weights.cu) vs the oracle's slab mode
    def same(Sg, Sref):
        a, b_ = Sg.export(), Sref.arrays()
        return all(np.array_equal(p_, q_) for p_, q_ in zip(a, b_))
    dX, dY, dZ = np.full(pb.nx, 100.0), np.full(pb.ny, 100.0), np.full(pb.nz, 50.0)
    m1 = rng.standard_normal(N); m2 = rng.standard_normal(N); lw = rng.uniform(0.5, 1.5, N); cw2 = rng.uniform(0.5, 2.0, N)
    sl = slice(cell0, cell0 + ncl)
    nlc = 5 * N
    Cg, Co = tfx.SparseMatrix(nlc, 2 * ncl, 30 * N), orc.SparseMatrix(nlc, 2 * ncl, 30 * N)
    bg, bo = np.zeros(nlc), np.zeros(nlc)
    c1 = tfx.damping_add(Cg, bg, 1e-2, 0.9, 2.0, 1, pb.nx, pb.ny, pb.nz, pb.cw[sl], m1[sl], m2[sl], 0, True, lw[sl], rank, world)
    c1o = orc.damping_add(Co, bo, 1e-2, 0.9, 2.0, 1, pb.nx, pb.ny, pb.nz, cell0, ncl, pb.cw, m1, m2, 0, True, lw)
    c2 = tfx.damping_gradient_add(Cg, bg, 3e-2, 0.9, pb.nx, pb.ny, pb.nz, dX, dY, dZ, m1, pb.cw[sl], lw, 0, 2, rank, world)
    c2o = orc.damping_gradient_add(Co, bo, 3e-2, 0.9, pb.nx, pb.ny, pb.nz, dX, dY, dZ, cell0, ncl, m1, pb.cw, lw, 0, 2)
    c3, cg_g = tfx.cross_gradient_calculate(Cg, bg, pb.nx, pb.ny, pb.nz, dX, dY, dZ, m1, m2, pb.cw[sl], cw2[sl], 1, 0.5,
                                            (0, 0), rank, world)
    c3o, cg_o, _, _ = orc.cross_gradient_calculate(Co, bo, pb.nx, pb.ny, pb.nz, dX, dY, dZ, cell0, ncl, m1, m2, pb.cw, cw2, 1, 0.5)
    Cg.finalize(); Co.finalize()
    assert same(Cg, Co), "device-built constraint slab differs from the oracle's"
    assert np.array_equal(bg, bo) and np.array_equal(cg_g, cg_o)
    assert np.allclose([c1, c2] + list(c3), [c1o, c2o] + list(c3o), rtol=1e-13)
    xyz = pb.data_xyz
    dwg = tfx.calculate_depth_weight(2, pb.grid, xyz, 3.0, 1.5, 0.0, cell0, ncl, rank, world)
    dwo = orc.depth_weight(2, pb.grid, xyz[0], xyz[1], xyz[2], 3.0, 1.5, 0.0)
    assert np.allclose(dwg, dwo[sl], rtol=1e-13), "distance weighting: the normalisation maximum must be global"

    # ---- the same system solved in the physical domain: wavelet transforms inside the loop (WAVELET_DOMAIN = F,
    # lsqr_solver2.F90:200-207,228-235) on the distributed vectors
    u2 = b.copy(); x2 = np.zeros(ncol)
    tfx.lsqr_solve_sensit(len(u2), ncol, niter, 1e-13, 0.0, 0.0, S, Cm, u2, x2, [1, 0], ncl, pb.nx, pb.ny, pb.nz, 1, 1, False,
                          myrank=rank, nbproc=world)
    h2, it2, _ = tfx.last_history()
    tfx.comm_finalize()
    parts = [None] * world
    td.all_gather_object(parts, (cell0, ncl, x[:ncl]))
    if rank == 0:
        xg = np.zeros(N)
        for c0, n, xl in parts:
            xg[c0:c0 + n] = xl
        Co = orc.SparseMatrix(N, 2 * N, N)
        for p in range(N):
            Co.add(float(alpha), p + 1); Co.new_row()
        Co.finalize()
        x_ref, h_ref, it_ref = orc.lsqr_solve_sensit(niter, 1e-13, 0.0, 0.0, So, Co, b, N, pb.nx, pb.ny, pb.nz, 1, 1, True)
        assert it == it_ref
        n = min(8, len(h_ref))
        # threshold flips (<= a couple of entries per row) perturb the matrix at the 1e-4 level of a row's norm
        assert np.allclose(h[:n], h_ref[:n], rtol=5e-3), (h[:n], h_ref[:n])
        x_ref2, h_ref2, it_ref2 = orc.lsqr_solve_sensit(niter, 1e-13, 0.0, 0.0, So, Co, b, N, pb.nx, pb.ny, pb.nz, 1, 1, False)
        assert it2 == it_ref2 and np.allclose(h2[:n], h_ref2[:n], rtol=5e-3), (h2[:n], h_ref2[:n])
        print("multi_rank_case ok: backend=nccl kind=dist_sensit world=%d nnz=%d slabs=%s iters=%d r_last=%.6e" %
              (world, tot, list(map(int, nel_at)), it, h[-1]), flush=True)


def dist_sensit_model(rank, world, td):
    """Host model of csrc/sensit_dist.cu on CPU (gloo): every rank owns the oracle's rows of its stations
    (even data split), cuts each row at the slab boundaries, sends the pieces to the column owners and
    concatenates what it receives in source-rank order; the result must equal the reference rule applied to
    the full matrix, already sorted by (row, column)."""
    import tomofastx_b200 as tfx                                      # host-only helpers (no GPU call)
    from oracle import oracle as orc
    from oracle import partition as orp
    from tests.synth import make_problem
    pb = make_problem(nx=8, ny=6, nz=4, ndata=7, compression_type=1, rate=0.25, problem_type=2, nmodel_components=3)
    N, nmc = pb.N, 3
    pb.par.param_shift = 0
    So = pb.oracle_matrix(orc)
    sa, ija, ijl, rowptr = So.arrays()
    full = {int(rowptr[i]): (ija[ijl[i] - 1:ijl[i + 1] - 1], sa[ijl[i] - 1:ijl[i + 1] - 1]) for i in range(len(rowptr))}
    nnz_col = np.zeros(N, dtype=np.int32)
    for c, _ in full.values():
        np.add.at(nnz_col, (c - 1) % N, 1)
    _, nel_at = tfx.get_load_balancing_nelements(nnz_col, world)
    assert np.array_equal(nel_at, orp.get_load_balancing_nelements(nnz_col, world)[1])
    d0 = tfx.get_nsmaller(pb.ndata, rank, world)
    nloc = tfx.calculate_nelements_at_cpu(pb.ndata, rank, world)
    mine = {r: full[r] for r in range(d0 + 1, d0 + nloc + 1) if r in full}
    # pack: destination-major, row order inside a destination (k_pack_pieces)
    send = []
    for dst in range(world):
        piece = orp.column_slab(mine, N, nmc, nel_at, dst, 2)
        send.append([(r, piece[r][0], piece[r][1]) for r in sorted(piece)])
    allsend = [None] * world
    td.all_gather_object(allsend, send)
    recv = [t for src in range(world) for t in allsend[src][rank]]     # source-rank order
    rows_seq = [t[0] for t in recv]
    assert rows_seq == sorted(rows_seq), "pieces must arrive in global row order"
    want = orp.column_slab(full, N, nmc, nel_at, rank, 2)
    assert [t[0] for t in recv] == sorted(want)
    for r, c, v in recv:
        assert np.array_equal(c, want[r][0]) and np.array_equal(v, want[r][1]) and np.all(np.diff(c) > 0)
        assert c.min() >= 3 * nel_at[rank] + 1 and c.max() <= 6 * nel_at[rank]       # problem 2 of a 3-component model
    print("multi_rank_case ok: backend=model kind=dist_sensit world=%d slabs=%s" % (world, list(map(int, nel_at))), flush=True)


def dist_wavelet_model(rank, world, td):
    """Host model of csrc/data.cu:wavelet_slab_dist on CPU (gloo): the 3-D transform of a volume held as contiguous cell
    slabs WITHOUT assembling it anywhere -- layout A (whole k-planes: axes 1 and 2 local), one all-to-all, layout B (all k
    for a range of plane cells: axis 3 local), one all-to-all back to the slabs. Same plan arithmetic (ka, pa, the
    contiguous B range of a slab) as the C++; the per-axis passes are the oracle's 3-D transform applied to shapes with
    the other axes' lengths 1. Result must be bit-identical to the oracle's transform of the whole volume."""
    from oracle import oracle as orc
    rng = np.random.default_rng(99)
    for (nx, ny, nz), wtype in (((12, 10, 9), 1), ((7, 9, 11), 2), ((16, 8, 6), 1)):
        plane, N = nx * ny, nx * ny * nz
        vol = rng.standard_normal(N)
        # unequal slabs (nnz-balanced in the product): cut points not aligned with planes
        cuts = np.sort(rng.choice(np.arange(plane, N - plane), size=world - 1, replace=False)) if world > 1 else np.array([], int)
        off = np.concatenate([[0], cuts, [N]]).astype(np.int64)
        ka = [(int(off[r]) + plane - 1) // plane for r in range(world)] + [nz]
        pa = [plane * r // world for r in range(world + 1)]
        qualifies = all(ka[r] < ka[r + 1] for r in range(world)) and all(
            (ka[r] - 1) * plane >= off[r - 1] for r in range(1, world))
        for fwd in (True, False):
            want = (orc.forward_wavelet if fwd else orc.inverse_wavelet)(vol.copy(), nx, ny, nz, wtype)
            if not qualifies:
                continue                                   # the product falls back to the gather path
            t = orc.forward_wavelet if fwd else orc.inverse_wavelet
            slab = vol[off[rank]:off[rank + 1]].copy()
            me = rank
            nk = ka[me + 1] - ka[me]
            lead = ka[me] * plane - int(off[me])
            # slabs -> A: my planes = my cells from `lead` on + the leading piece of rank me+1's slab
            pieces = [None] * world
            td.all_gather_object(pieces, slab[:lead])
            A = np.concatenate([slab[lead:], pieces[me + 1] if me + 1 < world else np.zeros(0)])
            assert A.size == nk * plane
            # axes 1 and 2 on my planes: a (nx, ny, nk) volume transformed along axis 1, then along axis 2
            A = A.reshape(nk, ny, nx)
            for k in range(nk):
                for j in range(ny):
                    A[k, j] = t(A[k, j].copy(), nx, 1, 1, wtype)
                for i in range(nx):
                    A[k, :, i] = t(A[k, :, i].copy(), 1, ny, 1, wtype)
            A = A.reshape(nk, plane)
            # A -> B: rank q receives the columns [pa[q], pa[q+1]) of everybody's planes, plane order = rank order
            send = [A[:, pa[q]:pa[q + 1]].copy() for q in range(world)]
            allsend = [None] * world
            td.all_gather_object(allsend, send)
            B = np.concatenate([allsend[r][me] for r in range(world)], axis=0)      # (nz, W)
            W = pa[me + 1] - pa[me]
            assert B.shape == (nz, W)
            for p in range(W):
                B[:, p] = t(B[:, p].copy(), 1, 1, nz, wtype)
            # B -> slabs: the cells of slab r inside my columns are ONE contiguous range of B
            Bf = B.ravel()

            def rng_in(q, r):
                Wq = pa[q + 1] - pa[q]
                k0, k1 = int(off[r]) // plane, (int(off[r + 1]) - 1) // plane
                s0, e1 = int(off[r]) - k0 * plane, int(off[r + 1]) - k1 * plane
                b0 = k0 * Wq + min(max(s0, pa[q]), pa[q + 1]) - pa[q]
                b1 = k1 * Wq + min(max(e1, pa[q]), pa[q + 1]) - pa[q]
                return b0, max(b0, b1)
            send = [Bf[slice(*rng_in(me, r))].copy() for r in range(world)]
            td.all_gather_object(allsend, send)
            # peer-memory path (csrc/data.cu k_scatter_ranges): the SENDER computes where its piece starts in rank r's
            # staging buffer -- behind the pieces of the ranks before it, rank r itself sending nothing; the receiver
            # unpacks the pieces in rank order. Both views must agree.
            for r in range(world):
                if r == me:
                    continue
                sender_off = sum(rng_in(s_, r)[1] - rng_in(s_, r)[0] for s_ in range(me) if s_ != r)
                recv_off = sum(allsend[s_][r].size for s_ in range(me) if s_ != r)
                assert sender_off == recv_off, (sender_off, recv_off)
            out = np.full(off[me + 1] - off[me], np.nan)
            k0, k1 = int(off[me]) // plane, (int(off[me + 1]) - 1) // plane
            for q in range(world):
                msg = allsend[q][me]
                pos = 0
                for k in range(k0, k1 + 1):
                    lo = max(pa[q], int(off[me]) - k * plane if k == k0 else 0)
                    hi = min(pa[q + 1], int(off[me + 1]) - k * plane if k == k1 else plane)
                    if hi > lo:
                        d0 = k * plane + lo - int(off[me])
                        out[d0:d0 + hi - lo] = msg[pos:pos + hi - lo]
                        pos += hi - lo
                assert pos == msg.size
            assert np.array_equal(out, want[off[me]:off[me + 1]]), ((nx, ny, nz), wtype, fwd)
    print("multi_rank_case ok: backend=model kind=dist_wavelet world=%d" % world, flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--backend", default="nccl", choices=["nccl", "model"])
    ap.add_argument("--niter", type=int, default=30)
    a = ap.parse_args()
    import torch
    import torch.distributed as td
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    td.init_process_group(backend="gloo", rank=rank, world_size=world)
    ok = True
    cases = [(k, solve_nccl) for k in ("dense", "sparse", "xgrad", "dense_xgrad")] if a.backend == "nccl" else \
            [("dense", solve_model), ("sparse", solve_model), ("xgrad", solve_model), ("xgrad", solve_model_owned),
             ("sparse", solve_model_owned)]
    for kind, fn in cases:
        pb = full_problem(kind)
        x_loc, h, it, cell0, ncl = fn(pb, kind, rank, world, td, a.niter)
        parts = [None] * world
        td.all_gather_object(parts, (cell0, ncl, x_loc))
        if rank == 0:
            x = np.zeros(pb["N"])
            covered = np.zeros(pb["N"], dtype=int)
            for c0, n, xl in parts:
                x[c0:c0 + n] = xl; covered[c0:c0 + n] += 1
            assert np.all(covered == 1), "column slabs must tile the model exactly once"
            x_ref, h_ref, it_ref = oracle_reference(pb, a.niter)
            assert it == it_ref, (it, it_ref)
            n = min(10, len(h_ref))
            assert np.allclose(h[:n], h_ref[:n], rtol=1e-6), (kind, h[:n], h_ref[:n])
            assert abs(h[-1] - h_ref[-1]) <= 1e-6 * h_ref[-1], (kind, h[-1], h_ref[-1])
            assert np.allclose(x, x_ref, rtol=1e-6, atol=1e-6 * np.abs(x_ref).max()), (kind, fn.__name__, np.abs(x - x_ref).max(), np.abs(x_ref).max())
            print("multi_rank_case ok: backend=%s kind=%s world=%d iters=%d r_last=%.6e" % (a.backend, kind, world, it, h[-1]),
                  flush=True)
    if a.backend == "nccl":
        dist_sensit_case(rank, world, td, a.niter)
    else:
        dist_sensit_model(rank, world, td)
        dist_wavelet_model(rank, world, td)
    td.barrier()
    td.destroy_process_group()
    if not ok:
        sys.exit(1)


if __name__ == "__main__":
    main()
