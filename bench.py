#!/usr/bin/env python
"""bench.py -- LSQR iterations/s on the Tomofast-x inversion hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repository (libtfx, sm_100a CUDA)
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the host cores

Workload (config.workload): BASELINE.json configs[1] -- synthetic gravity inversion, 256 x 256 x 64
cells, 10 000 stations, no compression, assembled ON the device with the gravity prism kernel
(SURVEY.md section 8d inputs: regular 100 x 100 x 50 m cells, stations on a lattice at z = -0.1, depth
weighting type 1, one 250 kg/m3 block as true model, Parfile-default model damping 1e-11).
A "step" is one LSQR iteration of lsqr_solve_sensit: both products with S (and with the damping block),
the norms and the x/w updates. With N > 1 the SAME problem is column-sharded over the ranks like the
reference's MPI decomposition (lsqr_solver2.F90:16) -> "scaling": "strong".

value  : iterations/s with u, x and S resident in HBM, timed with CUDA events on the library stream
         around the iteration loop (max over ranks).
e2e    : iterations/s through the C ABI call tfx_lsqr_solve_sensit with HOST buffers (pinned): H2D of the
         right-hand side, the initialisation before the loop, K iterations, D2H of x and u.
roofline: the fused sweep kernel (dense_sweep_kernel): ALGORITHMIC bytes (4 B per matrix entry read
         once + the vectors) / mean launch time from CUDA events, against the measured HBM peak.
spmv   : (extra object) the compressed half of BASELINE.json's metric on BASELINE config C's shape: 512 x 512 x 128
         cells, Haar compression 5 %, 6 250 stations PER GPU (weak scaling: 8 GPUs = the 50 000 stations of config C),
         assembled on the device in row blocks, kept in the T16 layouts (12 B/nnz for both products, ~126 GB per GPU):
         S x / S^T u GB/s by the reference's 8 B/nnz accounting and by the 6 B/nnz the layouts move, wavelet transform
         times, assembly rate, compressed LSQR it/s. --comp-grid / --comp-ndata / --comp-batch change the shape.
config_d: (extra object) parfiles/Parfile_2body_induced.txt with Daubechies-4 compression end to end (distance weights,
         magnetic 3-component assembly, 2 x 100 LSQR iterations), column slabs over the N ranks.
config_e: (extra object, --gpus >= 2 or --config-e) joint gravity + magnetic inversion with cross-gradient constraints
         and the wavelet transforms inside the LSQR loop (see tomofast-x_b200/configs.py).
cpu_baseline / --impl reference: the oracle port of the reference loops (sparse_matrix.f90:313-405,
         lsqr_solver2.F90:321-473) on the host cores, column-split over P processes like the reference's
         MPI ranks, on a bounded column sample of the same matrix shape, scaled linearly in nnz.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "LSQR iterations/sec"
UNIT = "it/s"
FALLBACK_HBM_GBS = 6650.0          # /opt/skills/guides/B200_PROFILING.md fallback


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


# ----------------------------------------------------------------------------------------------------
# clocks sampling during the timed region
# ----------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.FIELDS,
                                       "--format=csv,noheader,nounits", "-lms", "200"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.f:
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1])); mx.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"),
                                 parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.f.name)
        if sm:
            out["sm_mhz"] = float(np.median(sm))
            out["sm_max_mhz"] = float(max(mx))
            out["samples"] = len(sm)
        out["reasons"] = sorted(reasons)
        return out


# ----------------------------------------------------------------------------------------------------
# distributed plumbing (torchrun env); data-path collectives are NCCL inside libtfx
# ----------------------------------------------------------------------------------------------------
class Dist:
    def __init__(self):
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.td = None
        if self.world > 1:
            import torch.distributed as td
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            td.init_process_group(backend="gloo", rank=self.rank, world_size=self.world)
            self.td = td

    def barrier(self):
        if self.td:
            self.td.barrier()

    def bcast_obj(self, obj):
        if not self.td:
            return obj
        box = [obj]
        self.td.broadcast_object_list(box, src=0)
        return box[0]

    def max(self, x):
        if not self.td:
            return x
        import torch
        t = torch.tensor([float(x)], dtype=torch.float64)
        self.td.all_reduce(t, op=self.td.ReduceOp.MAX)
        return float(t[0])

    def min(self, x):
        return -self.max(-x)

    def sum(self, x):
        if not self.td:
            return x
        import torch
        t = torch.tensor([float(x)], dtype=torch.float64)
        self.td.all_reduce(t, op=self.td.ReduceOp.SUM)
        return float(t[0])

    def finish(self):
        if self.td:
            self.td.barrier()
            self.td.destroy_process_group()


# ----------------------------------------------------------------------------------------------------
# workload
# ----------------------------------------------------------------------------------------------------
def workload_name(a):
    return "synthetic gravity inversion %dx%dx%d cells, %d data, no compression" % (a.nx, a.ny, a.nz, a.ndata)


def run_ours(a):
    import tomofastx_b200 as tfx
    from tomofastx_b200.synth import depth_weight_type1, regular_grid, station_lattice

    d = Dist()
    tfx.init(d.local_rank)
    for kv in a.opt:
        k, v = kv.split("=")
        tfx.set_option(k, int(v))
    if a.dense_vec4 is not None:
        tfx.set_option("dense_vec4", a.dense_vec4)
    if a.dense_f2f_rows is not None:
        tfx.set_option("dense_f2f_rows", a.dense_f2f_rows)
    if d.world > 1:
        uid = d.bcast_obj(tfx.comm_unique_id() if d.rank == 0 else None)
        tfx.comm_init(d.world, d.rank, uid)

    if a.no_dense:
        # the extras alone (compressed section on config C's shape, config D, config E); not the driver's bench line
        extras = run_extras(a, tfx, d)
        if d.rank == 0:
            v = extras.get("spmv", {}).get("lsqr", {}).get("it_per_s")
            print(json.dumps(dict({"metric": METRIC, "unit": UNIT, "n_gpus": d.world, "value": v,
                                   "note": "extras only (--no-dense)"}, **extras)), flush=True)
        d.finish()
        return None
    nx, ny, nz, ndata = a.nx, a.ny, a.nz, a.ndata
    N = nx * ny * nz
    ncl = tfx.calculate_nelements_at_cpu(N, d.rank, d.world)          # column slab of this rank
    cell0 = tfx.get_nsmaller(N, d.rank, d.world)
    ncolumns = 2 * ncl                                                # joint_inverse_problem.F90:213-214
    grid = regular_grid(nx, ny, nz)
    data_xyz = station_lattice(ndata, 100.0 * nx, 100.0 * ny, z=-0.1)
    cw = depth_weight_type1(grid, 2.0, 0.0, 4.0e3)
    dw = np.ones((ndata, 1))
    problem_weight, alpha_damp = 1.0, 1.0e-11                         # parameters_init.f90:339-345,329

    par = tfx.SensitParams()
    par.problem_type = 1
    par.nx, par.ny, par.nz = nx, ny, nz
    par.ndata, par.ndata_components, par.nmodel_components, par.data_type = ndata, 1, 1, 1
    par.compression_type, par.compression_rate = 0, 1.0
    par.problem_weight = problem_weight
    par.cell0, par.ncells_local, par.param_shift, par.ncolumns = cell0, ncl, 0, ncolumns

    free_b, total_b = tfx.device_mem_info()
    need = ((ndata + 3) // 4 * 4) * ncl * 4 + 6 * ncolumns * 8 + 7 * N * 8 + (1 << 30)
    if need > free_b:
        raise SystemExit("bench: the %s block needs %.1f GB of HBM, only %.1f GB free" %
                         (workload_name(a), need / 1e9, free_b / 1e9))

    t0 = time.perf_counter()
    S, _, _, nnz_loc = tfx.calculate_sensit(par, grid, data_xyz, cw, dw)
    tfx.synchronize()
    t_assemble = time.perf_counter() - t0
    assert S.storage_kind() == 1

    # model damping block alpha*I (damping.F90:158-179): N rows, this rank stores its own ncl rows
    sa = np.full(ncl, alpha_damp * problem_weight, dtype=np.float32)
    C = tfx.SparseMatrix.from_arrays(N, ncolumns, sa, np.arange(1, ncl + 1, dtype=np.int32),
                                     np.arange(1, ncl + 2, dtype=np.int64),
                                     np.arange(cell0 + 1, cell0 + ncl + 1, dtype=np.int32))

    # observed data = S * (m_true / cw) summed over the column slabs (model.F90:243-293); start model 0
    m = np.zeros((nz, ny, nx))
    sl = lambda n: slice(max(0, n // 2 - max(1, n // 8)), n // 2 + max(1, n // 8))
    m[sl(nz), sl(ny), sl(nx)] = 250.0
    xs = np.zeros(ncolumns)
    xs[:ncl] = (m.ravel() / cw)[cell0:cell0 + ncl]
    d_obs = S.mult_vector(xs)
    if d.world > 1:
        tfx.comm_allreduce_sum(d_obs, ndata)
    nlines = ndata + N
    b = np.zeros(nlines)
    b[:ndata] = problem_weight * d_obs                                # calculate_b_RHS; damping RHS is 0 (m = m_prior)
    del grid, xs, m

    def solve(u, x, niter):
        tfx.lsqr_solve_sensit(nlines, ncolumns, niter, 1.0e-13, 0.0, 0.0, S, C, u, x, [1, 0], ncl, nx, ny, nz, 1,
                              0, True, myrank=d.rank, nbproc=d.world)

    # ---- device-resident measurement ("value")
    u_dev, x_dev = tfx.Buffer(nlines), tfx.Buffer(ncolumns)
    tfx.set_option("profile_sweeps", 1)
    tfx.copy(u_dev, b, nlines)
    solve(u_dev, x_dev, a.warmup)                                     # W untimed warm-up iterations
    tfx.copy(u_dev, b, nlines)
    tfx.synchronize()
    d.barrier()
    sampler = ClockSampler(d.local_rank)
    if d.rank == 0:
        sampler.start()
    l0 = tfx.launch_count()
    solve(u_dev, x_dev, a.steps)
    tfx.synchronize()
    launches = tfx.launch_count() - l0
    loop_ms, sweep_ms, nsweeps = tfx.last_timing()
    hist, iters, fused = tfx.last_history()
    d.barrier()
    clocks = sampler.stop() if d.rank == 0 else None
    loop_ms_max = d.max(loop_ms)
    assert iters == a.steps and fused, (iters, fused)
    # the fused sweep is launched once before the loop and once per iteration
    sweep_avg_ms = d.max(sweep_ms / max(nsweeps, 1))

    # ---- end to end through the C ABI with pinned HOST buffers ("e2e")
    tfx.set_option("profile_sweeps", 0)
    u_pin, x_pin = tfx.Buffer(nlines, "pinned"), tfx.Buffer(ncolumns, "pinned")
    e2e_runs = []
    for _ in range(2):          # first call also pays the one-time staging-buffer allocation; the second is reported
        u_pin.numpy()[:] = b
        d.barrier()
        t0 = time.perf_counter()
        solve(u_pin, x_pin, a.steps)
        e2e_runs.append(d.max(time.perf_counter() - t0))
    e2e_s = e2e_runs[-1]
    x_host = x_pin.numpy().copy()
    d.barrier()

    if d.rank == 0:
        peak, peak_src = hbm_peak()
        # algorithmic bytes of one fused sweep launch: every matrix entry once (4 B) + v, g, vhat (8 B each per
        # column) + u (8 B per row) + the per-CTA partial q vectors
        nnz_loc_f = float(((ndata + 3) // 4 * 4)) * ncl
        alg_bytes = 4.0 * nnz_loc_f + 24.0 * ncl + 8.0 * ndata + 8.0 * ndata * 148
        achieved = alg_bytes / (sweep_avg_ms * 1e-3) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(tp) and d.world == 1:
            try:
                tj = json.load(open(tp))
                if tj.get("workload") == workload_name(a):
                    traffic = tj.get("dram_bytes_per_launch")
            except Exception:
                pass
        value = a.steps / (loop_ms_max * 1e-3)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": d.world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": loop_ms_max / a.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64 (f32 matrix values, f64 vectors/accumulation)", "data": "synthetic",
            "config": {"workload": workload_name(a), "parallelism": "column slabs x%d" % d.world,
                       "matrix_bytes_per_gpu": int(nnz_loc_f * 4), "l2_policy": "inputs larger than L2 (matrix "
                       "%.1f GB per GPU streamed every step)" % (nnz_loc_f * 4 / 1e9),
                       "lsqr": "fused single-sweep path, damping block alpha=1e-11, rmin=1e-13",
                       "assemble_s": round(t_assemble, 2), "residual_last": float(hist[-1]) if len(hist) else None},
            "e2e": {"value": a.steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(nlines * 8 / a.steps),
                    "d2h_bytes_per_step": int((nlines + ncolumns) * 8 / a.steps),
                    "first_call_value": a.steps / e2e_runs[0],
                    "note": "tfx_lsqr_solve_sensit with pinned HOST u/x: H2D of the right-hand side, the "
                            "initialisation before the loop, K iterations, D2H of x and u; wall clock around the call"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "kernel": "dense_sweep_kernel<K,FUSED>", "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg_bytes, "launch_ms": sweep_avg_ms,
                         "note": "one launch = S^T u AND S vhat on one read of S; by the reference's 16 B/nnz-per-"
                                 "iteration CSR accounting the same launch is worth %.0f GB/s" %
                                 (16.0 * nnz_loc_f / (sweep_avg_ms * 1e-3) / 1e9)},
        }
        if d.world == 1 and not a.no_cpu_baseline:
            line["cpu_baseline"] = cpu_reference(a, steps=2, warmup=1, seconds_hint=20.0)
    # ---- extras: the other BASELINE configurations. A failure of an extra never costs the headline line.
    del S, C, u_dev, x_dev, u_pin, x_pin               # free the dense block first
    import gc
    gc.collect()
    extras = run_extras(a, tfx, d)
    if d.rank == 0:
        line.update(extras)
        print(json.dumps(line), flush=True)
    d.finish()
    return x_host


def run_extras(a, tfx, d):
    import traceback
    out = {}

    def guarded(name, fn):
        try:
            res = fn()
        except Exception as e:                          # noqa: BLE001 -- reported in the JSON line
            res = {"error": "%s: %s" % (type(e).__name__, str(e)[:300])}
            sys.stderr.write("bench extra %s failed:\n%s\n" % (name, traceback.format_exc()))
        failed = d.max(1.0 if "error" in res else 0.0) > 0
        if failed and "error" not in res:
            res = {"error": "failed on another rank"}
        out[name] = res
        tfx.synchronize()
        import gc
        gc.collect()

    if not a.no_compressed:
        guarded("spmv", lambda: compressed_spmv(a, tfx, d))
    if not a.no_config_d:
        guarded("config_d", lambda: config_d_extra(a, tfx, d))
    if (a.config_e == 1) or (a.config_e < 0 and d.world >= 2):
        guarded("config_e", lambda: config_e_extra(a, tfx, d))
    return out


def config_d_extra(a, tfx, d):
    """BASELINE config D: parfiles/Parfile_2body_induced.txt with Daubechies-4 compression, end to end."""
    from tomofastx_b200 import configs
    fixture = os.path.join(ROOT, "tests", "golden", "twobody_induced.npz")   # input DATA of the reference's Parfile
    if not os.path.exists(fixture):
        return {"skipped": "input fixture tests/golden/twobody_induced.npz not found"}
    c = configs.load_twobody(fixture)
    configs.run_config_d(tfx, dict(c, nmajor=1, niter=3), 2, d.rank, d.world, sync=d.barrier)   # warm-up (allocations, NCCL)
    r = configs.run_config_d(tfx, c, 2, d.rank, d.world, sync=d.barrier)
    loop_ms = d.max(r["loop_ms"])
    nnz = r["nnz"]
    peak, _ = hbm_peak()
    return {"workload": "Parfile_2body_induced: magnetic 3-component, 67x67x30 cells, 1681 data, Daubechies-4 rate 0.3, "
                        "distance weighting, 2 x 100 LSQR iterations, damping 1e-8",
            "n_gpus": d.world, "nnz": nnz, "compression_error": r.get("compression_error"),
            "depth_weight_s": round(d.max(r["depth_weight_s"]), 4), "assemble_s": round(d.max(r["assemble_s"]), 3),
            "inversion_s": round(d.max(r["inversion_s"]), 3), "iters": r["iters"],
            "lsqr_it_per_s": r["iters"] / (loop_ms * 1e-3), "ms_per_it": loop_ms / max(r["iters"], 1),
            "ref_accounting_gbs": 16.0 * nnz * r["iters"] / loop_ms / 1e6, "ref_accounting_frac": 16.0 * nnz * r["iters"] / loop_ms / 1e6 / (peak * d.world),
            "costs": [float(v) for v in r["costs"]], "residual_last": [float(h[-1]) for h in r["histories"]],
            "column_slabs": r.get("column_slabs"),
            "note": "launch/latency-bound at this size (2.0e8 nnz, SURVEY 8d: 3.3 GB per iteration by the reference's "
                    "accounting); parity vs the oracle: tests/test_gpu_config_d.py"}


def config_e_extra(a, tfx, d):
    """BASELINE config E in shape: joint gravity + magnetic inversion, cross-gradient constraint, wavelet transforms
    inside the LSQR loop, column slabs over the N ranks. The grid is weak-scaled with N (16.8 M cells per GPU:
    1024 x 1024 x 128 at 8 GPUs = config E's grid); the station count is a small fraction of config E's 200 000 (the
    assembly of the full kernel, 5.4e13 cell evaluations, does not fit a benchmark run), so the products with S are
    timed on their own and the full-size iteration is PROJECTED from them."""
    from tomofastx_b200 import configs
    grids = {1: (256, 512, 128), 2: (512, 512, 128), 4: (1024, 512, 128)}
    nx, ny, nz = a.e_grid if a.e_grid else grids.get(d.world, (1024, 1024, 128))
    per = a.e_ndata if a.e_ndata > 0 else 500
    nd1 = nd2 = per * d.world
    rate = 0.002
    r = configs.run_config_e(tfx, nx, ny, nz, nd1, nd2, rate=rate, niter=a.steps, warmup=max(3, a.warmup), rank=d.rank,
                             world=d.world, sync=d.barrier)
    loop_ms = d.max(r["loop_ms"])
    it = r["iters"]
    N = nx * ny * nz
    fwd, trn, wav = d.max(r["S_fwd_ms"]), d.max(r["S_trans_ms"]), d.max(r["wavelet_slab_ms"])
    ms_it = loop_ms / max(it, 1)
    # config E proper: 25 000 stations per GPU x int(0.0025 N) entries per row on this grid. Its products are
    # NOT scaled up from this thin kernel (a few hundred entries per column segment: the worst case of the layouts);
    # they are taken at the rate the config C section of this repository reaches on a kernel of that size (0.75 of the
    # HBM peak by bytes moved, profiles/r2_bench_8gpu*.jsonl); everything else of the iteration is as measured here.
    peak0, _ = hbm_peak()
    full_nnz = 25000.0 * d.world * int(0.0025 * N)
    non_product = ms_it - (fwd + trn)
    product_full = 2.0 * (6.0 * full_nnz / d.world) / (0.75 * peak0 * 1e9) * 1e3
    projected = non_product + product_full
    peak, _ = hbm_peak()
    return {"workload": "joint grav+mag %dx%dx%d cells, %d + %d data, Haar rate %g, damping + cross-gradient constraint "
                        "(%d rows), wavelet in the LSQR loop (WAVELET_DOMAIN = F)" % (nx, ny, nz, nd1, nd2, rate, r["constraint_rows"]),
            "n_gpus": d.world, "nnz": r["nnz"], "column_slabs": r["column_slabs"],
            "assemble_s": round(d.max(r["assemble_s"]), 2), "constraints_s": round(d.max(r["constraints_s"]), 2),
            "constraint_nnz": int(d.sum(float(r["constraint_nnz_local"]))),
            "lsqr_it_per_s": it / (loop_ms * 1e-3), "ms_per_it": ms_it, "iters": it,
            "residual_last": float(r["history"][-1]) if len(r["history"]) else None,
            "S_fwd_ms": fwd, "S_trans_ms": trn, "wavelet_transform_ms": wav,
            "wavelet_share": 4.0 * wav / ms_it, "wavelet_distributed": r.get("wavelet_distributed"),
            "wavelet_exchange": r.get("wavelet_exchange"),
            "non_product_ms_per_it": non_product,
            "projected_full_config_e": {"nnz": full_nnz, "ms_per_it": projected, "it_per_s": 1e3 / projected,
                                        "note": "measured non-product part of the iteration (wavelets, constraint block, vectors, "
                                                "collectives) + the two products of a 25 000-stations-per-GPU, 0.25 %% kernel at "
                                                "0.75 of the HBM peak by bytes moved (the config C section's measured rate); "
                                                "BASELINE.md: roofline 24 it/s, 60 %% target 14.6 it/s on 8 GPUs"},
            "note": "4 transforms of distributed vectors per iteration: axis-1/2 passes on the planes a rank owns, one all-to-all, "
                    "axis-3 pass on its columns, one all-to-all back (wavelet_distributed = true); falls back to all-gather + "
                    "full transform when a slab is thinner than a plane"}


def compressed_spmv(a, tfx, d):
    """SpMV / SpMV^T GB/s on a wavelet-compressed sensitivity matrix (the second half of BASELINE.json's
    metric): synthetic gravity, same grid and stations as the headline workload, Haar compression at 5 %
    (BASELINE config C's rate), assembled on the device by the reference's row pipeline, kept in the T16
    layouts (6 B/nnz per product); with N ranks the station count is N x comp_ndata (weak scaling). Times come from CUDA events on the library stream around back-to-back
    products (tfx_sparse_matrix_time_product); the matrix (2 x 12.6 GB) is far larger than L2.
    With N > 1 ranks: rows are assembled sharded by data, re-partitioned over NVLink to nnz-balanced column
    slabs (csrc/sensit_dist.cu) and every figure is the whole-job aggregate (total bytes / max time over ranks)."""
    from tomofastx_b200.synth import depth_weight_type1, regular_grid, station_lattice
    # weak scaling: the number of stations grows with the number of GPUs, so every GPU keeps a slab of the same nnz
    # (config C: 2 x 63 GB in the T16 layouts) -- a per-GPU fraction of the HBM peak means something only at that size
    nx, ny, nz = (a.comp_grid if a.comp_grid else (a.nx, a.ny, a.nz))
    rate = a.comp_rate
    N = nx * ny * nz
    per_rank = a.comp_ndata
    # Row blocks. A block is built from triplets: ~28 B per entry at the peak of its build (the CSR, the row ids, the
    # double-buffered (column, position) pairs of the transpose sort), 12 B per entry once its T16 layouts stand. --comp-batch -1 (default) sizes every block to the
    # memory that is free WHEN IT IS BUILT: the first blocks are large (thousands of stations: long segments, the
    # kernels' best case), the last ones small -- instead of equal thin blocks sized for the last one.
    auto_blocks = a.comp_batch < 0
    batch = a.comp_batch if a.comp_batch >= 0 else 0
    free_b, _ = tfx.device_mem_info()
    free_b = d.min(float(free_b))
    nel_row = max(1, int(rate * N))
    kBuild, kReserve = 32.0, 40.0 * N + (4 << 30)
    if auto_blocks:
        need = lambda rows: 12.0 * nel_row * rows + kReserve + kBuild * nel_row * 64
    else:
        blk_rows = (batch // d.world) if batch > 0 else per_rank
        need = lambda rows: 12.0 * nel_row * rows + 36.0 * nel_row * min(blk_rows, rows) + kReserve
    while per_rank > 16 and need(per_rank) > 0.94 * free_b:    # a smaller station count is reported, never silent
        per_rank = int(per_rank * 0.9)
    nd = per_rank * d.world
    a_comp_batch = batch if not auto_blocks else 1

    def next_block(rows_left):
        """Stations (all ranks together) of the next row block."""
        if not auto_blocks:
            return min(a_comp_batch, rows_left)
        fr, _ = tfx.device_mem_info()
        fr = d.min(float(fr))
        fit = int((0.92 * fr - kReserve) / (kBuild * nel_row))     # rows per rank whose build fits now
        fit = min(fit, int(4.0e9 / nel_row))                       # 32-bit positions inside a block's transpose sort
        nb = max(32, fit) * d.world
        if rows_left - nb < 64 * d.world:                          # no crumbs at the end
            nb = rows_left
        return min(nb, rows_left)
    grid = regular_grid(nx, ny, nz)
    xyz = station_lattice(nd, 100.0 * nx, 100.0 * ny, z=-0.1)
    cw = depth_weight_type1(grid, 2.0, 0.0, 4.0e3)
    par = tfx.SensitParams()
    par.problem_type = 1
    par.nx, par.ny, par.nz = nx, ny, nz
    par.ndata, par.ndata_components, par.nmodel_components, par.data_type = nd, 1, 1, 1
    par.compression_type, par.compression_rate = 1, rate
    par.problem_weight = 1.0
    par.cell0, par.ncells_local, par.param_shift, par.ncolumns = 0, N, 0, 2 * N
    tfx.synchronize()
    d.barrier()
    t0 = time.perf_counter()
    t_rows = t_part = None
    grid = tfx.grid_pin(grid)      # one upload of the 48 B/cell grid for the sample and every row block (inside the timing)
    if a_comp_batch > 0:
        # Row-blocked assembly (bounded build memory, csrc/sensit.cu matrix_append_block): the column partition comes
        # from a strided sample of the stations (the reference balances on the nnz counts of ALL rows, which it has on
        # disk before it reads the kernel back; the regular station lattice makes 1/16 of them representative), then
        # the kernel is assembled batch after batch, every batch re-partitioned over NVLink and built into its own
        # row block.
        import copy
        dw1 = np.ones((nd, 1))
        stride = max(1, nd // max(256, nd // 16)) | 1      # odd: never a divisor of an even lattice width (no aliasing)
        sx = tuple(np.ascontiguousarray(v[::stride]) for v in xyz)
        par_s = copy.copy(par); par_s.ndata = sx[0].size
        rows_s, nnz_col, _, _ = tfx.sensit_assemble_rows(par_s, grid, sx, cw, np.ones((sx[0].size, 1)), d.rank, d.world)
        del rows_s
        nnz_at, nel_at = tfx.get_load_balancing_nelements(nnz_col, d.world)
        tfx.synchronize()
        t_sample = d.max(time.perf_counter() - t0)
        ncl, cell0 = int(nel_at[d.rank]), int(nel_at[:d.rank].sum())
        slabs = [int(v) for v in nel_at]
        tfx.set_option("sensit_row_blocks", 1)
        S = tfx.SparseMatrix(nd, 2 * ncl, int(nd) * int(rate * N))
        nnz = 0
        cerr_sum = 0.0
        b0, block_sizes = 0, []
        while b0 < nd:
            nb = next_block(nd - b0)
            block_sizes.append(nb)
            par_b = copy.copy(par); par_b.ndata = nb
            xb = tuple(np.ascontiguousarray(v[b0:b0 + nb]) for v in xyz)
            rows_b, _, cerr_b, tot_b = tfx.sensit_assemble_rows(par_b, grid, xb, cw, dw1[b0:b0 + nb], d.rank, d.world)
            tfx.sensit_repartition_into(S, rows_b, 1, nel_at, d.rank, d.world)
            del rows_b
            nnz += int(tot_b); cerr_sum += cerr_b * nb
            b0 += nb
        S.finalize()
        tfx.set_option("sensit_row_blocks", 0)
        cerr = cerr_sum / nd
        nnz_loc = S.get_number_elements()
        t_rows = None
    elif d.world == 1:
        S, _, cerr, nnz = tfx.calculate_sensit(par, grid, xyz, cw, np.ones((nd, 1)))
        ncl, cell0, nnz_loc = N, 0, nnz
        slabs = [N]
    else:
        rows, nnz_col, cerr, nnz = tfx.sensit_assemble_rows(par, grid, xyz, cw, np.ones((nd, 1)), d.rank, d.world)
        tfx.synchronize()
        t_rows = d.max(time.perf_counter() - t0)
        nnz_at, nel_at = tfx.get_load_balancing_nelements(nnz_col, d.world)
        S = tfx.sensit_repartition(rows, 1, nel_at, d.rank, d.world)
        ncl, cell0, nnz_loc = int(nel_at[d.rank]), int(nel_at[:d.rank].sum()), int(nnz_at[d.rank])
        slabs = [int(v) for v in nel_at]
        del rows
    tfx.grid_unpin()
    tfx.synchronize()
    t_asm = d.max(time.perf_counter() - t0)
    if t_rows is not None:
        t_part = t_asm - t_rows
    ncol = 2 * ncl
    assert S.storage_kind() == 2, "compressed matrix must be in the T16 layouts"
    peak, peak_src = hbm_peak()
    rng = np.random.default_rng(1235)
    x = tfx.Buffer(ncol); u = tfx.Buffer(nd); q = tfx.Buffer(nd); t = tfx.Buffer(ncol)
    x_full = rng.uniform(-1.0, 1.0, N)
    x_loc = np.zeros(ncol); x_loc[:ncl] = x_full[cell0:cell0 + ncl]
    tfx.copy(x, x_loc, ncol)
    tfx.copy(u, rng.uniform(-1.0, 1.0, nd), nd)
    out = {"workload": "synthetic gravity %dx%dx%d cells, %d data, Haar wavelet compression %g" % (nx, ny, nz, nd, rate),
           "nnz": int(nnz), "compression_error": cerr, "assemble_s": round(t_asm, 2),
           "layout": "T16 (f32 value + u16 in-tile key = 6 B/nnz per product, one copy per direction)",
           "peak": peak * d.world, "peak_source": peak_src + (" x %d GPUs" % d.world if d.world > 1 else ""),
           "unit": "GB/s", "reps": a.comp_reps, "scaling": "weak (%d stations per GPU)" % per_rank}
    if per_rank != a.comp_ndata:
        out["note"] = "stations per GPU reduced from %d to %d to fit the free HBM (%.0f GB)" % (a.comp_ndata, per_rank, free_b / 1e9)
    if a_comp_batch > 0:
        out["row_blocks"] = {"stations_per_block": block_sizes, "blocks": len(block_sizes),
                             "sizing": "each block sized to the HBM free when it is built" if auto_blocks else "fixed",
                             "partition_sample_s": round(t_sample, 2), "device_bytes_per_rank_max": int(d.max(float(S.device_bytes())))}
    if d.world > 1:
        out["column_slabs"] = slabs
        out["nnz_per_rank_max_over_mean"] = d.max(float(nnz_loc)) / (float(nnz) / d.world)
        if t_rows is not None:
            out["assemble_rows_s"], out["repartition_s"] = round(t_rows, 2), round(t_part, 2)
    l0 = tfx.launch_count()
    for name, tr, xi, yo in (("forward", 0, x, q), ("transposed", 1, u, t)):
        d.barrier()
        ms = d.max(S.time_product(tr, xi, yo, a.comp_reps))
        # algorithmic bytes of SURVEY 8(d): 8 B/nnz (f32 value + int32 column) + the vectors; bytes moved: 6 B/nnz
        vec = 8.0 * (N + 2 * nd * d.world) if tr == 0 else 8.0 * (nd * d.world + 2 * N)
        alg = 8.0 * nnz + vec
        moved = 6.0 * nnz + vec
        out[name] = {"ms": ms, "achieved": alg / ms / 1e6, "frac": alg / ms / 1e6 / (peak * d.world),
                     "moved_gbs": moved / ms / 1e6, "moved_frac": moved / ms / 1e6 / (peak * d.world)}
    out["gpu_launches"] = int(tfx.launch_count() - l0)
    # size-independent parity property at full size: <S x, u> == <x, S^T u> (the two products use two
    # different copies of the matrix, so this also checks the layouts against each other)
    xh, uh = x.numpy(), u.numpy()
    lhs, rhs = d.sum(float(np.dot(q.numpy(), uh))), d.sum(float(np.dot(xh, t.numpy())))
    out["adjoint_rel_err"] = abs(lhs - rhs) / max(abs(lhs), abs(rhs), 1e-300)
    assert out["adjoint_rel_err"] < 1e-10, out["adjoint_rel_err"]
    if d.world == 1:
        # 3-D wavelet transform on a device-resident volume (SURVEY 8d metric iii: 16 B per element and transform)
        vol = tfx.Buffer(N)
        tfx.copy(vol, rng.uniform(-1.0, 1.0, N), N)
        out["wavelet"] = {}
        for wname, wtype in (("haar", 1), ("daubechies_d4", 2)):
            for _ in range(2):
                tfx.forward_wavelet(vol, nx, ny, nz, wtype); tfx.inverse_wavelet(vol, nx, ny, nz, wtype)
            tfx.timer_start()
            for _ in range(10):
                tfx.forward_wavelet(vol, nx, ny, nz, wtype); tfx.inverse_wavelet(vol, nx, ny, nz, wtype)
            ms = tfx.timer_stop() / 20.0
            # bytes moved per element and transform: D4 three axis passes (48 B); Haar axis 1 fused with the three lowest
            # axis-2 scales + the high axis-2 scales on every 8th row + axis 3: 34 B forward, 36 B inverse
            moved = 48.0 if wtype == 2 else 35.0
            out["wavelet"][wname] = {"ms_per_transform": ms, "achieved": 16.0 * N / ms / 1e6, "frac": 16.0 * N / ms / 1e6 / peak,
                                     "moved_frac": moved * N / ms / 1e6 / peak,
                                     "note": "algorithmic 16 B/element; bytes moved: %g B/element" % moved}
        del vol
    out["assembly"] = {"rows_per_s": nd / t_asm, "cell_evaluations_per_s": float(nd) * N / t_asm,
                       "note": "kernel line + column weight + Haar transform + exact k-th threshold + compaction per row"
                               + ("; rows sharded over %d GPUs + all-to-all re-partitioning" % d.world if d.world > 1 else "")}
    # LSQR on the compressed matrix (wavelet-domain solve, split path: one product per direction per iteration)
    m = np.zeros((nz, ny, nx))
    sl = lambda n: slice(max(0, n // 2 - max(1, n // 8)), n // 2 + max(1, n // 8))
    m[sl(nz), sl(ny), sl(nx)] = 250.0
    xs = np.zeros(ncol)
    xs[:ncl] = tfx.forward_wavelet((m.ravel() / cw).copy(), nx, ny, nz, 1)[cell0:cell0 + ncl]
    d_obs = S.mult_vector(xs)
    if d.world > 1:
        tfx.comm_allreduce_sum(d_obs, nd)
    Cm = tfx.SparseMatrix.from_arrays(N, ncol, np.full(ncl, 1.0e-11, dtype=np.float32), np.arange(1, ncl + 1, dtype=np.int32),
                                      np.arange(1, ncl + 2, dtype=np.int64), np.arange(cell0 + 1, cell0 + ncl + 1, dtype=np.int32))
    nlines = nd + N
    b = np.zeros(nlines); b[:nd] = d_obs
    ub, xb = tfx.Buffer(nlines), tfx.Buffer(ncol)
    for niter in (3, a.steps):
        tfx.copy(ub, b, nlines)
        d.barrier()
        tfx.lsqr_solve_sensit(nlines, ncol, niter, 1.0e-13, 0.0, 0.0, S, Cm, ub, xb, [1, 0], ncl, nx, ny, nz, 1, 1, True,
                              myrank=d.rank, nbproc=d.world)
    loop_ms, _, _ = tfx.last_timing()
    loop_ms = d.max(loop_ms)
    hist, iters, fused = tfx.last_history()
    out["lsqr"] = {"it_per_s": iters / (loop_ms * 1e-3), "ms_per_it": loop_ms / max(iters, 1), "iters": int(iters),
                   "residual_last": float(hist[-1]) if len(hist) else None,
                   "ref_accounting_gbs": 16.0 * nnz * iters / loop_ms / 1e6}
    return out


# ----------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the oracle port on the host cores
# ----------------------------------------------------------------------------------------------------
def _cpu_worker(args):
    """One 'MPI rank' of the reference's column split: LSQR iterations on its own column slab."""
    seed, nrows, ncols, iters = args
    sys.path.insert(0, ROOT)
    from oracle import oracle as orc
    rng = np.random.default_rng(seed)
    # the solver the inversion calls (lsqr_solve_sensit, lsqr_solver2.F90:47-308) on this rank's column slab: the
    # sensitivity block of problem 1 in a matrix of 2 * nelements columns (joint_inverse_problem.F90:213-214), the
    # damping block alpha * I below it (damping.F90:97-261), right-hand side [data residuals, 0]
    m = orc.SparseMatrix(nrows, 2 * ncols, nrows * ncols)
    cols = np.arange(1, ncols + 1, dtype=np.int32)
    for _ in range(nrows):
        m.add_row(rng.standard_normal(ncols, dtype=np.float32), cols)
        m.new_row()
    m.finalize()
    cm = orc.SparseMatrix(ncols, 2 * ncols, ncols)
    one = np.array([1.0e-11], dtype=np.float32)
    for i in range(ncols):
        cm.add_row(one, cols[i:i + 1])
        cm.new_row()
    cm.finalize()
    b = np.concatenate([rng.standard_normal(nrows), np.zeros(ncols)])
    solve = lambda n: orc.lsqr_solve_sensit(n, 1e-300, 0.0, 0.0, m, cm, b, ncols, 1, 1, 1, 1, 0, True, (1, 0))
    solve(1)                                                           # warm the caches / page in
    t0 = time.perf_counter()
    x, hist, it = solve(iters)
    dt = time.perf_counter() - t0
    # the solver does the initial S^T u (half an iteration's matrix traffic) plus `it` iterations
    return dt, it + 0.5, float(nrows) * ncols


def cpu_reference(a, steps, warmup, seconds_hint=20.0):
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    nrows = a.ndata
    # bounded sample: each worker owns a slab of columns sized for ~seconds_hint of work in total
    # (~2.5 ns per matrix entry and product on one core), capped by host memory (8 B per entry)
    per_iter_entries = seconds_hint / max(steps + warmup + 1.5, 1.0) / 2.5e-9 / 2.0
    ncols = int(max(256, min(per_iter_entries / nrows, 3.0e9 / (8.0 * nrows))))
    ctx = mp.get_context("spawn")
    with ctx.Pool(cores) as pool:
        # P = 1 (SURVEY 8d asks for both): one process alone on the box, a quarter of a slab
        dt1, its1, ent1 = pool.map(_cpu_worker, [(999, nrows, max(256, ncols // 4), steps + warmup)])[0]
        res = pool.map(_cpu_worker, [(1000 + i, nrows, ncols, steps + warmup) for i in range(cores)])
    # ranks run concurrently; an iteration of the whole slab set ends when the slowest rank ends
    t_iter = max(dt / its for dt, its, _ in res)
    sample_entries = sum(e for _, _, e in res)
    full_entries = float(a.ndata) * a.nx * a.ny * a.nz
    its_per_s_sample = 1.0 / t_iter
    value = its_per_s_sample * sample_entries / full_entries          # SpMV cost is linear in nnz
    return {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "projected": True,
            "projection": "measured on a column sample that fits the host RAM (the full CSR is 335 GB), scaled linearly in "
                          "nnz; workers are independent column slabs WITHOUT the per-iteration all-reduce of u (favours the "
                          "CPU arm)",
            "value_1core": (its1 / dt1) * ent1 / full_entries,
            "sample": "oracle lsqr_solve_sensit (C port of lsqr_solver2.F90:47-308 + sparse_matrix.f90:313-405) with the "
                      "damping block, %d column-slab processes x (%d rows x %d dense columns, CSR f32+i32), %d iterations "
                      "each; scaled linearly in nnz from %.3g to %.3g entries" %
                      (cores, nrows, ncols, steps + warmup, sample_entries, full_entries)}


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cb = cpu_reference(a, steps=a.steps, warmup=a.warmup, seconds_hint=60.0)
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": a.gpus,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 / cb["value"], "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64 (f32 matrix values, f64 vectors/accumulation)",
            "data": "synthetic", "config": {"workload": workload_name(a), "parallelism": "%d host processes "
                                            "(column split, lsqr_solver2.F90:16)" % cb["cores"]},
            "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            "note": "reference = Fortran 2008 + MPI, not buildable in this image nor on the GPU box (no Fortran compiler, no "
                    "MPI: profiles/r2_probe_fortran_gpubox.txt): the timed arm is the line-faithful C port (oracle/) and its "
                    "value is PROJECTED from a column sample (cpu_baseline.projection)"}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--nx", type=int, default=256)
    ap.add_argument("--ny", type=int, default=256)
    ap.add_argument("--nz", type=int, default=64)
    ap.add_argument("--ndata", type=int, default=10000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--dense-vec4", type=int, default=None, help="A/B switch of the dense sweep kernel shape (option dense_vec4)")
    ap.add_argument("--dense-f2f-rows", type=int, default=None, help="A/B switch (option dense_f2f_rows)")
    ap.add_argument("--no-compressed", action="store_true", help="skip the compressed SpMV section")
    ap.add_argument("--comp-ndata", type=int, default=6250, help="stations per GPU of the compressed section")
    ap.add_argument("--comp-rate", type=float, default=0.05)
    ap.add_argument("--comp-reps", type=int, default=20)
    ap.add_argument("--comp-grid", type=int, nargs=3, default=(512, 512, 128), metavar=("NX", "NY", "NZ"),
                    help="grid of the compressed section (default: BASELINE config C, 512 512 128)")
    ap.add_argument("--comp-batch", type=int, default=-1,
                    help="assemble the compressed kernel in row blocks of this many stations (all ranks together); "
                         "0 = one piece, -1 (default) = 400 per rank")
    ap.add_argument("--no-config-d", action="store_true", help="skip the config D extra")
    ap.add_argument("--opt", action="append", default=[], metavar="NAME=VALUE", help="tfx_set_option before the run (A/B switches)")
    ap.add_argument("--config-e", type=int, default=-1, help="1 / 0: run / skip the config E extra (default: run when --gpus >= 2)")
    ap.add_argument("--e-grid", type=int, nargs=3, default=None, metavar=("NX", "NY", "NZ"))
    ap.add_argument("--e-ndata", type=int, default=0, help="stations per problem and GPU of the config E extra")
    ap.add_argument("--no-dense", action="store_true", help="run only the compressed section")
    a = ap.parse_args()
    a.warmup = max(a.warmup, 3) if a.impl == "ours" else a.warmup
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
