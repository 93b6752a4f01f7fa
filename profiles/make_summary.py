#!/usr/bin/env python
"""Turns an .ncu-rep (ncu --set full --import-source on) into the small CSV summaries kept under profiles/.

    python profiles/make_summary.py gpurun_out/x.ncu-rep "comment line" [launch index] > profiles/r1_x_ncu_summary.csv
"""
import csv
import io
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__bytes_read.sum.per_second", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__cycles_elapsed.avg.per_second", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct",
]


def ncu_csv(rep, page, launch=None):
    sel = ["--launch-skip", str(launch), "--launch-count", "1"] if (launch is not None and page == "source") else []
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"] + sel, capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep, comment = sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else ""
    launch = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    rows = ncu_csv(rep, "raw")
    hdr, units, vals = rows[0], rows[1], rows[2 + launch]
    col = {h: i for i, h in enumerate(hdr)}
    print("# " + comment)
    print("Kernel Name,%s," % vals[col["Kernel Name"]])
    for m in METRICS:
        if m in col:
            print("%s,%s,%s" % (m, vals[col[m]], units[col[m]]))
    stalls = []
    for h, i in col.items():
        if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
            try:
                stalls.append((float(vals[i]), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
            except ValueError:
                pass
    tot = sum(v for v, _ in stalls) or 1.0
    for v, n in sorted(stalls, reverse=True)[:10]:
        print("stall_%s,%.2f,%%" % (n, 100.0 * v / tot))
    src = ncu_csv(rep, "source", launch if len(rows) > 3 else None)
    h = next(i for i, r in enumerate(src) if r and r[0] == "Address")
    idx = {n: j for j, n in enumerate(src[h])}
    ops = {}
    total = 0
    for r in src[h + 1:]:
        if len(r) <= idx["Instructions Executed"]:
            continue
        try:
            n = int(r[idx["Instructions Executed"]])
        except ValueError:
            continue
        toks = r[idx["Source"]].split()
        op = next((t for t in toks if not t.startswith("@")), "?").split(".")[0]
        ops[op] = ops.get(op, 0) + n
        total += n
    print("\n# executed warp instructions by opcode (source page), total %d" % total)
    for op, n in sorted(ops.items(), key=lambda kv: -kv[1])[:16]:
        print("op_%s,%.2f,%%" % (op, 100.0 * n / max(total, 1)))


if __name__ == "__main__":
    main()
