#!/usr/bin/env python
"""Turn an `ncu --set full` report (gpurun_out/*.ncu-rep) into the small text summary committed under
profiles/.  Usage: python profiles/summarize_ncu.py gpurun_out/k1.ncu-rep > profiles/r1_k1.md"""
import csv
import io
import subprocess
import sys
from collections import defaultdict

KEYS = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__cycles_elapsed.avg",
    "sm__cycles_elapsed.avg.per_second", "smsp__cycles_active.avg", "smsp__inst_executed.sum",
    "sm__inst_executed.avg.per_cycle_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "smsp__warps_active.avg.per_cycle_active",
    "smsp__warps_eligible.avg.per_cycle_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__bytes_write.sum.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_bytes.sum",
]


def ncu_csv(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main(rep):
    raw = ncu_csv(rep, "raw")
    hdr, units, vals = raw[0], raw[1], raw[2]
    d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
    print(f"# ncu summary of `{rep}`\n")
    print(f"kernel: `{d.get('Kernel Name', ('?',))[0]}`\n")
    print("| metric | value | unit |\n|---|---:|---|")
    for k in KEYS:
        if k in d:
            print(f"| {k} | {d[k][0]} | {d[k][1]} |")
    print("\nwarp stall reasons (average warps stalled per issue-active cycle):\n")
    print("| reason | ratio |\n|---|---:|")
    for h in hdr:
        if "average_warps_issue_stalled" in h and "per_issue_active" in h and "not_issued" not in h:
            v = float(d[h][0] or 0)
            if v >= 0.01:
                print(f"| {h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')} | {v:.3f} |")
    src = ncu_csv(rep, "source")
    if len(src) > 2:
        h = src[1]
        ix = {n: i for i, n in enumerate(h)}
        agg = defaultdict(lambda: [0, 0])
        for r in src[2:]:
            try:
                ex, n = int(r[ix["Instructions Executed"]]), int(r[ix["# Samples"]])
            except (ValueError, IndexError, KeyError):
                continue
            toks = r[ix["Source"]].split()
            if not toks:
                continue
            op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
            key = "IMAD.WIDE" if op.startswith("IMAD.WIDE") else op.split(".")[0]
            agg[key][0] += ex
            agg[key][1] += n
        tot = sum(a[0] for a in agg.values())
        print("\nexecuted warp-instructions by opcode (SASS, whole kernel):\n")
        print("| opcode | warp-inst (M) | share | stall samples |\n|---|---:|---:|---:|")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:18]:
            print(f"| {k} | {a[0] / 1e6:.1f} | {100 * a[0] / tot:.1f}% | {a[1]} |")
        print(f"| total | {tot / 1e6:.1f} | | |")


if __name__ == "__main__":
    main(sys.argv[1])
