#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --set full, one kernel) into the markdown kept under profiles/.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep "title" > profiles/rNN_<kernel>_ncu_summary.md
"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.avg", "smsp__cycles_active.avg",
    "sm__sass_thread_inst_executed_op_ffma_pred_on.sum", "smsp__inst_executed.sum",
]


def main():
    rep, title = sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    names, units, vals = rows[0], rows[1], rows[2]
    d = {n: (u, v) for n, u, v in zip(names, units, vals)}
    print(f"# {title}\n")
    print(f"Source: `{rep}` (ncu --set full --clock-control none --import-source on, 1 launch).\n")
    print(f"Kernel: `{d.get('Kernel Name', ('', '?'))[1]}`\n")
    print("| metric | value | unit |\n|---|---|---|")
    for k in KEYS:
        if k in d:
            print(f"| {k} | {d[k][1]} | {d[k][0]} |")
    print("\n## Warp stall reasons (per issued instruction)\n\n| reason | ratio |\n|---|---|")
    stalls = [(n, float(v[1])) for n, v in d.items()
              if n.startswith("smsp__average_warps_issue_stalled_") and n.endswith("_per_issue_active.ratio") and "not_issued" not in n]
    for n, v in sorted(stalls, key=lambda t: -t[1])[:10]:
        print(f"| {n[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]} | {v:.3f} |")
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    h = rows[hdr]
    ix = {n: i for i, n in enumerate(h)}
    data = [r for r in rows[hdr + 1:] if len(r) >= len(h)]
    tot = sum(int(r[ix["# Samples"]] or 0) for r in data)
    print(f"\n## Hottest SASS instructions (of {tot} samples)\n\n| samples | % | instruction |\n|---|---|---|")
    for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]] or 0))[:12]:
        s = int(r[ix["# Samples"]] or 0)
        print(f"| {s} | {100.0 * s / max(tot, 1):.1f} | `{r[ix['Source']].strip()[:90]}` |")
    ops = {}
    for r in data:
        op = r[ix["Source"]].strip().split()[0] if r[ix["Source"]].strip() else "?"
        if op.startswith("@"):
            op = r[ix["Source"]].strip().split()[1]
        ops[op.split(".")[0]] = ops.get(op.split(".")[0], 0) + int(r[ix["Instructions Executed"]] or 0)
    print("\n## Executed warp instructions by opcode\n\n| opcode | warp instructions |\n|---|---|")
    for op, c in sorted(ops.items(), key=lambda t: -t[1])[:10]:
        print(f"| {op} | {c} |")


if __name__ == "__main__":
    main()
