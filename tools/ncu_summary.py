"""Condense an .ncu-rep (ncu --set full) into the handful of numbers DESIGN.md / bench.py quote.
usage: python tools/ncu_summary.py file.ncu-rep [more.ncu-rep ...]  > profiles/rNN_ncu_summary.txt"""
import csv
import io
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__shared_mem_per_block_dynamic",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
]

for path in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    if len(rows) < 3:
        print(f"## {path}: no kernels")
        continue
    h, units = rows[0], rows[1]
    for r in rows[2:]:
        print(f"## {path} :: {r[h.index('Kernel Name')][:90]}")
        for w in WANT:
            if w in h:
                print(f"   {w:82s} {r[h.index(w)]:>16s} {units[h.index(w)]}")
        stalls = []
        for i, name in enumerate(h):
            if "pcsamp_warps_issue_stalled" in name and not name.endswith("not_issued"):
                try:
                    stalls.append((float(r[i]), name.split("issue_stalled_")[1]))
                except ValueError:
                    pass
        tot = sum(v for v, _ in stalls) or 1.0
        print("   stall samples: " + ", ".join(f"{n} {100 * v / tot:.1f}%" for v, n in sorted(stalls, reverse=True)[:8]))
        print()
