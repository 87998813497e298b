"""Summarise an .ncu-rep (from `ncu --set full ...`) as the per-kernel metric table committed under profiles/.

    ncu -i gpurun_out/prof.ncu-rep --page raw --csv > raw.csv ; python tools/ncu_summary.py raw.csv
"""
import csv
import sys

METRICS = [
    "gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.pct_of_peak_sustained_elapsed",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_sector_hit_rate.pct", "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__t_sector_hit_rate.pct",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.max",
]
rows = list(csv.reader(open(sys.argv[1])))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
h, units = rows[hdr], rows[hdr + 1]
col = {c: i for i, c in enumerate(h)}
for r in rows[hdr + 2:]:
    if len(r) < len(h):
        continue
    print("== %s   grid %s block %s" % (r[col["Kernel Name"]][:110], r[col["Grid Size"]].replace(" ", ""), r[col["Block Size"]].replace(" ", "")))
    for m in METRICS:
        if m in col:
            print("%-78s %-16s %s" % (m, units[col[m]], r[col[m]]))
    print()
