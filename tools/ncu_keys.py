#!/usr/bin/env python3
"""Print the key metrics of every kernel in an `ncu --page raw --csv` export (dev tool)."""
import csv
import sys

KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__icc_request_hit_rate.pct",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__cycles_elapsed.avg.per_second",
        "lts__t_bytes.sum", "l1tex__t_bytes_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_bytes_pipe_lsu_mem_local_op_st.sum"]
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("==", d.get("Kernel Name", "")[:100])
    for k in KEYS:
        if k in d:
            print(f"   {k:90s} {d[k]} {units[hdr.index(k)]}")
    for k in hdr:
        if "issue_stalled" in k and k.endswith("per_issue_active.ratio") and "not_issued" not in k:
            try:
                if float(d[k]) > 0.2:
                    print(f"   {k:90s} {d[k]}")
            except ValueError:
                pass
