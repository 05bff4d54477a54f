#!/usr/bin/env python
"""Condense `ncu -i X.ncu-rep --page raw --csv` output into the short per-kernel text summaries kept under profiles/.

    python tools/ncu_summary.py gpurun_out/r01c_reg_pinhole_k11_k12.raw.csv "header line" > profiles/r01c_....txt
"""
import csv
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__cycles_active.avg", "sm__cycles_elapsed.avg"]


def main():
    path = sys.argv[1]
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    if len(sys.argv) > 2:
        print("# " + sys.argv[2])
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print("--- %s   [grid %s x %s]" % (d.get("Kernel Name", "?")[:110], d.get("launch__grid_size", "?"), d.get("launch__block_size", "?")))
        for k in KEYS:
            if k in d and d[k] != "":
                print("%-78s %s %s" % (k, d[k], units[hdr.index(k)]))


if __name__ == "__main__":
    main()
