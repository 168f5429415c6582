#!/usr/bin/env python
"""Pick the roofline-relevant counters out of `ncu -i X.ncu-rep --page raw --csv` dumps.
Usage: python tools/summarize_ncu.py name=raw.csv [name=raw.csv ...] > ncu_full_summary.json"""
import csv
import json
import sys

KEEP = ("Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__bytes_read.sum.per_second", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__t_bytes.sum", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_off.avg.per_cycle_elapsed",
        "sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_off.avg.peak_sustained",
        "smsp__inst_executed_op_tma_ld.sum", "smsp__sass_inst_executed_op_utcmma.sum",
        "smsp__sass_inst_executed_op_tmem_ldt.sum", "smsp__inst_executed_op_ldgsts.sum", "gpc__cycles_elapsed.max",
        "sm__cycles_active.avg")
out = {}
for arg in sys.argv[1:]:
    name, path = arg.split("=", 1)
    rows = list(csv.reader(open(path)))
    if len(rows) < 3:
        continue
    hdr, units = rows[0], rows[1]
    recs = []
    for r in rows[2:]:
        d = {}
        for h, u, v in zip(hdr, units, r):
            if h in KEEP and v != "":
                d[h + (f" [{u}]" if u else "")] = v
        recs.append(d)
    out[name] = recs
json.dump(out, sys.stdout, indent=1)
