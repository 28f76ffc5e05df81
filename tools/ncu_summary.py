#!/usr/bin/env python3
"""Key metrics per kernel launch of an ncu report.  usage: tools/ncu_summary.py <report.ncu-rep>"""
import csv, io, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
want = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("smsp__inst_executed.sum", "warp_inst"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"), ("launch__registers_per_thread", "regs"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "bank_conf"), ("smsp__thread_inst_executed_per_inst_executed.ratio", "thr/inst"),
        ("sm__inst_executed_pipe_alu.sum", "alu"), ("sm__inst_executed_pipe_fma.sum", "fma"), ("sm__inst_executed_pipe_lsu.sum", "lsu")]
for r in data:
    name = r[idx["Kernel Name"]].split("(")[0]
    print(name, " ".join("%s=%s%s" % (n, r[idx[k]], units[idx[k]].replace("byte", "B") if "dram_" in n or n == "time" else "") for k, n in want if k in idx))
