#!/usr/bin/env python3
"""Per-source-line summary of one kernel in an ncu report (needs -lineinfo and --import-source on).
usage: tools/ncu_lines.py <report.ncu-rep> <kernel-regex> [top-N]"""
import csv, io, subprocess, sys

rep, kre = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", "regex:" + kre, "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None; cur_file = None; agg = {}
for r in rows:
    if len(r) >= 2 and r[0] == "File Name":
        cur_file = r[1].split("/")[-1]; continue
    if len(r) > 10 and r[0] == "Line No":
        hdr = r; continue
    if hdr is None or len(r) != len(hdr) or r[0] == "":
        continue
    d = dict(zip(hdr, r))
    try:
        inst = int(d["Instructions Executed"]); smp = int(d["# Samples"]); thr = d["Avg. Threads Executed"]
    except ValueError:
        continue
    key = (cur_file, int(r[0]))
    a = agg.setdefault(key, [0, 0, r[1], thr])
    a[0] += inst; a[1] += smp
tot_i = sum(a[0] for a in agg.values()) or 1; tot_s = sum(a[1] for a in agg.values()) or 1
print("total warp-instructions %d, samples %d" % (tot_i, tot_s))
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%5.1f%% inst %5.1f%% smp  thr %-4s %s:%d  %s" % (100.0 * a[0] / tot_i, 100.0 * a[1] / tot_s, a[3], f, ln, a[2].strip()[:110]))
