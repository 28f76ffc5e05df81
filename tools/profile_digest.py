#!/usr/bin/env python3
"""Digest one GPU round's ncu outputs into tracked files under profiles/.
usage: tools/profile_digest.py <tag> <frames-per-launch>
  gpurun_out/launches_<tag>.csv  -> profiles/<tag>_launches.csv (copy) + profiles/<tag>_launch_summary.csv
  gpurun_out/prof_<tag>.ncu-rep  -> profiles/<tag>_full_raw.csv (selected metrics per launch) + profiles/traffic.json (DRAM bytes per frame per kernel)"""
import csv, io, json, os, shutil, subprocess, sys, collections

tag, frames = sys.argv[1], int(sys.argv[2])
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")

def short(name):
    n = name.split("(")[0].replace("void ", "")
    return n.split("<")[0]

lc = os.path.join(G, "launches_%s.csv" % tag)
if os.path.exists(lc):
    shutil.copy(lc, os.path.join(P, "%s_launches.csv" % tag))
    lines = [l for l in open(lc) if not l.startswith("==")]
    rows = list(csv.DictReader(io.StringIO("".join(lines))))
    agg = collections.OrderedDict()
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", "")); u = r["Metric Unit"]
        us = v / 1e3 if u in ("ns", "nsecond") else v * 1e3 if u in ("ms", "msecond") else v
        a = agg.setdefault(short(r["Kernel Name"]), [0, 0.0]); a[0] += 1; a[1] += us
    tot = sum(a[1] for a in agg.values()) or 1.0
    with open(os.path.join(P, "%s_launch_summary.csv" % tag), "w") as f:
        f.write("# %s: ncu --metrics gpu__time_duration.sum --clock-control none, bench.py --steps 1 --warmup 1 --batch %d --no-cpu-baseline --no-e2e\n" % (tag, frames))
        f.write("# per-launch times are cold-cache and serialised: compare SHARES with bench.py's stage_share, not absolutes\n")
        f.write("kernel,launches,total_us,avg_us,share_of_own_kernels\n")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("%s,%d,%.1f,%.1f,%.4f\n" % (k, a[0], a[1], a[1] / a[0], a[1] / tot))
rep = os.path.join(G, "prof_%s.ncu-rep" % tag)
if os.path.exists(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__thread_inst_executed_per_inst_executed.ratio"]
    want = [w for w in want if w in idx]
    traffic = {}; issue = {}; dram_pct = {}
    def to_bytes(v, u):
        v = float(v.replace(",", "")); m = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        return v * m.get(u, 1)
    with open(os.path.join(P, "%s_full_raw.csv" % tag), "w") as f:
        f.write("# %s: ncu --set full --clock-control none --import-source on, %d frames per launch\n" % (tag, frames))
        f.write(",".join(want) + "\n"); f.write(",".join(units[idx[w]] for w in want) + "\n")
        for r in data:
            f.write(",".join('"%s"' % short(r[idx[w]]) if w == "Kernel Name" else r[idx[w]].replace(",", "") for w in want) + "\n")
            k = short(r[idx["Kernel Name"]])
            b = to_bytes(r[idx["dram__bytes_read.sum"]], units[idx["dram__bytes_read.sum"]]) + to_bytes(r[idx["dram__bytes_write.sum"]], units[idx["dram__bytes_write.sum"]])
            traffic.setdefault(k, []).append(b / frames)
            try:
                issue.setdefault(k, []).append(float(r[idx["smsp__issue_active.avg.pct_of_peak_sustained_active"]]))
                dram_pct.setdefault(k, []).append(float(r[idx["gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]]))
            except Exception:
                pass
    tj = {k: sum(v) / len(v) for k, v in traffic.items()}
    # k_pyr_resize launches once per level: report the per-frame SUM over the 7 levels
    for k in list(tj):
        if "resize" in k:
            tj[k] = sum(traffic[k]) / max(1, len(traffic[k]) // 7)
    json.dump({"source": "%s_full_raw.csv" % tag, "frames_per_launch": frames, "dram_bytes_per_frame": tj,
               "issue_active_pct": {k: sum(v) / len(v) for k, v in issue.items()}, "dram_throughput_pct": {k: sum(v) / len(v) for k, v in dram_pct.items()}},
              open(os.path.join(P, "traffic.json"), "w"), indent=1)
    print(json.dumps(tj, indent=1))
