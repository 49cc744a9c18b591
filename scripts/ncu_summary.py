"""Summarise ncu reports (`--page raw --csv`) into profiles/: raw CSV per report + one JSON with the metrics the docs quote.
Usage: ncu_summary.py TAG name1 name2 ...   (reads gpurun_out/TAG_name.ncu-rep)"""
import csv
import io
import json
import subprocess
import sys

KEYS = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__waves_per_multiprocessor", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "l1tex__t_sector_hit_rate.pct"]

tag, names = sys.argv[1], sys.argv[2:]
out = {}
for n in names:
    rep = f"gpurun_out/{tag}_{n}.ncu-rep"
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    open(f"profiles/{tag}_{n}_ncu_full_raw.csv", "w").write(raw)
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = {}
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                d[k] = r[i] + ((" " + units[i]) if units[i] else "")
        res.append(d)
    out[f"{tag}_{n}"] = res
json.dump(out, open(f"profiles/{tag}_ncu_summary.json", "w"), indent=1)
for k, v in out.items():
    for d in v:
        print(k, {x: d.get(x) for x in ("Kernel Name", "gpu__time_duration.sum", "smsp__inst_executed.sum", "launch__registers_per_thread",
                                        "smsp__issue_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum")})
