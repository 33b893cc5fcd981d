"""Selected raw metrics of an .ncu-rep as JSON (what profiles/ncu_*_summary.json hold).
    python scripts/ncu_summary.py gpurun_out/prof_r1_scan.ncu-rep > profiles/ncu_r1_scan_summary.json"""
import csv
import io
import json
import subprocess
import sys

KEYS = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "lts__t_sector_hit_rate.pct",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
res = []
for r in rows[2:]:
    d = {}
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            d[k] = (r[i] + " " + units[i]).strip()
    if len(sys.argv) > 2 and not any(s in d.get("Kernel Name", "") for s in sys.argv[2:]):
        continue
    res.append(d)
json.dump(res, sys.stdout, indent=1)
