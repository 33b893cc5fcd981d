"""DRAM bytes per launch of the fused scan kernel's Q1 and Q2 instances, from an `ncu --set full` report, stamped with the
SASS hash of the library the report was taken from (bench.py reports the number as `roofline.traffic` only while the built
library still hashes to the stamp).

    python scripts/make_traffic.py gpurun_out/prof_r2_scan.ncu-rep        -> profiles/q1_scan_traffic.json, q2_scan_traffic.json
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scripts"))
from sass_hash import kernel_sass_hash  # noqa: E402

LIB = os.path.join(ROOT, "bo-sql_b200", "libbosql_b200.so")
WHICH = {   # file, demangled-name fragment in the report, regex on the mangled name in the library, algorithmic bytes per row
    "q1": ("q1_scan_traffic.json", "k_scan<12324, 65, 40, 1, 1>", r"k_scanILj12324ELj65ELj40ELi1ELb1E", 16),
    "q2": ("q2_scan_traffic.json", "k_scan<16777745, 0, 105, 2, 0>", r"k_scanILj16777745ELj0ELj105ELi2ELb0E", 32),
}


def gb(text, unit):
    v = float(text.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[unit]


def main():
    rep = sys.argv[1]
    rows_per_launch = int(float(sys.argv[2])) if len(sys.argv) > 2 else 1_000_000_000
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    ir, iw, it, ik = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("gpu__time_duration.sum"), hdr.index("Kernel Name")
    for key, (fname, frag, regex, bpr) in WHICH.items():
        hits = [r for r in rows[2:] if frag.replace(" ", "") in r[ik].replace(" ", "")]
        if not hits:
            print(f"{key}: no launch of {frag} in {rep}")
            continue
        r = hits[-1]
        rd, wr = gb(r[ir], units[ir]), gb(r[iw], units[iw])
        rec = {"kernel": "bq::" + frag, "kernel_regex": regex, "sass_sha256": kernel_sass_hash(LIB, regex),
               "source": f"profiles/{os.path.basename(rep)} (ncu --set full --clock-control none, {rows_per_launch} rows per launch)",
               "duration_under_ncu": f"{r[it]} {units[it]}", "dram_bytes_read": rd, "dram_bytes_write": wr,
               "dram_bytes_per_launch": rd + wr, "algorithmic_bytes_per_launch": bpr * rows_per_launch}
        json.dump(rec, open(os.path.join(ROOT, "profiles", fname), "w"), indent=1)
        print(key, json.dumps(rec))


if __name__ == "__main__":
    main()
