"""Configuration 1 end to end: `<binary> orders.csv --sql Q1` wall time, product CLI beside the reference CLI (same file, same
statement, outputs compared).  Needs a GPU for the product binary.  python scripts/c1_cli.py [--rows 1000000]"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import datagen  # noqa: E402  (measurement harness: data restatement only)

SQL = ("SELECT order_date, SUM(total) AS revenue FROM table WHERE status = 'COMPLETE' AND order_date >= 20240101 "
       "AND order_date <= 20240131 GROUP BY order_date ORDER BY order_date")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=1_000_000)
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    cols = {name: arr for name, _, arr in datagen.host_table(datagen.orders_schema(a.rows), a.rows, seed=20240101)}
    status = np.array(datagen.STATUS_DICT)[cols["status"]]
    path = os.path.join(tempfile.mkdtemp(), "orders.csv")
    total = np.char.mod("%.2f", cols["total"])
    with open(path, "w") as f:
        f.write("order_id,status,order_date,total\n")
        for i in range(0, a.rows, 100000):
            j = min(a.rows, i + 100000)
            f.write("\n".join(f"{o},{s},{d},{t}" for o, s, d, t in zip(cols["order_id"][i:j], status[i:j], cols["order_date"][i:j], total[i:j])) + "\n")
    out = {"config": "C1: CLI, orders CSV, Q1", "rows": a.rows, "file_mb": os.path.getsize(path) / 1e6, "sql": SQL}
    texts = {}
    for name, binary in (("ours", os.path.join(ROOT, "bo-sql_b200", "bq_b200")), ("reference", os.path.join(ROOT, "oracle", "_ref", "bq_ref"))):
        if not os.path.exists(binary):
            continue
        times = []
        for _ in range(a.reps):
            t0 = time.perf_counter()
            r = subprocess.run([binary, path, "--sql", SQL, "--output-format", "csv"], capture_output=True, text=True, timeout=300)
            times.append(time.perf_counter() - t0)
            assert r.returncode == 0 and r.stdout.count("\n") > 5, r.stderr[-500:]
        texts[name] = r.stdout
        if name == "ours":      # where the wall time went: load / CUDA context (concurrent) / statement, printed by the binary itself
            tr = subprocess.run([binary, path, "--sql", SQL, "--output-format", "csv"], capture_output=True, text=True, timeout=300,
                                env=dict(os.environ, BOSQL_TRACE="1"))
            out["ours_phases"] = [ln for ln in tr.stderr.splitlines() if "bosql trace" in ln][-1:]
        out[name + "_wall_s"] = statistics.median(times)
        out[name + "_wall_all"] = times
    if len(texts) == 2:
        a_rows = [l.split(",") for l in texts["ours"].strip().splitlines()]
        b_rows = [l.split(",") for l in texts["reference"].strip().splitlines()]
        assert len(a_rows) == len(b_rows) and a_rows[0] == b_rows[0]
        for x, y in zip(a_rows[1:], b_rows[1:]):
            assert x[0] == y[0] and abs(float(x[1]) - float(y[1])) <= 2e-6 + 1e-12 * abs(float(y[1])), (x, y)
        out["checked"] = "same rows (dates exact, revenue within the six printed decimals)"
        out["speedup"] = out["reference_wall_s"] / out["ours_wall_s"]
    print(json.dumps(out))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "c1_cli.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
