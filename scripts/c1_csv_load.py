"""Configuration 1's load step on the host CPU: load_csv of a 1 M-row orders file (the reference spends ~1.85 s of its ~1.9 s
CLI run here).  Times the product's loader (bqx_catalog_load_csv: bo-sql_b200/host/csv_ingest.cpp) beside the
compiled reference's (oracle/_ref) on the same file and checks that both infer the same schema.  CPU only - no GPU involved.

    python scripts/c1_csv_load.py [--rows 1000000] [--out profiles/c1_csv_load_r1.json]
"""
import argparse
import json
import os
import statistics
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=1_000_000)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    bq = load_package()
    from oracle import datagen, ref_engine        # checker + data restatement: this script is a measurement harness, not product
    cols = {name: arr for name, _, arr in datagen.host_table(datagen.orders_schema(a.rows), a.rows, seed=20240101)}
    status = np.array(datagen.STATUS_DICT)[cols["status"]]
    path = os.path.join(tempfile.mkdtemp(), "orders.csv")
    t0 = time.perf_counter()
    with open(path, "w") as f:
        f.write("order_id,status,order_date,total\n")
        total = np.char.mod("%.2f", cols["total"])
        for i in range(0, a.rows, 100000):
            j = min(a.rows, i + 100000)
            f.write("\n".join(f"{o},{s},{d},{t}" for o, s, d, t in zip(cols["order_id"][i:j], status[i:j], cols["order_date"][i:j], total[i:j])) + "\n")
    size = os.path.getsize(path)
    print(f"wrote {path}: {size / 1e6:.1f} MB in {time.perf_counter() - t0:.1f}s", file=sys.stderr)

    def ours():
        eng = bq.Engine()
        t = time.perf_counter()
        eng.load_csv(path, "table")
        dt = time.perf_counter() - t
        return dt, eng.table_columns("table")

    def ref():
        eng = ref_engine.RefEngine()
        t = time.perf_counter()
        eng.load_csv(path, "table")
        dt = time.perf_counter() - t
        return dt, eng.table_columns("table")

    o = [ours() for _ in range(a.reps)]
    out = {"config": "C1 load step: load_csv of the orders CSV", "rows": a.rows, "file_mb": size / 1e6, "host": f"{os.cpu_count()} logical CPUs, one thread used",
           "ours_s": statistics.median(x[0] for x in o), "ours_mb_per_s": size / 1e6 / statistics.median(x[0] for x in o)}
    schema = [(c[0], c[1]) for c in o[0][1]]
    out["schema"] = schema
    if ref_engine.available():
        r = [ref() for _ in range(max(1, a.reps // 2))]
        out["reference_s"] = statistics.median(x[0] for x in r)
        out["speedup"] = out["reference_s"] / out["ours_s"]
        ref_schema = [(c[0], c[1]) for c in r[0][1]]
        assert schema == ref_schema, (schema, ref_schema)
        for co, cr in zip(o[0][1], r[0][1]):
            assert np.array_equal(co[2], cr[2]), co[0]
            assert co[3:] == cr[3:], (co[0], co[3:], cr[3:])
        out["checked"] = "same types, values, dictionary ids and min/max/ndv as the compiled reference loader"
    print(json.dumps(out))
    if a.out:
        with open(os.path.join(ROOT, a.out), "w") as f:
            json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
