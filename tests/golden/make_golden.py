"""Generates tests/golden/ref_vectors.json by running the COMPILED REFERENCE (oracle/_ref) in this container.

    python tests/golden/make_golden.py

The reference cannot travel to the GPU box as source (and /root/reference does not exist there), so its answers on a
fixed set of tables and queries are frozen here as small fixtures.  They pin the numpy oracle (tests/test_oracle.py) and
the product's SQL front end (tests/test_frontend.py) without a GPU, and the GPU operators (tests/test_golden_gpu.py).
Tables are rebuilt from oracle/datagen.py specs (seeded, counter-based), whose first values are frozen too.
DOUBLE values are stored as float.hex() so the fixtures are bit-exact.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import datagen, ref_engine  # noqa: E402
from tests.golden import cases  # noqa: E402


def enc(col):
    if col.dtype.kind == "f":
        return [float(x).hex() for x in col.tolist()]
    return [int(x) for x in col.tolist()]


def main():
    out = {"tables": {}, "queries": [], "explain": []}
    engines = {}
    for tset, builder in cases.TABLE_SETS.items():
        eng = ref_engine.RefEngine()
        tables = builder()
        for name, cols, dict_key in tables:
            d = eng.new_dict(cases.DICTS[dict_key]) if dict_key else None
            # tables of one set that name the same dictionary share it
            eng._shared = getattr(eng, "_shared", {})
            if dict_key:
                d = eng._shared.setdefault(dict_key, d)
            eng.add_table(name, cols, d)
        engines[tset] = eng
        out["tables"][tset] = {name: {c[0]: enc(np.asarray(c[2])[:5]) for c in cols} for name, cols, _ in tables}
    for tset, sql in cases.QUERIES:
        entry = {"tables": tset, "sql": sql}
        try:
            r = engines[tset].query(sql)
            entry.update(names=r.names, types=r.types, rows=r.rows, cols=[enc(c) for c in r.cols], has_dict=r.has_dict)
        except RuntimeError as e:
            entry["error"] = str(e)
        out["queries"].append(entry)
    for sql in cases.EXPLAIN:
        entry = {"sql": sql}
        try:
            entry["plan"] = engines["fixture"].explain(sql)
        except RuntimeError as e:
            entry["error"] = str(e)
        out["explain"].append(entry)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_vectors.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=0, separators=(",", ":"))
    print(f"wrote {path}: {len(out['queries'])} queries, {len(out['explain'])} plans, {os.path.getsize(path)} bytes")


if __name__ == "__main__":
    main()
