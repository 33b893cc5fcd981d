"""Loading tests/golden/ref_vectors.json (answers of the compiled reference, frozen by tests/golden/make_golden.py)."""
import json
import os

import numpy as np

from tests.golden import cases

NP = {0: np.int64, 1: np.float64, 2: np.uint32, 3: np.int32}
PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_vectors.json")


def load():
    with open(PATH) as f:
        return json.load(f)


def decode(entry):
    cols = []
    for t, c in zip(entry["types"], entry["cols"]):
        if t == 1:
            cols.append(np.array([float.fromhex(x) for x in c], dtype=np.float64))
        else:
            cols.append(np.array(c, dtype=NP[t]))
    return cols


def build_engines(make_engine):
    """One engine per table set, filled with the fixture tables; dictionaries shared per set as in the generator."""
    out = {}
    for tset, builder in cases.TABLE_SETS.items():
        eng = make_engine()
        shared = {}
        for name, cols, dict_key in builder():
            d = None
            if dict_key:
                if dict_key not in shared:
                    shared[dict_key] = eng.new_dict(cases.DICTS[dict_key])
                d = shared[dict_key]
            eng.add_table(name, cols, d)
        out[tset] = eng
    return out


def order_spec(sql, names):
    if " ORDER BY " not in sql:
        return None
    tail = sql.split(" ORDER BY ", 1)[1].split(" LIMIT ")[0]
    spec = []
    for item in tail.split(","):
        parts = item.strip().split()
        if parts[0] not in names:
            return None
        spec.append((names.index(parts[0]), not (len(parts) > 1 and parts[1] == "DESC")))
    return spec
