"""Seeded random statements through the GPU operator layer against the compiled reference; prints every disagreement instead of
stopping at the first (a discovery tool; the pinned subset lives in tests/test_fuzz_gpu.py).  python scripts/fuzz_gpu.py [n]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402
from oracle import ref_engine  # noqa: E402  (checker)
from tests import golden_util as G  # noqa: E402
from tests.golden import cases  # noqa: E402
from tests.parity import assert_same_rows  # noqa: E402
from tests.test_oracle_fuzz import _statement, _sweep_statement  # noqa: E402

bq = load_package()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
bad = 0


def run(family, tables, gen, seed):
    global bad
    r, g = ref_engine.RefEngine(), bq.Engine()
    for eng in (r, g):
        dicts = {}
        for name, cols, dname in tables:
            d = dicts.setdefault(dname, eng.new_dict(cases.DICTS[dname])) if dname else None
            eng.add_table(name, cols, d)
    rng = np.random.default_rng(seed)
    for i in range(n):
        sql = gen(rng)
        sql = sql[0] if isinstance(sql, tuple) else sql
        try:
            want, werr = r.query(sql), None
        except RuntimeError as e:
            want, werr = None, str(e)
        try:
            got, gerr = g.query(sql), None
        except Exception as e:  # noqa: BLE001
            got, gerr = None, str(e)
        try:
            if werr or gerr:
                assert werr == gerr, f"errors differ: reference {werr!r}, GPU {gerr!r}"
            else:
                assert got.names == want.names and got.types == want.types, "schema differs"
                order = [(0, True), (1, True)] if " ORDER BY l.order_id, l.sku" in sql else G.order_spec(sql, want.names)
                assert_same_rows(got.cols, want.cols, ordered_by=order, what="rows")
        except AssertionError as e:
            bad += 1
            print(f"[{family} #{i}] {sql}\n    {str(e)[:300]}", flush=True)


star = [t for t in cases.star_tables() if t[0] in ("orders", "lineitem")]
run("star", star, _statement, 20240101)
run("sweep", cases.sweep_tables(), _sweep_statement, 7)
print(f"{bad} disagreement(s) in {2 * n} statements")
