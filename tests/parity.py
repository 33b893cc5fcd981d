"""Parity helpers shared by the tests, __graft_entry__.smoke() and bench.py's checker legs.

Comparison contract (hazards H1-H4):
  * INT64 / COUNT / StrId / Date32 columns and group keys: bit-exact;
  * DOUBLE SUM / AVG: |got - want| <= REL_TOL * |want| (order of addition differs);
  * no ORDER BY  -> rows compared as a multiset keyed by the non-DOUBLE columns (emit order is unspecified);
  * ORDER BY     -> the sort-key sequence must match exactly; payload within ties is unordered.
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

REL_TOL = 1e-12      # BASELINE.json north_star: DOUBLE SUM/AVG within 1e-12 relative

INT64, DOUBLE, STRING, DATE32 = 0, 1, 2, 3


def _is_float(a):
    return a.dtype.kind == "f"


def assert_close(got, want, what=""):
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    assert got.shape == want.shape, f"{what}: shape {got.shape} != {want.shape}"
    if got.size == 0:
        return
    both_inf = np.isinf(got) & np.isinf(want) & (np.sign(got) == np.sign(want))
    err = np.abs(got - want)
    lim = REL_TOL * np.abs(want)
    bad = ~((err <= lim) | both_inf)
    assert not bad.any(), (f"{what}: {bad.sum()} values outside {REL_TOL} relative; first: "
                           f"got {got[bad][0]!r} want {want[bad][0]!r}")


def _sort_rows(cols, key_idx):
    if not key_idx:
        return cols
    order = np.lexsort([cols[i] for i in reversed(key_idx)])
    return [c[order] for c in cols]


def assert_same_rows(got_cols, want_cols, ordered_by=None, what=""):
    """got_cols / want_cols: lists of numpy arrays (one per output column, same column order).

    ordered_by: None (multiset compare) or list of (column index, asc) describing the ORDER BY.
    """
    assert len(got_cols) == len(want_cols), f"{what}: {len(got_cols)} columns != {len(want_cols)}"
    n = len(want_cols[0]) if want_cols else 0
    for g, w in zip(got_cols, want_cols):
        assert len(g) == n == len(w), f"{what}: row count {len(g)} != {len(w)}"
        assert g.dtype == w.dtype, f"{what}: dtype {g.dtype} != {w.dtype}"
    if n == 0:
        return
    exact = [i for i, w in enumerate(want_cols) if not _is_float(w)]
    if ordered_by:
        # the sort-key sequence is pinned (ties: payload order is unspecified, H4) ...
        for i, _asc in ordered_by:
            if _is_float(want_cols[i]):
                assert_close(got_cols[i], want_cols[i], f"{what}: sort key column {i}")
            else:
                assert np.array_equal(got_cols[i], want_cols[i]), f"{what}: sort key column {i} differs"
    # ... and the rows are the same multiset: canonical order = exact columns first, then the DOUBLE columns
    # (group keys of type DOUBLE are bit-exact; DOUBLE aggregates only break ties among otherwise equal rows)
    floats = [i for i in range(len(want_cols)) if i not in exact]
    def canon(cols):
        order = np.lexsort([cols[i] for i in reversed(exact + floats)])
        return [c[order] for c in cols]
    g2, w2 = canon(got_cols), canon(want_cols)
    for i, (g, w) in enumerate(zip(g2, w2)):
        if _is_float(w):
            assert_close(g, w, f"{what}: column {i}")
        else:
            assert np.array_equal(g, w), f"{what}: column {i} differs (first rows got {g[:5]} want {w[:5]})"


# ---- smoke: one tiny Q1-shaped pipeline on cuda:0 through the C ABI -----------------------------------
def q1_kernel_spec(bq, status_col, date_col, total_col, n, status_id, d_lo, d_hi, date_min, date_max):
    """The fused-kernel spec of Q1: status = id AND date BETWEEN d_lo AND d_hi GROUP BY date SUM(total)."""
    s = bq.ScanSpec()
    s.key = bq.make_slot(date_col, [(d_lo, d_hi, 0)])
    s.a = bq.make_slot(total_col)
    s.pred[0] = bq.make_slot(status_col, [(status_id, status_id, 0)])
    s.row_begin, s.row_end = 0, n
    s.n_v = 1
    s.v[0] = bq.VExpr(op=bq.V_A)
    s.group_mode = bq.GROUP_DENSE
    s.key_min, s.key_max = date_min, date_max
    s.n_out = 1
    s.out[0] = bq.AggOut(func=bq.AGG_SUM, v=0, as_int=0)
    return s


def run_smoke(bq):
    """One small Q1 on cuda:0, twice: through the kernel layer's C ABI (bq_scan_aggregate + bq_rel_sort) and through the
    operator layer (bqx_*: parse -> plan -> open / next / close), both checked against the ORACLE - the numpy restatement of
    the reference's executor (oracle/oracle.py, itself pinned to the compiled reference by tests/test_oracle.py)."""
    from oracle import datagen, oracle as ORA
    n, seed = 200_000, 7
    tab = datagen.host_table(datagen.orders_schema(n), n, seed)
    cols = {name: arr for name, _t, arr in tab}
    sql = ("SELECT order_date, SUM(total) AS revenue FROM orders WHERE status = 'COMPLETE' AND order_date >= 20240101 "
           "AND order_date <= 20240131 GROUP BY order_date ORDER BY order_date")
    ora = ORA.Oracle()
    ora.add_table("orders", tab, ora.new_dict(datagen.STATUS_DICT))
    want = ora.query(sql)
    assert want.rows > 0
    ctx = bq.Context(0)
    try:
        status = ctx.upload(STRING, cols["status"])
        date = ctx.upload(DATE32, cols["order_date"])
        total = ctx.upload(DOUBLE, cols["total"])
        spec = q1_kernel_spec(bq, status, date, total, n, 0, 20240101, 20240131, 20240101, 20241228)
        rel = ctx.scan_aggregate(spec)
        srt = ctx.rel_sort(rel, [0], [1])
        got = srt.to_numpy()
        assert_same_rows(got, want.cols, ordered_by=[(0, True)], what="smoke, kernel layer vs oracle")
    finally:
        ctx.close()
    eng = bq.Engine(0)
    eng.add_table("orders", tab, eng.new_dict(datagen.STATUS_DICT))
    res = eng.query(sql)
    assert res.names == want.names, (res.names, want.names)
    assert_same_rows(res.cols, want.cols, ordered_by=[(0, True)], what="smoke, operator layer vs oracle")
