"""Kernel-layer parity on the GPU: every call goes through the C ABI of include/bosql_b200.h.

Oracles: oracle/datagen.py (generator restatement), the compiled reference executor (oracle/_ref, through SQL on the
same arrays) and plain numpy restatements of the operator being tested.
"""
import numpy as np
import pytest

from oracle import datagen
from tests.parity import DATE32, DOUBLE, INT64, STRING, assert_close, assert_same_rows, q1_kernel_spec

pytestmark = pytest.mark.gpu


# ---- generator -----------------------------------------------------------------------------------
@pytest.mark.parametrize("schema_fn,n", [(lambda: datagen.orders_schema(1000), 10007),
                                          (lambda: datagen.lineitem_schema(1000), 4099),
                                          (lambda: datagen.sweep_schema(), 5003)])
def test_generator_matches_numpy(bq, ctx, schema_fn, n):
    schema = schema_fn()
    for row0 in (0, 123456789):
        host = datagen.host_table(schema, n, seed=42, row0=row0)
        for i, (name, typ, spec) in enumerate(schema):
            col = ctx.alloc(typ, n).generate(seed=42, stream=i, row0=row0, **spec)
            got = col.to_numpy()
            assert got.dtype == host[i][2].dtype
            assert np.array_equal(got, host[i][2]), f"{name} differs at row0={row0}"


def test_generator_table_and_hashed(bq, ctx):
    n = 20011
    cdf = datagen.zipf_cdf(1000, 1.1)
    col = ctx.alloc(INT64, n).generate(dist=bq.GEN_TABLE, seed=5, stream=3, lo=1, cdf=cdf)
    want = datagen.generate(INT64, n, datagen.GEN_TABLE, 5, 3, lo=1, cdf=cdf)
    assert np.array_equal(col.to_numpy(), want)
    col = ctx.alloc(INT64, n).generate(dist=bq.GEN_HASHED, seed=5, stream=4, lo=0, hi=9999, modulus=1 << 61)
    want = datagen.generate(INT64, n, datagen.GEN_HASHED, 5, 4, lo=0, hi=9999, modulus=1 << 61)
    assert np.array_equal(col.to_numpy(), want)
    assert len(np.unique(want)) <= 10000
    # Zipf(1.1) over a whole 50 M-key domain: exact head, geometric buckets in the tail, uniform inside a bucket
    cdf, starts = datagen.zipf_buckets(50_000_000, 1.1, head=1 << 12)
    col = ctx.alloc(INT64, n).generate(dist=bq.GEN_BUCKETS, seed=5, stream=5, lo=1, cdf=cdf, starts=starts, row0=12345)
    want = datagen.generate(INT64, n, datagen.GEN_BUCKETS, 5, 5, lo=1, cdf=cdf, starts=starts, row0=12345)
    assert np.array_equal(col.to_numpy(), want)
    assert want.min() >= 1 and want.max() <= 50_000_000 and want.max() > 1 << 20


def test_minmax_and_f64_key(bq, ctx):
    rng = np.random.default_rng(1)
    a = rng.normal(size=10001) * 1e6
    col = ctx.upload(DOUBLE, a)
    lo, hi = col.minmax()
    assert lo == bq.f64_key(a.min()) and hi == bq.f64_key(a.max())
    keys = np.array([bq.f64_key(x) for x in a[:500]])
    assert np.array_equal(np.argsort(keys, kind="stable"), np.argsort(a[:500], kind="stable"))
    assert bq.f64_key(-0.0) == bq.f64_key(0.0) == 0
    b = rng.integers(-10**12, 10**12, size=7777)
    assert ctx.upload(INT64, b).minmax() == (b.min(), b.max())


# ---- fused filter + global aggregate (configuration 2) ----------------------------------------------
def _sweep_cols(ctx, n, seed):
    host = {name: arr for name, _t, arr in datagen.host_table(datagen.sweep_schema(), n, seed)}
    types = {name: t for name, t, _ in datagen.sweep_schema()}
    dev = {name: ctx.upload(types[name], arr) for name, arr in host.items()}
    return host, dev


def _range_for(bq, op, lit, is_f):
    """compare_values (src/exec/expression.cpp:60-120) reduced to a key range, as the host compiler does."""
    k = bq.f64_key(lit) if is_f else int(lit)
    lo_all, hi_all = -(1 << 63), (1 << 63) - 1
    if is_f:
        lo_all, hi_all = bq.f64_key(-np.inf), bq.f64_key(np.inf)
    return {"<": (lo_all, k - 1, 0), "<=": (lo_all, k, 0), ">": (k + 1, hi_all, 0), ">=": (k, hi_all, 0),
            "=": (k, k, 0), "!=": (k, k, 1)}[op]


@pytest.mark.parametrize("n", [0, 1, 3, 127, 131, 1024, 100003])
def test_filter_agg_sizes(bq, ctx, n):
    host, dev = _sweep_cols(ctx, n, seed=11)
    s = bq.ScanSpec()
    s.a = bq.make_slot(dev["v"])
    s.b = bq.make_slot(dev["w"])
    s.pred[0] = bq.make_slot(dev["c_i64"], [_range_for(bq, "<", 500000, False)])
    s.row_begin, s.row_end = 0, n
    s.n_v = 2
    s.v[0] = bq.VExpr(op=bq.V_A)
    s.v[1] = bq.VExpr(op=bq.V_B)
    s.group_mode = bq.GROUP_NONE
    s.n_out = 4
    s.out[0] = bq.AggOut(func=bq.AGG_COUNT)
    s.out[1] = bq.AggOut(func=bq.AGG_SUM, v=0)
    s.out[2] = bq.AggOut(func=bq.AGG_SUM, v=1, as_int=1)
    s.out[3] = bq.AggOut(func=bq.AGG_AVG, v=0)
    got = ctx.scan_aggregate(s).to_numpy()
    m = host["c_i64"] < 500000
    if m.sum() == 0:
        assert all(len(c) == 0 for c in got)       # zero qualifying rows -> zero output rows (H5)
        return
    assert got[0][0] == m.sum()
    assert_close(got[1], [host["v"][m].sum()], "SUM(v)")
    assert got[2][0] == host["w"][m].sum()
    assert_close(got[3], [host["v"][m].sum() / m.sum()], "AVG(v)")


@pytest.mark.parametrize("pred,op,lit", [
    ("c_i64", "<", 10000), ("c_i64", ">=", 990000), ("c_i64", "=", 697221), ("c_i64", "!=", 5),
    ("c_f64", "<", 5000.0), ("c_f64", ">", 9899.995), ("c_f64", "<=", 1146.09), ("c_f64", "!=", 1146.09),
    ("c_str", "=", 7), ("c_str", "!=", 7),
    ("c_date", ">=", 20200101), ("c_date", "<", 20150201), ("c_date", "=", 20170925),
])
def test_filter_agg_predicate_types(bq, ctx, pred, op, lit):
    n = 300007
    host, dev = _sweep_cols(ctx, n, seed=7)
    is_f = pred == "c_f64"
    s = bq.ScanSpec()
    s.a = bq.make_slot(dev["v"])
    s.pred[0] = bq.make_slot(dev[pred], [_range_for(bq, op, lit, is_f)])
    s.row_begin, s.row_end = 0, n
    s.n_v = 1
    s.v[0] = bq.VExpr(op=bq.V_A)
    s.n_out = 2
    s.out[0] = bq.AggOut(func=bq.AGG_COUNT)
    s.out[1] = bq.AggOut(func=bq.AGG_SUM, v=0)
    got = ctx.scan_aggregate(s).to_numpy()
    x = host[pred]
    m = {"<": x < lit, "<=": x <= lit, ">": x > lit, ">=": x >= lit, "=": x == lit, "!=": x != lit}[op]
    assert m.sum() > 0
    assert got[0][0] == m.sum()
    assert_close(got[1], [host["v"][m].sum()], "SUM(v)")


def test_filter_agg_vs_reference_sql(bq, ctx, ref):
    n = 50021
    host, dev = _sweep_cols(ctx, n, seed=3)
    eng = ref.RefEngine()
    eng.add_table("t", [(name, t, host[name]) for name, t, _ in datagen.sweep_schema()])
    r = eng.query("SELECT COUNT(*), SUM(v), SUM(w), AVG(v) FROM t WHERE c_date >= 20180101 AND c_date <= 20191231 AND c_str != 3")
    s = bq.ScanSpec()
    s.a = bq.make_slot(dev["v"])
    s.b = bq.make_slot(dev["w"])
    s.pred[0] = bq.make_slot(dev["c_date"], [(20180101, 20191231, 0)])
    s.pred[1] = bq.make_slot(dev["c_str"], [(3, 3, 1)])
    s.row_begin, s.row_end = 0, n
    s.n_v = 2
    s.v[0] = bq.VExpr(op=bq.V_A)
    s.v[1] = bq.VExpr(op=bq.V_B)
    s.n_out = 4
    s.out[0] = bq.AggOut(func=bq.AGG_COUNT)
    s.out[1] = bq.AggOut(func=bq.AGG_SUM, v=0)
    s.out[2] = bq.AggOut(func=bq.AGG_SUM, v=1, as_int=1)
    s.out[3] = bq.AggOut(func=bq.AGG_AVG, v=0)
    got = ctx.scan_aggregate(s).to_numpy()
    assert_same_rows(got, r.cols, what="filter+agg vs reference")


def test_filter_agg_subrange_and_mask(bq, ctx):
    n = 70001
    host, dev = _sweep_cols(ctx, n, seed=5)
    mask_h = (host["c_str"] % 3 == 0).astype(np.int64)
    mask = ctx.upload(INT64, mask_h)
    for rb, re in [(1, n), (5, 69999), (130, 131), (4, 4), (1000, 50000)]:
        s = bq.ScanSpec()
        s.a = bq.make_slot(dev["v"], [_range_for(bq, ">", 100.0, True)])
        s.mask = mask.h
        s.row_begin, s.row_end = rb, re
        s.n_v = 1
        s.v[0] = bq.VExpr(op=bq.V_MUL, l_src=bq.L_A, r_src=bq.R_IMM, imm_i=3)
        s.n_out = 2
        s.out[0] = bq.AggOut(func=bq.AGG_COUNT)
        s.out[1] = bq.AggOut(func=bq.AGG_SUM, v=0)
        got = ctx.scan_aggregate(s).to_numpy()
        v = host["v"][rb:re]
        m = (v > 100.0) & (mask_h[rb:re] != 0)
        if m.sum() == 0:
            assert len(got[0]) == 0
            continue
        assert got[0][0] == m.sum()
        assert_close(got[1], [(v[m] * 3.0).sum()], f"SUM(v*3) [{rb},{re})")


# ---- Q1 shape: dense GROUP BY in shared memory -------------------------------------------------------
@pytest.mark.parametrize("n", [1000, 250_003])
def test_q1_kernel_vs_reference(bq, ctx, ref, n):
    tab = datagen.host_table(datagen.orders_schema(n), n, seed=2024)
    cols = {name: arr for name, _t, arr in tab}
    eng = ref.RefEngine()
    d = eng.new_dict(datagen.STATUS_DICT)
    eng.add_table("orders", tab, d)
    r = eng.query("SELECT order_date, SUM(total) AS revenue FROM orders WHERE status = 'COMPLETE' AND "
                  "order_date >= 20240101 AND order_date <= 20240131 GROUP BY order_date ORDER BY order_date")
    status = ctx.upload(STRING, cols["status"])
    date = ctx.upload(DATE32, cols["order_date"])
    total = ctx.upload(DOUBLE, cols["total"])
    lo, hi = date.minmax()
    spec = q1_kernel_spec(bq, status, date, total, n, 0, 20240101, 20240131, lo, hi)
    rel = ctx.scan_aggregate(spec)
    got = ctx.rel_sort(rel, [0], [1]).to_numpy()
    assert_same_rows(got, r.cols, ordered_by=[(0, True)], what="Q1")


def test_group_dense_global_and_hash_agree(bq, ctx):
    """The same GROUP BY through the three table kinds (smem / dense global / hash) gives the same rows."""
    n = 400_009
    rng = np.random.default_rng(9)
    key = rng.integers(-50, 20000, size=n).astype(np.int64)
    val = rng.integers(1, 1000, size=n).astype(np.float64) / 8.0      # dyadic: sums are exact in any order
    kc, vc = ctx.upload(INT64, key), ctx.upload(DOUBLE, val)
    uk, inv = np.unique(key, return_inverse=True)
    want = [uk, np.bincount(inv).astype(np.int64), np.bincount(inv, weights=val), np.bincount(inv, weights=val) / np.bincount(inv)]
    for mode, kmin, kmax in [(bq.GROUP_DENSE, -50, 19999), (bq.GROUP_HASH, 0, -1)]:
        s = bq.ScanSpec()
        s.key = bq.make_slot(kc)
        s.a = bq.make_slot(vc)
        s.row_begin, s.row_end = 0, n
        s.n_v = 1
        s.v[0] = bq.VExpr(op=bq.V_A)
        s.group_mode = mode
        s.key_min, s.key_max = kmin, kmax
        s.ndv_hint = 20050
        s.n_out = 3
        s.out[0] = bq.AggOut(func=bq.AGG_COUNT)
        s.out[1] = bq.AggOut(func=bq.AGG_SUM, v=0)
        s.out[2] = bq.AggOut(func=bq.AGG_AVG, v=0)
        got = ctx.scan_aggregate(s).to_numpy()
        assert_same_rows(got, want, what=f"group mode {mode}")
        assert np.array_equal(np.sort(got[0]), uk)
        g = dict(zip(got[0].tolist(), got[2].tolist()))
        assert all(g[k] == w for k, w in zip(uk.tolist(), want[2].tolist())), "dyadic sums must be bit-exact"
    # small domain -> shared-memory tables
    key2 = (key % 300).astype(np.int64)
    kc2 = ctx.upload(INT64, key2)
    s = bq.ScanSpec()
    s.key = bq.make_slot(kc2)
    s.a = bq.make_slot(vc)
    s.row_begin, s.row_end = 0, n
    s.n_v = 1
    s.v[0] = bq.VExpr(op=bq.V_A)
    s.group_mode = bq.GROUP_DENSE
    s.key_min, s.key_max = 0, 299
    s.n_out = 2
    s.out[0] = bq.AggOut(func=bq.AGG_COUNT)
    s.out[1] = bq.AggOut(func=bq.AGG_SUM, v=0)
    got = ctx.scan_aggregate(s).to_numpy()
    uk2, inv2 = np.unique(key2, return_inverse=True)
    assert np.array_equal(got[0], uk2)
    assert np.array_equal(got[1], np.bincount(inv2))
    assert np.array_equal(got[2], np.bincount(inv2, weights=val))


def test_group_hash_special_keys(bq, ctx):
    key = np.array([np.iinfo(np.int64).min, 5, np.iinfo(np.int64).min, np.iinfo(np.int64).max, 5, 0], dtype=np.int64)
    val = np.array([1.0, 2.0, 3.0, 4.0, 5.0, 6.0])
    s = bq.ScanSpec()
    s.key = bq.make_slot(ctx.upload(INT64, key))
    s.a = bq.make_slot(ctx.upload(DOUBLE, val))
    s.row_begin, s.row_end = 0, len(key)
    s.n_v = 1
    s.v[0] = bq.VExpr(op=bq.V_A)
    s.group_mode = bq.GROUP_HASH
    s.n_out = 1
    s.out[0] = bq.AggOut(func=bq.AGG_SUM, v=0)
    # keep the uploaded columns alive through the call
    kc, vc = ctx.upload(INT64, key), ctx.upload(DOUBLE, val)
    s.key, s.a = bq.make_slot(kc), bq.make_slot(vc)
    got = ctx.scan_aggregate(s).to_numpy()
    want = [np.array([np.iinfo(np.int64).min, 0, 5, np.iinfo(np.int64).max], dtype=np.int64), np.array([4.0, 6.0, 7.0, 4.0])]
    assert_same_rows(got, want)


# ---- selection vectors, gather, slice ------------------------------------------------------------------
@pytest.mark.parametrize("n", [0, 1, 31, 32, 33, 8191, 100_001])
def test_select_gather(bq, ctx, n):
    host, dev = _sweep_cols(ctx, n, seed=21)
    ids = ctx.select([bq.make_slot(dev["c_i64"], [_range_for(bq, "<", 300000, False)]),
                      bq.make_slot(dev["c_str"], [(10, 60, 0)])], row_begin=0, row_end=n)
    m = (host["c_i64"] < 300000) & (host["c_str"] >= 10) & (host["c_str"] <= 60)
    want_ids = np.nonzero(m)[0].astype(np.uint32)
    assert np.array_equal(ids.to_numpy(), want_ids)       # stable: scan order
    for name in ("v", "c_date", "w", "c_str"):
        assert np.array_equal(ctx.gather(dev[name], ids).to_numpy(), host[name][m])
    if n > 40:
        assert np.array_equal(ctx.slice(dev["w"], 7, n - 3).to_numpy(), host["w"][7:n - 3])
        sub = ctx.select([bq.make_slot(dev["c_i64"], [_range_for(bq, "<", 300000, False)])], row_begin=33, row_end=n - 5)
        want = (np.nonzero(host["c_i64"][33:n - 5] < 300000)[0] + 33).astype(np.uint32)
        assert np.array_equal(sub.to_numpy(), want)


# ---- joins -----------------------------------------------------------------------------------------
def test_join_kinds_and_probe(bq, ctx):
    rng = np.random.default_rng(3)
    nb, npb = 5000, 40_003
    bkey = rng.permutation(np.arange(100, 100 + nb)).astype(np.int64)          # unique, dense
    pkey = rng.integers(0, 100 + nb + 200, size=npb).astype(np.int64)
    bk, pk = ctx.upload(INT64, bkey), ctx.upload(INT64, pkey)
    pos = {k: i for i, k in enumerate(bkey.tolist())}
    want_p = np.array([i for i, k in enumerate(pkey.tolist()) if k in pos], dtype=np.uint32)
    want_b = np.array([pos[k] for k in pkey.tolist() if k in pos], dtype=np.uint32)
    for kind, expect in [(bq.JOIN_AUTO, bq.JOIN_DIRECT), (bq.JOIN_DIRECT, bq.JOIN_DIRECT), (bq.JOIN_HASH, bq.JOIN_HASH)]:
        j = ctx.join_build(bk, kind=kind, need_rows=True, unique=True, key_min=100, key_max=100 + nb - 1)
        assert j.kind == expect
        p, b = ctx.join_probe(j, pk)
        assert np.array_equal(p.to_numpy(), want_p) and np.array_equal(b.to_numpy(), want_b)
    j = ctx.join_build(bk, kind=bq.JOIN_AUTO, need_rows=False, unique=True, key_min=100, key_max=100 + nb - 1)
    assert j.kind == bq.JOIN_BITMAP and j.bytes <= (nb // 8) + 8


def test_join_duplicates_insertion_order(bq, ctx):
    rng = np.random.default_rng(4)
    bkey = rng.integers(0, 50, size=400).astype(np.int64)         # heavy duplicates
    pkey = rng.integers(-5, 60, size=3001).astype(np.int64)
    bk, pk = ctx.upload(INT64, bkey), ctx.upload(INT64, pkey)
    j = ctx.join_build(bk, kind=bq.JOIN_AUTO, need_rows=True, key_min=0, key_max=49)    # duplicates force HASH
    assert j.kind == bq.JOIN_HASH
    p, b = ctx.join_probe(j, pk)
    wp, wb = [], []
    for i, k in enumerate(pkey.tolist()):
        for r in np.nonzero(bkey == k)[0].tolist():        # build insertion order (src/exec/operator.cpp:802-816)
            wp.append(i)
            wb.append(r)
    assert np.array_equal(p.to_numpy(), np.array(wp, dtype=np.uint32))
    assert np.array_equal(b.to_numpy(), np.array(wb, dtype=np.uint32))


@pytest.mark.parametrize("sku_type", [INT64, STRING])
def test_q2_kernel_vs_reference(bq, ctx, ref, sku_type):
    n_orders, n_line, n_sku = 20_000, 150_007, 500
    orders = datagen.host_table(datagen.orders_schema(n_orders, prefix="o."), n_orders, seed=1)
    line = datagen.host_table(datagen.lineitem_schema(n_orders, n_sku, sku_type=sku_type), n_line, seed=2)
    eng = ref.RefEngine()
    d = eng.new_dict(datagen.STATUS_DICT + [f"sku{i}" for i in range(n_sku)] if sku_type == STRING else datagen.STATUS_DICT)
    eng.add_table("orders", orders, d)
    eng.add_table("lineitem", line, d)
    r = eng.query("SELECT l.sku, SUM(l.qty * l.price) AS rev FROM lineitem l JOIN orders o ON l.order_id = o.order_id "
                  "WHERE o.status = 'COMPLETE' GROUP BY l.sku ORDER BY rev DESC LIMIT 20")
    o = {name: ctx.upload(t, a) for name, t, a in orders}
    l = {name: ctx.upload(t, a) for name, t, a in line}
    j = ctx.join_build(o["o.order_id"], preds=[bq.make_slot(o["o.status"], [(0, 0, 0)])], unique=True,
                       key_min=1, key_max=n_orders)
    assert j.kind == bq.JOIN_BITMAP
    s = bq.ScanSpec()
    s.key = bq.make_slot(l["l.sku"])
    s.a = bq.make_slot(l["l.qty"])
    s.b = bq.make_slot(l["l.price"])
    s.jkey = bq.make_slot(l["l.order_id"])
    s.join = j.h
    s.row_begin, s.row_end = 0, n_line
    s.n_v = 1
    s.v[0] = bq.VExpr(op=bq.V_MUL)
    s.group_mode = bq.GROUP_DENSE
    s.key_min, s.key_max = 0, n_sku - 1
    s.n_out = 1
    s.out[0] = bq.AggOut(func=bq.AGG_SUM, v=0)
    rel = ctx.scan_aggregate(s)
    top = ctx.rel_sort(rel, [1], [0], limit=20).to_numpy()
    assert_same_rows(top, r.cols, ordered_by=[(1, False)], what="Q2")
    # the same join through DIRECT and HASH tables, status read from the build side is not needed: same result
    for kind in (bq.JOIN_DIRECT, bq.JOIN_HASH):
        j2 = ctx.join_build(o["o.order_id"], preds=[bq.make_slot(o["o.status"], [(0, 0, 0)])], kind=kind, need_rows=True,
                            unique=True, key_min=1, key_max=n_orders)
        s.join = j2.h
        top2 = ctx.rel_sort(ctx.scan_aggregate(s), [1], [0], limit=20).to_numpy()
        assert_same_rows(top2, r.cols, ordered_by=[(1, False)], what=f"Q2 join kind {kind}")


@pytest.mark.parametrize("sku_type", [INT64, STRING])
@pytest.mark.parametrize("n_line", [0, 100, 128, 150_007])
def test_q2_probe_in_key_range_passes(bq, ctx, sku_type, n_line):
    """bq_join_probe_bits (the bitmap cut into slices, one streaming pass over the probe key per slice) followed by the fused
    scan over row bits must equal the single fused probe - at any slice count, incl. ragged tails and an empty probe side."""
    n_orders, n_sku = 200_000, 300
    orders = datagen.host_table(datagen.orders_schema(n_orders, prefix="o."), n_orders, seed=11)
    line = datagen.host_table(datagen.lineitem_schema(n_orders, n_sku, sku_type=sku_type), max(n_line, 1), seed=12)
    o = {name: ctx.upload(t, a) for name, t, a in orders}
    l = {name: ctx.upload(t, a[:n_line]) for name, t, a in line}
    j = ctx.join_build(o["o.order_id"], preds=[bq.make_slot(o["o.status"], [(0, 0, 0)])], unique=True, key_min=1, key_max=n_orders)
    assert j.kind == bq.JOIN_BITMAP and j.popcount() == j.build_rows

    def spec():
        s = bq.ScanSpec()
        s.key = bq.make_slot(l["l.sku"])
        s.a = bq.make_slot(l["l.qty"])
        s.b = bq.make_slot(l["l.price"])
        s.row_begin, s.row_end = 0, n_line
        s.n_v = 1
        s.v[0] = bq.VExpr(op=bq.V_MUL)
        s.group_mode = bq.GROUP_DENSE
        s.key_min, s.key_max = 0, n_sku - 1
        s.n_out = 2
        s.out[0] = bq.AggOut(func=bq.AGG_SUM, v=0)
        s.out[1] = bq.AggOut(func=bq.AGG_COUNT)
        return s
    fused = spec()
    fused.jkey = bq.make_slot(l["l.order_id"])
    fused.join = j.h
    want = ctx.scan_aggregate(fused).to_numpy()
    key = line[0][2][:n_line]
    status = orders[1][2]
    expect_bits = status[key - 1] == 0
    for slice_bytes in (1 << 30, 8 << 10, 1 << 10):            # 1, 4 and 25 passes over the 25 KB bitmap
        bits = j.probe_bits(l["l.order_id"], 0, n_line, slice_bytes=slice_bytes)
        words = bits.to_numpy()
        got_bits = np.unpackbits(words.view(np.uint8), bitorder="little")[:n_line].astype(bool)
        assert np.array_equal(got_bits, expect_bits), f"slice {slice_bytes}"
        assert not np.unpackbits(words.view(np.uint8), bitorder="little")[n_line:].any(), "bits beyond the last row must be clear"
        s = spec()
        s.row_bits = bits.h
        got = ctx.scan_aggregate(s).to_numpy()
        assert_same_rows(got, want, what=f"row bits, slice {slice_bytes}")


def test_join_payload_from_build_side(bq, ctx):
    """SUM(p.v * b.w) with b.w read from the matched build row (configuration 5's shape)."""
    rng = np.random.default_rng(6)
    nb, npb = 3000, 50_001
    bkey = np.arange(1, nb + 1, dtype=np.int64)
    bw = rng.integers(1, 64, size=nb).astype(np.float64) / 4.0
    pkey = rng.integers(1, nb + 500, size=npb).astype(np.int64)
    pv = rng.integers(1, 64, size=npb).astype(np.float64) / 4.0
    bk, bwc, pk, pvc = ctx.upload(INT64, bkey), ctx.upload(DOUBLE, bw), ctx.upload(INT64, pkey), ctx.upload(DOUBLE, pv)
    m = pkey <= nb
    want_cnt, want_sum = m.sum(), (pv[m] * bw[pkey[m] - 1]).sum()
    for kind in (bq.JOIN_DIRECT, bq.JOIN_HASH):
        j = ctx.join_build(bk, kind=kind, need_rows=True, unique=True, key_min=1, key_max=nb)
        s = bq.ScanSpec()
        s.a = bq.make_slot(pvc)
        s.b = bq.make_slot(bwc, from_build=True)
        s.jkey = bq.make_slot(pk)
        s.join = j.h
        s.row_begin, s.row_end = 0, npb
        s.n_v = 1
        s.v[0] = bq.VExpr(op=bq.V_MUL)
        s.n_out = 2
        s.out[0] = bq.AggOut(func=bq.AGG_COUNT)
        s.out[1] = bq.AggOut(func=bq.AGG_SUM, v=0)
        got = ctx.scan_aggregate(s).to_numpy()
        assert got[0][0] == want_cnt and got[1][0] == want_sum       # dyadic values: exact


def test_bitmap_build_without_round_trip(bq, ctx):
    """bq_join_build_bitmap_nosync + bq_join_bitmap_verdict: the counters packed behind the bitmap words say what was inserted;
    a repeated key shows as fewer bits than rows, a key outside the catalog's bounds as flag 2 - the two conditions on which
    ranks that summed their bitmaps fall back to a broadcast join."""
    n = 100_003
    rng = np.random.default_rng(21)
    keys = rng.permutation(np.arange(1, n + 1, dtype=np.int64))
    status = rng.integers(0, 4, size=n).astype(np.uint32)
    kc, sc = ctx.upload(INT64, keys), ctx.upload(STRING, status)
    j = ctx.join_build_bitmap_nosync(kc, preds=[bq.make_slot(sc, [(0, 0, 0)])], key_min=1, key_max=n)
    bits, ins, flags = j.verdict()
    want = int((status == 0).sum())
    assert (bits, ins, flags) == (want, want, 0) and j.build_rows == want
    p, words = j.bitmap()
    assert words == (n + 31) // 32
    # the bits themselves: the same probe answers as the synchronous build
    j2 = ctx.join_build(kc, preds=[bq.make_slot(sc, [(0, 0, 0)])], unique=True, key_min=1, key_max=n)
    probe = ctx.upload(INT64, np.arange(1, n + 1, dtype=np.int64))
    a = j.probe_bits(probe, 0, n, slice_bytes=1 << 30).to_numpy()
    b = j2.probe_bits(probe, 0, n, slice_bytes=1 << 30).to_numpy()
    assert np.array_equal(a, b)
    dup = keys.copy()
    dup[5] = dup[77]
    dc = ctx.upload(INT64, dup)
    bits, ins, flags = ctx.join_build_bitmap_nosync(dc, key_min=1, key_max=n).verdict()
    assert ins == n and bits == n - 1 and flags == 0
    far = keys.copy()
    far[9] = n + 1000
    fc = ctx.upload(INT64, far)
    bits, ins, flags = ctx.join_build_bitmap_nosync(fc, key_min=1, key_max=n).verdict()
    assert flags == 2 and ins == n - 1 == bits


@pytest.mark.parametrize("domain", [1, 31, 16384, 16385, 70_001, 1_048_576, 1_100_003])
def test_dense_group_by_emit_paths(bq, ctx, domain):
    """Dense states finish in one launch per 16384-slot chunk (k_finish_small; a counting launch first when there are several
    chunks) up to 2^20 slots, larger ones through presence bits, compaction and emit: every path must give the groups that
    exist, in key order, with COUNT / SUM / AVG as the reference defines them."""
    n = 200_003
    rng = np.random.default_rng(domain)
    present = np.sort(rng.choice(domain, size=max(1, (domain * 2) // 3), replace=False))
    k = (present[rng.integers(0, present.size, size=n)] + 1000).astype(np.int64)
    v = (rng.integers(-500, 500, size=n) / 8.0).astype(np.float64)
    kc, vc = ctx.upload(INT64, k), ctx.upload(DOUBLE, v)        # (kept alive: a slot holds the handle, not the Column)
    s = bq.ScanSpec()
    s.key = bq.make_slot(kc)
    s.a = bq.make_slot(vc)
    s.row_begin, s.row_end = 0, n
    s.n_v = 1
    s.v[0] = bq.VExpr(op=bq.V_A)
    s.group_mode = bq.GROUP_DENSE
    s.key_min, s.key_max = 1000, 1000 + domain - 1
    s.n_out = 3
    s.out[0] = bq.AggOut(func=bq.AGG_COUNT)
    s.out[1] = bq.AggOut(func=bq.AGG_SUM, v=0)
    s.out[2] = bq.AggOut(func=bq.AGG_AVG, v=0)
    got = ctx.scan_aggregate(s).to_numpy()
    keys, inv, counts = np.unique(k, return_inverse=True, return_counts=True)
    sums = np.bincount(inv, weights=v, minlength=keys.size)
    assert np.array_equal(got[0], keys) and np.array_equal(got[1], counts)          # key order, exact counts
    assert np.array_equal(got[2], sums)                                              # multiples of 1/8: sums are exact in any order
    assert np.array_equal(got[3], sums / counts)
    s.n_out = 1                                                                      # SUM only: presence rides on the sums (-0.0 marks)
    s.out[0] = bq.AggOut(func=bq.AGG_SUM, v=0)
    got = ctx.scan_aggregate(s).to_numpy()
    assert np.array_equal(got[0], keys) and np.array_equal(got[1], sums)


# ---- sort / limit --------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [0, 1, 2, 500, 4096, 4097, 300_001])
def test_sort_multi_key(bq, ctx, n):
    rng = np.random.default_rng(n + 1)
    a = rng.integers(-5, 5, size=n).astype(np.int64)
    b = (rng.integers(-1000, 1000, size=n) / 7.0).astype(np.float64)
    c = rng.integers(0, 2**32 - 1, size=n, dtype=np.uint64).astype(np.uint32)
    d = np.arange(n, dtype=np.int32)
    rel = ctx.rel_create([ctx.upload(INT64, a), ctx.upload(DOUBLE, b), ctx.upload(STRING, c), ctx.upload(DATE32, d)])
    got = ctx.rel_sort(rel, [0, 1], [0, 1]).to_numpy()       # a DESC, b ASC ; ties keep input order
    order = np.lexsort((d, b, -a))
    for g, w in zip(got, (a, b, c, d)):
        assert np.array_equal(g, w[order])
    got = ctx.rel_sort(rel, [2], [1], limit=10).to_numpy()
    order = np.argsort(c, kind="stable")[:10]
    for g, w in zip(got, (a, b, c, d)):
        assert np.array_equal(g, w[order])


def test_radix_sort_wide_keys(bq, ctx):
    """The LSD radix passes over keys that use all eight bytes: 64-bit integers (with a heavily repeated value: stability)
    and doubles of both signs, ascending and descending, over several hundred 4096-key tiles with a ragged last one."""
    n = 1_000_003
    rng = np.random.default_rng(5)
    k = rng.integers(-2**62, 2**62, size=n).astype(np.int64)
    k[::7] = k[3]
    f = rng.standard_normal(n) * 1e6
    f[::11] = -f[5]
    tag = np.arange(n, dtype=np.int32)
    rel = ctx.rel_create([ctx.upload(INT64, k), ctx.upload(DOUBLE, f), ctx.upload(DATE32, tag)])
    got = ctx.rel_sort(rel, [0], [1]).to_numpy()
    order = np.argsort(k, kind="stable")
    assert np.array_equal(got[0], k[order]) and np.array_equal(got[2], tag[order])
    got = ctx.rel_sort(rel, [1], [0]).to_numpy()                   # DESC: ties keep input order
    order = np.argsort(-f, kind="stable")
    assert np.array_equal(got[1], f[order]) and np.array_equal(got[2], tag[order])
    got = ctx.rel_sort(rel, [1, 0], [1, 0]).to_numpy()             # f ASC, then k DESC
    order = np.lexsort((tag, -k, f))
    assert np.array_equal(got[1], f[order]) and np.array_equal(got[0], k[order]) and np.array_equal(got[2], tag[order])


def test_full_sort_keeps_negative_zero_bits(bq, ctx):
    """A full sort writes its first sort column back from the sorted keys - except when the column holds -0.0, whose sign the
    key does not carry (-0.0 and +0.0 compare equal and must tie): then the column is gathered, and the bits survive."""
    n = 50_001
    rng = np.random.default_rng(9)
    f = rng.integers(-3, 4, size=n).astype(np.float64)
    f[::5] = -0.0
    tag = np.arange(n, dtype=np.int64)
    rel = ctx.rel_create([ctx.upload(DOUBLE, f), ctx.upload(INT64, tag)])
    for asc in (1, 0):
        got = ctx.rel_sort(rel, [0], [asc]).to_numpy()
        order = np.argsort(f if asc else -f, kind="stable")
        assert np.array_equal(got[0].view(np.int64), f[order].view(np.int64)) and np.array_equal(got[1], tag[order])
    g = np.where(f == 0.0, 0.0, f)                      # no -0.0 left: the reconstruction path, checked bit for bit as well
    rel = ctx.rel_create([ctx.upload(DOUBLE, g), ctx.upload(INT64, tag)])
    got = ctx.rel_sort(rel, [0], [0]).to_numpy()
    order = np.argsort(-g, kind="stable")
    assert np.array_equal(got[0].view(np.int64), g[order].view(np.int64)) and np.array_equal(got[1], tag[order])


# ---- expression programs -----------------------------------------------------------------------------
def test_eval_programs(bq, ctx):
    n = 10_007
    rng = np.random.default_rng(8)
    x = rng.integers(-1000, 1000, size=n).astype(np.int64)
    y = (rng.integers(-1000, 1000, size=n) / 4.0).astype(np.float64)
    dte = rng.integers(20200101, 20201231, size=n).astype(np.int32)
    xc, yc, dc = ctx.upload(INT64, x), ctx.upload(DOUBLE, y), ctx.upload(DATE32, dte)
    # (x * 2 + 1)  int arithmetic
    got = ctx.eval([("COL", 0, 0), ("IMM_I", 0, 2), ("MUL_I", 0, 0), ("IMM_I", 0, 1), ("ADD_I", 0, 0)], [xc], 0, n, INT64).to_numpy()
    assert np.array_equal(got, x * 2 + 1)
    # x * y -> double(x) * y
    got = ctx.eval([("COL", 0, 0), ("COL", 1, 0), ("I2F_2", 0, 0), ("MUL_F", 0, 0)], [xc, yc], 0, n, DOUBLE).to_numpy()
    assert np.array_equal(got, x.astype(np.float64) * y)
    # x < y : INT64-left compare truncates the double right operand (H6)
    got = ctx.eval([("COL", 0, 0), ("COL", 1, 0), ("F2I", 0, 0), ("LT_I", 0, 0)], [xc, yc], 0, n, INT64).to_numpy()
    assert np.array_equal(got, (x < np.trunc(y).astype(np.int64)).astype(np.int64))
    # y / 0 = +inf regardless of sign (H10)
    got = ctx.eval([("COL", 0, 0), ("IMM_F", 0, 0.0), ("DIV_F", 0, 0)], [yc], 0, n, DOUBLE).to_numpy()
    assert np.all(np.isposinf(got))
    # (x > 0 OR y < 0.0) AND date >= 20200601
    prog = [("COL", 0, 0), ("IMM_I", 0, 0), ("GT_I", 0, 0), ("COL", 1, 0), ("IMM_F", 0, 0.0), ("LT_F", 0, 0), ("OR", 0, 0),
            ("COL", 2, 0), ("IMM_I", 0, 20200601), ("SX32", 0, 0), ("GE_I", 0, 0), ("AND", 0, 0)]
    got = ctx.eval(prog, [xc, yc, dc], 0, n, INT64).to_numpy()
    assert np.array_equal(got, (((x > 0) | (y < 0.0)) & (dte >= 20200601)).astype(np.int64))
    # integer division by zero raises the reference's message (H10)
    with pytest.raises(bq.BqError, match="Division by zero"):
        ctx.eval([("COL", 0, 0), ("IMM_I", 0, 0), ("DIV_I", 0, 0)], [xc], 0, n, INT64)
    # malformed programs are rejected on the host
    with pytest.raises(bq.BqError):
        ctx.eval([("ADD_I", 0, 0)], [xc], 0, n, INT64)


def test_partial_merge(bq, ctx):
    """Two row-range partials merged == one scan (the multi-GPU merge path, run on one device)."""
    n = 120_001
    tab = datagen.host_table(datagen.orders_schema(n), n, seed=77)
    cols = {name: arr for name, _t, arr in tab}
    status, date, total = ctx.upload(STRING, cols["status"]), ctx.upload(DATE32, cols["order_date"]), ctx.upload(DOUBLE, cols["total"])
    full = q1_kernel_spec(bq, status, date, total, n, 0, 20240101, 20241231, 20240101, 20241228)
    want = ctx.scan_aggregate(full).to_numpy()
    parts = []
    for rb, re in [(0, 50_001), (50_001, n)]:
        sp = q1_kernel_spec(bq, status, date, total, n, 0, 20240101, 20241231, 20240101, 20241228)
        sp.row_begin, sp.row_end = rb, re
        parts.append(ctx.scan_aggregate(sp, partial=True))
    got = ctx.agg_finish(parts, True, DATE32, [bq.AggOut(func=bq.AGG_SUM, v=0)]).to_numpy()
    assert_same_rows(got, want, what="merged partials")


# ---- hash partitioning (the exchange step) and the partition-major group table ----------------------------------
def _key_hash_np(k):
    """key_hash of bq_common.cuh (murmur3 finaliser), vectorised."""
    x = k.astype(np.int64).view(np.uint64).copy()
    with np.errstate(over="ignore"):
        x ^= x >> np.uint64(33)
        x *= np.uint64(0xFF51AFD7ED558CCD)
        x ^= x >> np.uint64(33)
        x *= np.uint64(0xC4CEB9FE1A85EC53)
        x ^= x >> np.uint64(33)
    return x


@pytest.mark.parametrize("n,log2p", [(0, 3), (1, 2), (2047, 4), (2049, 8), (300_001, 8), (1_000_003, 10), (70_000, 1)])
def test_partition(bq, ctx, n, log2p):
    rng = np.random.default_rng(n + log2p)
    k = rng.integers(-10**12, 10**12, size=n).astype(np.int64)
    v = rng.normal(size=n)
    w = rng.integers(0, 2**32 - 1, size=n, dtype=np.uint64).astype(np.uint32)
    kc, vc, wc = ctx.upload(INT64, k), ctx.upload(DOUBLE, v), ctx.upload(STRING, w)
    ko, (vo, wo), off = ctx.partition(kc, [vc, wc], log2_parts=log2p)
    P = 1 << log2p
    offs = off.to_numpy()
    assert len(offs) == P + 1 and offs[0] == 0 and offs[-1] == n and np.all(np.diff(offs) >= 0)
    gk, gv, gw = ko.to_numpy(), vo.to_numpy(), wo.to_numpy()
    # the same rows, each in the partition its key hashes to
    a = np.lexsort((w, v, k))
    b = np.lexsort((gw, gv, gk))
    assert np.array_equal(k[a], gk[b]) and np.array_equal(v[a], gv[b]) and np.array_equal(w[a], gw[b])
    if n:
        for x in k[:5].tolist():
            assert int(_key_hash_np(np.array([x]))[0]) == bq.kernel_lib().bq_key_hash(x)
        part = (_key_hash_np(gk) >> np.uint64(64 - log2p)).astype(np.int64) & (P - 1)
        want = np.repeat(np.arange(P), np.diff(offs))
        assert np.array_equal(part, want)


def test_group_by_over_partitioned_rows(bq, ctx):
    """High-cardinality GROUP BY the L2-friendly way: partition by hash, then the partition-major hash table."""
    n, ids = 1_200_007, 150_000
    rng = np.random.default_rng(44)
    k = (rng.integers(0, ids, size=n) * 7919 - 10**9).astype(np.int64)
    v = rng.integers(1, 1000, size=n).astype(np.float64) / 8.0
    kc, vc = ctx.upload(INT64, k), ctx.upload(DOUBLE, v)
    log2p = 6
    ko, (vo,), _off = ctx.partition(kc, [vc], log2_parts=log2p)
    s = bq.ScanSpec()
    s.key = bq.make_slot(ko)
    s.a = bq.make_slot(vo)
    s.row_begin, s.row_end = 0, n
    s.n_v = 1
    s.v[0] = bq.VExpr(op=bq.V_A)
    s.group_mode = bq.GROUP_HASH
    s.ndv_hint = ids
    s.hash_part_log2, s.hash_part_shift = log2p, 64 - log2p
    s.n_out = 3
    s.out[0] = bq.AggOut(func=bq.AGG_COUNT)
    s.out[1] = bq.AggOut(func=bq.AGG_SUM, v=0)
    s.out[2] = bq.AggOut(func=bq.AGG_AVG, v=0)
    got = ctx.scan_aggregate(s).to_numpy()
    uk, inv = np.unique(k, return_inverse=True)
    cnt, sm = np.bincount(inv), np.bincount(inv, weights=v)
    o = np.argsort(got[0])
    assert np.array_equal(got[0][o], uk) and np.array_equal(got[1][o], cnt) and np.array_equal(got[2][o], sm)
    assert np.array_equal(got[3][o], sm / cnt)


# ---- shared-memory group tables over partitioned rows (bq_partition_aggregate, csrc/bq_groupby.cuh) -----------------
def _group_tables_case(bq, ctx, n, ids, log2p, splits, n_args, key_type=INT64, heavy=0.0, seed=0):
    """Rows -> partition() -> partition_aggregate(); numpy restatement of HashAggregate (src/exec/operator.cpp:984-1062).
    Values are dyadic, so every order of addition gives the same bits: the comparison is exact."""
    rng = np.random.default_rng(seed + n + ids)
    base = rng.integers(0, ids, size=n)
    if key_type == INT64:
        k = (base * 7919 - 1_000_000_007 * (base % 3)).astype(np.int64)
        if n > 2000:
            k[5::977] = np.iinfo(np.int64).min          # the key that equals the tables' empty marker
            k[7::1013] = np.iinfo(np.int64).max
        kd = k
    else:
        k = (20200101 + base).astype(np.int32)
        kd = k.astype(np.int64)
    if heavy:
        hot = rng.random(n) < heavy
        k[hot] = 42
        kd = k.astype(np.int64)
    a = rng.integers(-1000, 1001, size=n).astype(np.int64)
    b = rng.integers(-25600, 25601, size=n).astype(np.float64) / 8.0
    kc, ac, bc = ctx.upload(key_type, k), ctx.upload(INT64, a), ctx.upload(DOUBLE, b)
    pay = [bc] if n_args == 1 else ([ac, bc] if n_args == 2 else [])
    ko, po, off = ctx.partition(kc, pay, log2_parts=log2p)
    if n_args == 0:
        outs = [bq.AggOut(func=bq.AGG_COUNT)]
    elif n_args == 1:
        outs = [bq.AggOut(func=bq.AGG_COUNT), bq.AggOut(func=bq.AGG_SUM, v=0), bq.AggOut(func=bq.AGG_AVG, v=0)]
    else:
        outs = [bq.AggOut(func=bq.AGG_SUM, v=0, as_int=1), bq.AggOut(func=bq.AGG_COUNT), bq.AggOut(func=bq.AGG_SUM, v=1),
                bq.AggOut(func=bq.AGG_AVG, v=1)]
    got = ctx.partition_aggregate(ko, po, off, log2p, splits, outs).to_numpy()
    uk, inv = np.unique(kd, return_inverse=True)
    cnt = np.bincount(inv, minlength=len(uk))
    o = np.argsort(got[0].astype(np.int64), kind="stable")
    assert len(got[0]) == len(uk), f"{len(got[0])} groups, expected {len(uk)}"
    assert np.array_equal(got[0].astype(np.int64)[o], uk)
    if n_args == 0:
        assert np.array_equal(got[1][o], cnt)
    elif n_args == 1:
        sb = np.bincount(inv, weights=b, minlength=len(uk))
        assert np.array_equal(got[1][o], cnt) and np.array_equal(got[2][o], sb) and np.array_equal(got[3][o], sb / cnt)
    else:
        sa = np.bincount(inv, weights=a.astype(np.float64), minlength=len(uk))
        sb = np.bincount(inv, weights=b, minlength=len(uk))
        assert got[1].dtype == np.int64 and np.array_equal(got[1][o], sa.astype(np.int64))
        assert np.array_equal(got[2][o], cnt) and np.array_equal(got[3][o], sb) and np.array_equal(got[4][o], sb / cnt)


@pytest.mark.parametrize("n,ids,log2p,splits,n_args,key_type,heavy", [
    (600_011, 40_000, 4, 1, 1, INT64, 0.0),        # 2500 groups per 8192-slot table
    (600_011, 50_000, 2, 3, 1, INT64, 0.0),        # three splits per partition, load 0.5
    (400_003, 15_000, 3, 1, 2, INT64, 0.25),       # two sums (4096-slot tables), a quarter of the rows on ONE key
    (300_007, 9_000, 1, 2, 0, DATE32, 0.02),       # COUNT only, 4-byte keys
    (4_099, 700, 0, 1, 1, INT64, 0.0),             # one partition, ragged tail
    (37, 5, 4, 2, 1, INT64, 0.0),                  # mostly empty partitions
    (1, 1, 0, 1, 1, INT64, 0.0),
])
def test_partition_aggregate(bq, ctx, n, ids, log2p, splits, n_args, key_type, heavy):
    _group_tables_case(bq, ctx, n, ids, log2p, splits, n_args, key_type, heavy)


def test_partition_aggregate_sizing_rule_and_overflow(bq, ctx):
    # the sizing rule: expected load of every table <= 0.55, at most four splits, up to 1024 partitions
    assert ctx.group_tables_plan(12_500_000, 1) == (10, 3)          # configuration 4 on one GPU
    assert ctx.group_tables_plan(4_000_000, 1) == (10, 1)
    assert ctx.group_tables_plan(4_000_000, 2) == (10, 2)
    assert ctx.group_tables_plan(100_000_000, 1) is None            # would need 22 splits: the L2-resident table stays
    lp, sp = ctx.group_tables_plan(300_000, 1)
    _group_tables_case(bq, ctx, 1_200_007, 300_000, lp, sp, 1)
    # far more keys than the tables hold: a clean error (the operator layer falls back to bq_scan_aggregate), no wrong answer
    with pytest.raises(RuntimeError, match="table overflow"):
        _group_tables_case(bq, ctx, 300_000, 60_000, 1, 1, 1)
    # ... and the context is usable afterwards
    _group_tables_case(bq, ctx, 50_000, 3_000, 2, 1, 1)


def test_partition_aggregate_empty_input(bq, ctx):
    kc, vc = ctx.upload(INT64, np.zeros(0, np.int64)), ctx.upload(DOUBLE, np.zeros(0))
    ko, (vo,), off = ctx.partition(kc, [vc], log2_parts=3)
    got = ctx.partition_aggregate(ko, [vo], off, 3, 2, [bq.AggOut(func=bq.AGG_COUNT), bq.AggOut(func=bq.AGG_SUM, v=0)])
    assert got.rows == 0 and got.ncols == 3


@pytest.mark.gpu
@pytest.mark.parametrize("log2p", [1, 4, 9])
def test_partition_count_then_scatter_to_chosen_destinations(bq, ctx, log2p):
    """The two halves of the partition pass as the multi-GPU shuffle uses them: counts on the host first, then every
    partition is written where the caller says - here in REVERSE partition order inside one buffer (a peer's buffer in the
    shuffle).  Rows must arrive complete and each partition must be contiguous at its destination."""
    import ctypes as C
    K = bq.kernel_lib()
    n, P = 300_017, 1 << log2p
    rng = np.random.default_rng(log2p)
    k = rng.integers(-10**15, 10**15, size=n).astype(np.int64)
    v = rng.normal(size=n)
    d = rng.integers(0, 2**31 - 1, size=n).astype(np.int32)
    kc, vc, dc = ctx.upload(INT64, k), ctx.upload(DOUBLE, v), ctx.upload(DATE32, d)
    counts = (C.c_int64 * P)()
    plan = C.c_void_p()
    assert K.bq_partition_count(ctx.h, kc.h, 0, n, log2p, 40, counts, C.byref(plan)) == 0, K.bq_last_error()
    counts = np.array(list(counts), dtype=np.int64)
    part = ((_key_hash_np(k) >> np.uint64(40)).astype(np.int64)) & (P - 1)
    assert np.array_equal(counts, np.bincount(part, minlength=P))
    ko, vo, do = ctx.alloc(INT64, n), ctx.alloc(DOUBLE, n), ctx.alloc(DATE32, n)
    start = np.concatenate([[0], np.cumsum(counts[::-1])])[:-1][::-1]          # partition P-1 first, partition 0 last
    mk = lambda base, w: (C.c_void_p * P)(*[base + int(start[q]) * w for q in range(P)])
    pay = (C.c_void_p * 2)(vc.h, dc.h)
    assert K.bq_partition_scatter(ctx.h, plan, pay, 2, mk(ko.ptr, 8), mk(vo.ptr, 8), mk(do.ptr, 4)) == 0, K.bq_last_error()
    K.bq_part_plan_free(plan)
    gk, gv, gd = ko.to_numpy(), vo.to_numpy(), do.to_numpy()
    a, b = np.lexsort((d, v, k)), np.lexsort((gd, gv, gk))
    assert np.array_equal(k[a], gk[b]) and np.array_equal(v[a], gv[b]) and np.array_equal(d[a], gd[b])
    gpart = ((_key_hash_np(gk) >> np.uint64(40)).astype(np.int64)) & (P - 1)
    assert np.array_equal(gpart, np.repeat(np.arange(P)[::-1], counts[::-1]))


@pytest.mark.gpu
def test_ipc_export_needs_a_block_of_its_own(bq, ctx):
    import ctypes as C
    K = bq.kernel_lib()
    small = ctx.alloc(INT64, 1000)                       # from the stream-ordered pool: not exportable
    handle = (C.c_char * 64)()
    assert K.bq_col_ipc_export(ctx.h, small.h, handle) != 0 and b"shared" in K.bq_last_error()
    h = C.c_void_p()
    assert K.bq_col_alloc_shared(ctx.h, INT64, 1000, C.byref(h)) == 0
    assert K.bq_col_ipc_export(ctx.h, h, handle) == 0, K.bq_last_error()
    assert any(bytes(handle))                            # a real handle came back
    K.bq_col_free(ctx.h, h)
    r, u = C.c_size_t(), C.c_size_t()
    assert K.bq_ctx_pool_stats(ctx.h, C.byref(r), C.byref(u)) == 0 and r.value >= u.value
