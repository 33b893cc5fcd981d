"""GPU, BASELINE.json's full sizes (1 B rows): size-independent properties, since no CPU executor can check these
sizes row by row (the reference needs ~100 s and 250 B/row for them).

DOUBLE columns here are generated as k/64 (dyadic), so every partial sum is exactly representable: sums are bit-identical
in ANY order of addition, and different kernels / table kinds / partitionings must agree exactly — not within a tolerance.
Row count: $BOSQL_FULL_ROWS (default 1e9).
"""
import os

import numpy as np
import pytest

from oracle import datagen
from tests.parity import DATE32, DOUBLE, INT64, STRING, q1_kernel_spec

pytestmark = pytest.mark.gpu
N = int(float(os.environ.get("BOSQL_FULL_ROWS", "1e9")))
SEED = 99


def gen(ctx, schema, n, seed, row0=0):
    cols = {}
    for i, (name, typ, spec) in enumerate(schema):
        cols[name] = ctx.alloc(typ, n).generate(seed=seed, stream=i, row0=row0, **spec)
    ctx.sync()
    return cols


def agg(bq, ctx, n, a=None, b=None, preds=(), key=None, key_range=None, group=0, vexprs=(), outs=(), join=None, jkey=None,
        rows=None, ndv=0):
    s = bq.ScanSpec()
    if key is not None:
        s.key = bq.make_slot(key)
        s.group_mode = group
        if key_range:
            s.key_min, s.key_max = key_range
        s.ndv_hint = ndv
    if a is not None:
        s.a = bq.make_slot(a)
    if b is not None:
        s.b = bq.make_slot(b)
    for i, p in enumerate(preds):
        s.pred[i] = p
    if join is not None:
        s.join = join.h
        s.jkey = bq.make_slot(jkey)
    s.row_begin, s.row_end = rows if rows else (0, n)
    s.n_v = len(vexprs)
    for i, v in enumerate(vexprs):
        s.v[i] = v
    s.n_out = len(outs)
    for i, o in enumerate(outs):
        s.out[i] = o
    return ctx.scan_aggregate(s).to_numpy()


@pytest.fixture(scope="module")
def orders(ctx):
    return gen(ctx, datagen.orders_schema(N, div=64.0), N, SEED)


def test_q1_full_size_properties(bq, ctx, orders):
    o = orders
    spec = q1_kernel_spec(bq, o["status"], o["order_date"], o["total"], N, 0, 20240101, 20240131, 20240101, 20241228)
    spec.n_out = 2
    spec.out[1] = bq.AggOut(func=bq.AGG_COUNT)
    keys, sums, cnts = ctx.scan_aggregate(spec).to_numpy()
    # keys: exactly the 28 January days (1 B rows hit every day), ascending
    assert keys.tolist() == [20240100 + d for d in range(1, 29)]
    # the same predicate as a global aggregate: totals must agree exactly (dyadic values)
    tot = agg(bq, ctx, N, a=o["total"],
              preds=[bq.make_slot(o["status"], [(0, 0, 0)]), bq.make_slot(o["order_date"], [(20240101, 20240131, 0)])],
              vexprs=[bq.VExpr(op=bq.V_A)], outs=[bq.AggOut(func=bq.AGG_COUNT), bq.AggOut(func=bq.AGG_SUM, v=0)])
    assert int(cnts.sum()) == int(tot[0][0])
    assert float(sums.sum()) == float(tot[1][0])
    # expected selectivity 1/4 * 28/336 within 5 sigma
    p = 0.25 * 28 / 336
    assert abs(cnts.sum() - N * p) < 5 * np.sqrt(N * p)
    # partition invariance: three row ranges merged == one scan, bit for bit
    cuts = [0, N // 3 + 1, 2 * N // 3 + 7, N]
    parts = []
    for lo, hi in zip(cuts, cuts[1:]):
        sp = q1_kernel_spec(bq, o["status"], o["order_date"], o["total"], N, 0, 20240101, 20240131, 20240101, 20241228)
        sp.row_begin, sp.row_end = lo, hi
        parts.append(ctx.scan_aggregate(sp, partial=True))
    merged = ctx.agg_finish(parts, True, DATE32, [bq.AggOut(func=bq.AGG_SUM, v=0), bq.AggOut(func=bq.AGG_COUNT)])
    mk, ms, mc = ctx.rel_sort(merged, [0], [1]).to_numpy()
    assert np.array_equal(mk, keys) and np.array_equal(ms, sums) and np.array_equal(mc, cnts)
    # the hash table kind gives the same groups as the shared-memory kind
    sp = q1_kernel_spec(bq, o["status"], o["order_date"], o["total"], N, 0, 20240101, 20240131, 20240101, 20241228)
    sp.group_mode = bq.GROUP_HASH
    sp.ndv_hint = 400
    hk, hs = ctx.scan_aggregate(sp).to_numpy()
    order = np.argsort(hk)
    assert np.array_equal(hk[order], keys) and np.array_equal(hs[order], sums)


def test_filter_sweep_full_size(bq, ctx, orders):
    """COUNT(p) + COUNT(not p) = N and SUM(p) + SUM(not p) = SUM(all), exactly, at 1 % ... 99 % selectivity."""
    o = orders
    outs = [bq.AggOut(func=bq.AGG_COUNT), bq.AggOut(func=bq.AGG_SUM, v=0)]
    ve = [bq.VExpr(op=bq.V_A)]
    all_ = agg(bq, ctx, N, a=o["total"], vexprs=ve, outs=outs)
    assert int(all_[0][0]) == N
    lo_all, hi_all = -(1 << 63), (1 << 63) - 1
    for sel in (0.01, 0.1, 0.5, 0.9, 0.99):
        t = int(N * sel)
        yes = agg(bq, ctx, N, a=o["total"], preds=[bq.make_slot(o["order_id"], [(lo_all, t, 0)])], vexprs=ve, outs=outs)
        no = agg(bq, ctx, N, a=o["total"], preds=[bq.make_slot(o["order_id"], [(t + 1, hi_all, 0)])], vexprs=ve, outs=outs)
        assert int(yes[0][0]) == t and int(no[0][0]) == N - t          # order_id = row + 1
        assert float(yes[1][0]) + float(no[1][0]) == float(all_[1][0])
    # predicate on the DOUBLE column itself and on the date / status columns
    k = bq.f64_key(500.0)
    below = agg(bq, ctx, N, a=o["total"], preds=[], vexprs=ve, outs=outs, key=None)
    s = bq.ScanSpec()
    s.a = bq.make_slot(o["total"], [(bq.f64_key(-np.inf), k, 0)])
    s.row_begin, s.row_end, s.n_v, s.n_out = 0, N, 1, 2
    s.v[0], s.out[0], s.out[1] = ve[0], outs[0], outs[1]
    le = ctx.scan_aggregate(s).to_numpy()
    s.a = bq.make_slot(o["total"], [(k + 1, bq.f64_key(np.inf), 0)])
    gt = ctx.scan_aggregate(s).to_numpy()
    assert int(le[0][0]) + int(gt[0][0]) == N and float(le[1][0]) + float(gt[1][0]) == float(below[1][0])


def test_q2_full_size_properties(bq, ctx):
    n_orders, n_sku = max(4, N // 4), 100_000
    od = gen(ctx, datagen.orders_schema(n_orders, prefix="o.", div=64.0)[:2], n_orders, SEED + 1)
    li = gen(ctx, datagen.lineitem_schema(n_orders, n_sku, div=64.0), N, SEED + 2)
    status0 = [bq.make_slot(od["o.status"], [(0, 0, 0)])]
    j = ctx.join_build(od["o.order_id"], preds=status0, unique=True, key_min=1, key_max=n_orders)
    assert j.kind == bq.JOIN_BITMAP and j.bytes <= n_orders // 8 + 8
    mul = [bq.VExpr(op=bq.V_MUL)]
    sku, rev = agg(bq, ctx, N, key=li["l.sku"], key_range=(0, n_sku - 1), group=bq.GROUP_DENSE, a=li["l.qty"], b=li["l.price"],
                   vexprs=mul, outs=[bq.AggOut(func=bq.AGG_SUM, v=0)], join=j, jkey=li["l.order_id"])
    assert np.array_equal(sku, np.arange(n_sku))                   # every sku occurs, ascending
    # global aggregate over the same join: COUNT and SUM agree exactly with the grouped result
    cnt, tot = agg(bq, ctx, N, a=li["l.qty"], b=li["l.price"], vexprs=mul,
                   outs=[bq.AggOut(func=bq.AGG_COUNT), bq.AggOut(func=bq.AGG_SUM, v=0)], join=j, jkey=li["l.order_id"])
    assert float(rev.sum()) == float(tot[0])
    assert abs(int(cnt[0]) - N / 4) < 5 * np.sqrt(N * 0.25 * 0.75)
    # the hash-table join and the hash group table reproduce the bitmap / dense answer bit for bit (on a 1/8 slice to
    # keep the 12 B/slot table small)
    m = N // 8
    j2 = ctx.join_build(od["o.order_id"], preds=status0, kind=bq.JOIN_HASH, need_rows=True, unique=True, key_min=1, key_max=n_orders)
    a1 = agg(bq, ctx, N, key=li["l.sku"], key_range=(0, n_sku - 1), group=bq.GROUP_DENSE, a=li["l.qty"], b=li["l.price"],
             vexprs=mul, outs=[bq.AggOut(func=bq.AGG_SUM, v=0), bq.AggOut(func=bq.AGG_COUNT)], join=j, jkey=li["l.order_id"], rows=(0, m))
    a2 = agg(bq, ctx, N, key=li["l.sku"], group=bq.GROUP_HASH, ndv=n_sku, a=li["l.qty"], b=li["l.price"],
             vexprs=mul, outs=[bq.AggOut(func=bq.AGG_SUM, v=0), bq.AggOut(func=bq.AGG_COUNT)], join=j2, jkey=li["l.order_id"], rows=(0, m))
    o2 = np.argsort(a2[0])
    assert np.array_equal(a1[0], a2[0][o2]) and np.array_equal(a1[1], a2[1][o2]) and np.array_equal(a1[2], a2[2][o2])
    # top-20 by revenue == the 20 largest of the full group table
    rel = ctx.rel_create([ctx.upload(INT64, sku), ctx.upload(DOUBLE, rev)])
    top = ctx.rel_sort(rel, [1], [0], limit=20).to_numpy()
    order = np.argsort(-rev, kind="stable")[:20]
    assert np.array_equal(top[0], sku[order]) and np.array_equal(top[1], rev[order])


def test_high_cardinality_group_by(bq, ctx):
    """Configuration 4's shape on one GPU: GROUP BY a sparse INT64 key (hash table in HBM), SUM / COUNT / AVG."""
    n = max(1000, N // 4)
    ids = max(10, n // 20)                       # ~20 rows per key
    k = ctx.alloc(INT64, n).generate(dist=bq.GEN_HASHED, seed=SEED, stream=0, lo=0, hi=ids - 1, modulus=1 << 61)
    v = ctx.alloc(DOUBLE, n).generate(dist=bq.GEN_UNIFORM_DIV, seed=SEED, stream=1, lo=1, hi=6400, div=64.0)
    ctx.sync()
    keys, cnt, sm, avg = agg(bq, ctx, n, key=k, group=bq.GROUP_HASH, ndv=ids, a=v, vexprs=[bq.VExpr(op=bq.V_A)],
                             outs=[bq.AggOut(func=bq.AGG_COUNT), bq.AggOut(func=bq.AGG_SUM, v=0), bq.AggOut(func=bq.AGG_AVG, v=0)])
    assert len(keys) == len(np.unique(keys)) <= ids
    assert len(keys) > ids * 0.99                # 20 draws per id: practically every id occurs
    assert int(cnt.sum()) == n
    tot = agg(bq, ctx, n, a=v, vexprs=[bq.VExpr(op=bq.V_A)], outs=[bq.AggOut(func=bq.AGG_SUM, v=0)])
    assert float(sm.sum()) == float(tot[0][0])   # dyadic: exact
    assert np.array_equal(avg, sm / cnt)
    # a slice small enough for numpy: exact comparison per key
    m = min(n, 2_000_000)
    kk, vv = k.to_numpy(0, m), v.to_numpy(0, m)
    g = agg(bq, ctx, n, key=k, group=bq.GROUP_HASH, ndv=ids, a=v, vexprs=[bq.VExpr(op=bq.V_A)],
            outs=[bq.AggOut(func=bq.AGG_COUNT), bq.AggOut(func=bq.AGG_SUM, v=0)], rows=(0, m))
    uk, inv = np.unique(kk, return_inverse=True)
    o = np.argsort(g[0])
    assert np.array_equal(g[0][o], uk) and np.array_equal(g[1][o], np.bincount(inv)) and np.array_equal(g[2][o], np.bincount(inv, weights=vv))


@pytest.mark.parametrize("tables", ["0", "1"])
def test_sql_high_cardinality_group_by_takes_the_partitioned_path(bq, ctx, monkeypatch, tables):
    """Through the operator layer: catalog NDV says the table will not fit in L2, so the planner partitions first - then
    either the partition-major table in L2 (BOSQL_GROUP_TABLES=0) or one shared-memory table per partition (=1).
    The answer must equal the unpartitioned kernel's, bit for bit (dyadic values)."""
    monkeypatch.setenv("BOSQL_GROUP_TABLES", tables)
    n, ids = 8_000_003, 4_000_000
    k = ctx.alloc(INT64, n).generate(dist=bq.GEN_HASHED, seed=SEED, stream=0, lo=0, hi=ids - 1, modulus=1 << 61)
    v = ctx.alloc(DOUBLE, n).generate(dist=bq.GEN_UNIFORM_DIV, seed=SEED, stream=1, lo=1, hi=6400, div=64.0)
    ctx.sync()
    eng = bq.Engine()
    eng.add_table("t", [("k", INT64, k), ("v", DOUBLE, v)], stats={"k": (0, (1 << 61) - 1, ids)})
    before = bq.wrap_context(bq.exec_lib().bqx_context()).launches
    r = eng.query("SELECT k, COUNT(*), SUM(v), AVG(v) FROM t GROUP BY k")
    assert r.names == ["k", "COUNT(*)", "SUM(v)", "AVG(v)"]
    want = agg(bq, ctx, n, key=k, group=bq.GROUP_HASH, ndv=ids, a=v, vexprs=[bq.VExpr(op=bq.V_A)],
               outs=[bq.AggOut(func=bq.AGG_COUNT), bq.AggOut(func=bq.AGG_SUM, v=0), bq.AggOut(func=bq.AGG_AVG, v=0)])
    a, b = np.argsort(r.cols[0]), np.argsort(want[0])
    for g, w in zip(r.cols, want):
        assert np.array_equal(g[a], w[b])
    assert int(r.cols[1].sum()) == n
    # the shared-memory tables finish in ONE launch after the partition pass; the L2-resident table needs seven more
    launches = bq.wrap_context(bq.exec_lib().bqx_context()).launches - before
    assert (launches <= 10) == (tables == "1"), launches


@pytest.mark.parametrize("tables", ["0", "1"])
def test_sql_group_by_survives_a_stale_ndv(bq, ctx, monkeypatch, tables):
    """The catalog claims 3.2 M distinct keys; the rows hold ten million.  The shared-memory tables fill up, the L2-resident
    table sized from the claim fills up as well, and the third attempt sizes the table from the rows: slower, same answer."""
    monkeypatch.setenv("BOSQL_GROUP_TABLES", tables)
    n = 10_000_019
    k = ctx.alloc(INT64, n).generate(dist=bq.GEN_HASHED, seed=SEED + 9, stream=0, lo=0, hi=(1 << 40) - 1, modulus=1 << 61)
    v = ctx.alloc(DOUBLE, n).generate(dist=bq.GEN_UNIFORM_DIV, seed=SEED + 9, stream=1, lo=1, hi=6400, div=64.0)
    ctx.sync()
    eng = bq.Engine()
    eng.add_table("t", [("k", INT64, k), ("v", DOUBLE, v)], stats={"k": (0, (1 << 61) - 1, 3_200_000)})
    r = eng.query("SELECT k, COUNT(*), SUM(v) FROM t GROUP BY k")
    kk, vv = k.to_numpy(), v.to_numpy()
    uk, inv = np.unique(kk, return_inverse=True)
    assert len(uk) > 9_990_000
    o = np.argsort(r.cols[0])
    assert np.array_equal(r.cols[0][o], uk) and np.array_equal(r.cols[1][o], np.bincount(inv))
    assert np.array_equal(r.cols[2][o], np.bincount(inv, weights=vv))
