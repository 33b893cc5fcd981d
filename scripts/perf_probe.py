"""Quick wall-clock probe of the fused kernels at full size (not the bench; used while developing)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402
from tests.parity import q1_kernel_spec  # noqa: E402

bq = load_package()
from bosql_b200 import synthetic as datagen  # noqa: E402
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000_000
ctx = bq.Context(0)
print("sm, free, total:", ctx.info())


def gen(schema, n, seed):
    out = {}
    for i, (name, typ, spec) in enumerate(schema):
        t0 = time.perf_counter()
        out[name] = ctx.alloc(typ, n).generate(seed=seed, stream=i, **spec)
        ctx.sync()
        print(f"  gen {name}: {time.perf_counter() - t0:.3f}s")
    return out


def timeit(fn, reps=5):
    fn()
    ts = []
    for _ in range(reps):
        ctx.sync()
        t0 = time.perf_counter()
        fn()
        ctx.sync()
        ts.append(time.perf_counter() - t0)
    return min(ts), sorted(ts)[len(ts) // 2]


res = {}
o = gen(datagen.orders_schema(n), n, 1)
spec = q1_kernel_spec(bq, o["status"], o["order_date"], o["total"], n, 0, 20240101, 20240131, 20240101, 20241228)
best, med = timeit(lambda: ctx.scan_aggregate(spec).free())
res["q1"] = dict(rows=n, best_ms=best * 1e3, med_ms=med * 1e3, gbs=16 * n / best / 1e9, grows=n / best / 1e9)
print("Q1", res["q1"])

# filter sweep on the same table: SUM(total) WHERE order_id < T
for sel in (0.01, 0.5, 0.99):
    s = bq.ScanSpec()
    s.a = bq.make_slot(o["total"])
    s.pred[0] = bq.make_slot(o["order_id"], [(-(1 << 63), int(n * sel), 0)])
    s.row_begin, s.row_end = 0, n
    s.n_v = 1
    s.v[0] = bq.VExpr(op=bq.V_A)
    s.n_out = 2
    s.out[0] = bq.AggOut(func=bq.AGG_COUNT)
    s.out[1] = bq.AggOut(func=bq.AGG_SUM, v=0)
    best, med = timeit(lambda: ctx.scan_aggregate(s).free())
    res[f"filter_i64_{sel}"] = dict(best_ms=best * 1e3, gbs=16 * n / best / 1e9)
    print("filter", sel, res[f"filter_i64_{sel}"])
del o

# Q2: lineitem(n) x orders(n/4)
no = n // 4
od = gen(datagen.orders_schema(no, prefix="o.")[:2], no, 2)
li = gen(datagen.lineitem_schema(no), n, 3)
def jb():
    return ctx.join_build(od["o.order_id"], preds=[bq.make_slot(od["o.status"], [(0, 0, 0)])], unique=True, key_min=1, key_max=no)
best, med = timeit(lambda: jb().free())
j = jb()
res["join_build"] = dict(best_ms=best * 1e3, med_ms=med * 1e3, gbs=12 * no / best / 1e9)
print("join build", res["join_build"], "kind", j.kind, "bytes", j.bytes)
s = bq.ScanSpec()
s.key = bq.make_slot(li["l.sku"])
s.a = bq.make_slot(li["l.qty"])
s.b = bq.make_slot(li["l.price"])
s.jkey = bq.make_slot(li["l.order_id"])
s.join = j.h
s.row_begin, s.row_end = 0, n
s.n_v = 1
s.v[0] = bq.VExpr(op=bq.V_MUL)
s.group_mode = bq.GROUP_DENSE
s.key_min, s.key_max = 0, 99999
s.n_out = 1
s.out[0] = bq.AggOut(func=bq.AGG_SUM, v=0)
best, med = timeit(lambda: ctx.scan_aggregate(s).free())
res["q2_probe"] = dict(best_ms=best * 1e3, gbs=32 * n / best / 1e9)
print("Q2 probe", res["q2_probe"])


def q2_all():
    jj = ctx.join_build(od["o.order_id"], preds=[bq.make_slot(od["o.status"], [(0, 0, 0)])], unique=True, key_min=1, key_max=no)
    s.join = jj.h
    rel = ctx.scan_aggregate(s)
    top = ctx.rel_sort(rel, [1], [0], limit=20)
    top.free(); rel.free(); jj.free()


best, med = timeit(q2_all)
res["q2"] = dict(best_ms=best * 1e3, gbs=(32 * n + 12 * no) / best / 1e9, mrows=(n + no) / best / 1e6)
print("Q2 all", res["q2"])
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "perf_probe.json"), "w"), indent=1)

# ---- configuration 4's shape on one GPU: GROUP BY a sparse INT64 key (100 M distinct), SUM / COUNT / AVG
del li, od, j
ngb = n
ids = max(10, n // 10)
k = ctx.alloc(bq.INT64, ngb).generate(dist=bq.GEN_HASHED, seed=5, stream=0, lo=0, hi=ids - 1, modulus=1 << 61)
v = ctx.alloc(bq.DOUBLE, ngb).generate(dist=bq.GEN_UNIFORM_DIV, seed=5, stream=1, lo=1, hi=6400, div=64.0)
ctx.sync()
s = bq.ScanSpec()
s.key = bq.make_slot(k)
s.a = bq.make_slot(v)
s.row_begin, s.row_end = 0, ngb
s.n_v = 1
s.v[0] = bq.VExpr(op=bq.V_A)
s.group_mode = bq.GROUP_HASH
s.ndv_hint = ids
s.n_out = 3
s.out[0] = bq.AggOut(func=bq.AGG_COUNT)
s.out[1] = bq.AggOut(func=bq.AGG_SUM, v=0)
s.out[2] = bq.AggOut(func=bq.AGG_AVG, v=0)
best, med = timeit(lambda: ctx.scan_aggregate(s).free(), reps=3)
res["groupby_100M"] = dict(rows=ngb, keys=ids, best_ms=best * 1e3, mrows=ngb / best / 1e6, gbs=(16 * ngb + 32 * ids) / best / 1e9)
print("GROUP BY high-cardinality", res["groupby_100M"])
del k, v

# ---- configuration 5's shape on one GPU: Zipf(1.1) probe keys against a unique build side, SUM(p.v * b.w)
nb = max(10, n // 4)
bk = ctx.alloc(bq.INT64, nb).generate(dist=bq.GEN_SEQ, seed=6, stream=0, lo=1)
bw = ctx.alloc(bq.DOUBLE, nb).generate(dist=bq.GEN_UNIFORM_DIV, seed=6, stream=1, lo=1, hi=64, div=4.0)
cdf = datagen.zipf_cdf(min(nb, 1 << 22), 1.1)     # hot head of the key domain (table lookup keeps it reproducible)
pk = ctx.alloc(bq.INT64, n).generate(dist=bq.GEN_TABLE, seed=7, stream=0, lo=1, cdf=cdf)
pv = ctx.alloc(bq.DOUBLE, n).generate(dist=bq.GEN_UNIFORM_DIV, seed=7, stream=1, lo=1, hi=64, div=4.0)
ctx.sync()
t0 = time.perf_counter()
jz = ctx.join_build(bk, need_rows=True, unique=True, key_min=1, key_max=nb)
ctx.sync()
print("zipf join build", time.perf_counter() - t0, "kind", jz.kind, "bytes", jz.bytes)
s = bq.ScanSpec()
s.a = bq.make_slot(pv)
s.b = bq.make_slot(bw, from_build=True)
s.jkey = bq.make_slot(pk)
s.join = jz.h
s.row_begin, s.row_end = 0, n
s.n_v = 1
s.v[0] = bq.VExpr(op=bq.V_MUL)
s.n_out = 2
s.out[0] = bq.AggOut(func=bq.AGG_COUNT)
s.out[1] = bq.AggOut(func=bq.AGG_SUM, v=0)
best, med = timeit(lambda: ctx.scan_aggregate(s).free(), reps=3)
res["zipf_join"] = dict(probe=n, build=nb, best_ms=best * 1e3, gbs=(16 * n + 16 * nb) / best / 1e9)
print("Zipf join probe", res["zipf_join"])
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "perf_probe.json"), "w"), indent=1)

# ---- the same high-cardinality GROUP BY through the partitioner (L2-resident table regions)
k = ctx.alloc(bq.INT64, ngb).generate(dist=bq.GEN_HASHED, seed=5, stream=0, lo=0, hi=ids - 1, modulus=1 << 61)
v = ctx.alloc(bq.DOUBLE, ngb).generate(dist=bq.GEN_UNIFORM_DIV, seed=5, stream=1, lo=1, hi=6400, div=64.0)
ctx.sync()
for log2p in (6, 8, 10):
    def part_only():
        ko, (vo,), off = ctx.partition(k, [v], log2_parts=log2p)
        return ko, vo, off
    best_p, _ = timeit(lambda: part_only(), reps=3)

    def gb_part():
        ko, vo, off = part_only()
        s2 = bq.ScanSpec()
        s2.key = bq.make_slot(ko)
        s2.a = bq.make_slot(vo)
        s2.row_begin, s2.row_end = 0, ngb
        s2.n_v = 1
        s2.v[0] = bq.VExpr(op=bq.V_A)
        s2.group_mode = bq.GROUP_HASH
        s2.ndv_hint = ids
        s2.hash_part_log2, s2.hash_part_shift = log2p, 64 - log2p
        s2.n_out = 3
        s2.out[0] = bq.AggOut(func=bq.AGG_COUNT)
        s2.out[1] = bq.AggOut(func=bq.AGG_SUM, v=0)
        s2.out[2] = bq.AggOut(func=bq.AGG_AVG, v=0)
        ctx.scan_aggregate(s2).free()
    best, med = timeit(gb_part, reps=3)
    res[f"groupby_100M_part{log2p}"] = dict(partition_ms=best_p * 1e3, total_ms=best * 1e3, mrows=ngb / best / 1e6)
    print("GROUP BY high-cardinality, partitioned", log2p, res[f"groupby_100M_part{log2p}"])
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "perf_probe.json"), "w"), indent=1)
