"""Configuration 4 on one GPU, both ways: the L2-resident partition-major table (bq_scan_aggregate) against the
shared-memory group tables (bq_partition_aggregate).  Kernel layer first (partition pass and aggregation timed apart, wall
clock around synchronised calls), then the whole statement through the operator layer under BOSQL_GROUP_TABLES=0 / 1.
usage: python scripts/group_tables_probe.py [rows] [out.json]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402

bq = load_package()
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 250_000_000
out_path = sys.argv[2] if len(sys.argv) > 2 else None
ids = max(16, n // 20)
ctx = bq.wrap_context(bq.exec_lib().bqx_context())
k = ctx.alloc(bq.INT64, n).generate(dist=bq.GEN_HASHED, seed=46, stream=0, lo=0, hi=ids - 1, modulus=1 << 61)
v = ctx.alloc(bq.DOUBLE, n).generate(dist=bq.GEN_UNIFORM_DIV, seed=46, stream=1, lo=1, hi=6400, div=64.0)
ctx.sync()


def timeit(fn, reps=4):
    fn()
    ts = []
    for _ in range(reps):
        ctx.sync()
        t0 = time.perf_counter()
        fn()
        ctx.sync()
        ts.append((time.perf_counter() - t0) * 1e3)
    return min(ts)


res = {"rows": n, "distinct_keys_nominal": ids}
outs = [bq.AggOut(func=bq.AGG_SUM, v=0), bq.AggOut(func=bq.AGG_COUNT), bq.AggOut(func=bq.AGG_AVG, v=0)]
plan = ctx.group_tables_plan(ids, 1)
res["plan_log2_parts_splits"] = plan

# ---- kernel layer: the partition pass at both widths
for log2p in sorted({5, plan[0]} if plan else {5}):
    def part():
        ko, po, off = ctx.partition(k, [v], log2_parts=log2p)
        return ko, po, off
    res[f"partition_log2p_{log2p}_ms"] = timeit(lambda: part())
    print(f"partition 2^{log2p}: {res[f'partition_log2p_{log2p}_ms']:.3f} ms", flush=True)

# ---- kernel layer: aggregation over partitioned rows
ko, (vo,), off = ctx.partition(k, [v], log2_parts=5)
s = bq.ScanSpec()
s.key = bq.make_slot(ko)
s.a = bq.make_slot(vo)
s.row_begin, s.row_end = 0, n
s.n_v = 1
s.v[0] = bq.VExpr(op=bq.V_A)
s.group_mode = bq.GROUP_HASH
s.ndv_hint = ids
s.hash_part_log2, s.hash_part_shift = 5, 64 - 5
s.n_out = 3
for i, o in enumerate(outs):
    s.out[i] = o
res["l2_table_aggregate_ms"] = timeit(lambda: ctx.scan_aggregate(s).free())
ref = ctx.scan_aggregate(s).to_numpy()
print(f"L2-resident table over 2^5 partitions: {res['l2_table_aggregate_ms']:.3f} ms, {len(ref[0])} groups", flush=True)
del ko, vo, off
if plan:
    for splits in sorted({plan[1], plan[1] + 1}):
        ko, (vo,), off = ctx.partition(k, [v], log2_parts=plan[0])
        try:
            ms = timeit(lambda: ctx.partition_aggregate(ko, [vo], off, plan[0], splits, outs).free())
            got = ctx.partition_aggregate(ko, [vo], off, plan[0], splits, outs).to_numpy()
            a, b = np.argsort(got[0]), np.argsort(ref[0])
            same = all(np.array_equal(g[a], w[b]) for g, w in zip(got, ref))        # dyadic values: exact in any order
            res[f"smem_tables_aggregate_splits_{splits}_ms"] = ms
            res[f"smem_tables_splits_{splits}_equal_to_l2_table"] = bool(same)
            print(f"shared-memory tables, 2^{plan[0]} partitions x {splits} splits: {ms:.3f} ms, {len(got[0])} groups, equal: {same}", flush=True)
        except Exception as e:  # noqa: BLE001
            res[f"smem_tables_splits_{splits}_error"] = str(e)
            print("shared-memory tables failed:", e, flush=True)
        del ko, vo, off

# ---- operator layer: the whole statement
eng = bq.Engine()
eng.add_table("t", [("k", bq.INT64, k), ("v", bq.DOUBLE, v)], stats={"k": (0, (1 << 61) - 1, ids)})
for mode in ("0", "1"):
    os.environ["BOSQL_GROUP_TABLES"] = mode
    p = eng.plan("SELECT k, SUM(v), COUNT(*), AVG(v) FROM t GROUP BY k")
    try:
        res[f"statement_ms_group_tables_{mode}"] = timeit(lambda: p.run_device().free(), reps=5)
        print(f"statement, BOSQL_GROUP_TABLES={mode}: {res[f'statement_ms_group_tables_{mode}']:.3f} ms", flush=True)
    except Exception as e:  # noqa: BLE001
        res[f"statement_group_tables_{mode}_error"] = str(e)
        print("statement failed:", e, flush=True)
print(json.dumps(res))
if out_path:
    with open(out_path, "w") as f:
        json.dump(res, f, indent=1)
