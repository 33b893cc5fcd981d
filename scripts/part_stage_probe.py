"""k_part_scatter with and without the cp.async staging of the next tile (BOSQL_PART_STAGE=0|1): 250 M (int64, double) rows
into 32 partitions (configuration 4's local partition) and into 8 (a shuffle's fan-out), then the whole C4 statement."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402

bq = load_package()
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 250_000_000
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


res = {"rows": n}
eng = bq.Engine()
eng.add_table("t", [("k", bq.INT64, k), ("v", bq.DOUBLE, v)], stats={"k": (0, (1 << 61) - 1, ids)})
plan = eng.plan("SELECT k, SUM(v), COUNT(*), AVG(v) FROM t GROUP BY k")
for mode in ("0", "1", "0", "1"):
    os.environ["BOSQL_PART_STAGE"] = mode
    for log2p in (5, 3):
        ms = timeit(lambda: ctx.partition(k, [v], log2_parts=log2p))
        res.setdefault(f"partition_2^{log2p}_stage_{mode}_ms", []).append(round(ms, 3))
    ms = timeit(lambda: plan.run_device().free())
    res.setdefault(f"c4_statement_stage_{mode}_ms", []).append(round(ms, 3))
print(json.dumps(res, indent=1))
