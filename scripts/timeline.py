"""Development probe: the device timeline of one query step (kernels, copies, gaps), through torch.profiler's CUPTI
activity records - they cover every launch in the process, the library's own kernels included.

    python scripts/timeline.py q1|q2|q2hash|c4|sort [rows_total] [--strong]          (also under torch.distributed.run)

Prints, for the last profiled step, every device activity in start order with its duration and the idle gap before it,
and the host-side runtime calls that synchronise.  Numbers taken under the profiler are for attribution only.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from __graft_entry__ import load_package  # noqa: E402

bq = load_package()
from bosql_b200 import synthetic as datagen  # noqa: E402

sys.path.insert(0, ROOT)
import bench  # noqa: E402  (SQL strings only)

which = sys.argv[1] if len(sys.argv) > 1 else "q2"
rows_total = int(float(sys.argv[2])) if len(sys.argv) > 2 and not sys.argv[2].startswith("-") else 1_000_000_000
strong = "--strong" in sys.argv
world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
torch.cuda.set_device(local)
xl = bq.exec_lib()
if xl.bqx_init(local):
    raise RuntimeError(xl.bqx_last_error().decode())
ctx = bq.wrap_context(xl.bqx_context())
stream = torch.cuda.Stream()
ctx.set_stream(stream.cuda_stream)
if world > 1:
    from bosql_b200 import distributed as DIST
    DIST.install_native(xl)
SEED = bench.SEED


def gen_table(schema, n, seed, r0):
    cols = {}
    for i, (name, typ, spec) in enumerate(schema):
        cols[name] = ctx.alloc(typ, n).generate(seed=seed, stream=i, row0=r0, **spec)
    ctx.sync()
    return cols


n_all = rows_total if strong else rows_total * world
n_loc = n_all // world
eng = bq.Engine()
d = eng.new_dict(datagen.STATUS_DICT)
if which == "q1":
    o = gen_table(datagen.orders_schema(n_all), n_loc, SEED, rank * n_loc)
    eng.add_table("orders", [("status", bq.STRING, o["status"]), ("order_date", bq.DATE32, o["order_date"]), ("total", bq.DOUBLE, o["total"])], d,
                  stats={"order_date": (20240101, 20241228, 336), "total": (1.0, 1000.0, 99901)})
    sql = bench.Q1_SQL
elif which in ("q2", "q2hash"):
    n_ord_all, n_ord = n_all // 4, n_loc // 4
    stride = 7919 if which == "q2hash" else 1           # sparse order ids: no bitmap, the open-addressing hash join
    o = gen_table(datagen.orders_schema(n_ord_all, prefix="o.", key_stride=stride)[:2], n_ord, SEED + 1, rank * n_ord)
    li = gen_table(datagen.lineitem_schema(n_ord_all, bench.N_SKU, key_stride=stride), n_loc, SEED + 2, rank * n_loc)
    eng.add_table("orders", [("o.order_id", bq.INT64, o["o.order_id"]), ("o.status", bq.STRING, o["o.status"])], d,
                  stats={"o.order_id": (1, 1 + (n_ord_all - 1) * stride, n_ord_all)})
    eng.add_table("lineitem", [(n, t, li[n]) for n, t, _ in datagen.lineitem_schema(n_ord_all, bench.N_SKU)], d,
                  stats={"l.sku": (0, bench.N_SKU - 1, bench.N_SKU), "l.order_id": (1, 1 + (n_ord_all - 1) * stride, n_ord_all)})
    sql = bench.Q2_SQL
elif which == "sort":
    k = ctx.alloc(bq.INT64, n_loc).generate(dist=bq.GEN_UNIFORM, seed=SEED + 7, stream=0, lo=0, hi=(1 << 40), row0=rank * n_loc)
    v = ctx.alloc(bq.DOUBLE, n_loc).generate(dist=bq.GEN_UNIFORM_DIV, seed=SEED + 7, stream=1, lo=1, hi=(1 << 52), div=4096.0, row0=rank * n_loc)
    ctx.sync()
    eng.add_table("s", [("k", bq.INT64, k), ("v", bq.DOUBLE, v)], d)
    sql = "SELECT k, v FROM s ORDER BY v DESC"
else:
    ids = max(16, n_all // 20)
    k = ctx.alloc(bq.INT64, n_loc).generate(dist=bq.GEN_HASHED, seed=SEED + 4, stream=0, lo=0, hi=ids - 1, modulus=1 << 61, row0=rank * n_loc)
    v = ctx.alloc(bq.DOUBLE, n_loc).generate(dist=bq.GEN_UNIFORM_DIV, seed=SEED + 4, stream=1, lo=1, hi=6400, div=64.0, row0=rank * n_loc)
    ctx.sync()
    eng.add_table("t", [("k", bq.INT64, k), ("v", bq.DOUBLE, v)], d, stats={"k": (0, (1 << 61) - 1, ids)})
    sql = "SELECT k, SUM(v), COUNT(*), AVG(v) FROM t GROUP BY k"
    xl.bqx_exchange_keep_sharded(1)

plan = eng.plan(sql)
if which in ("c4", "sort"):
    def run():
        plan.run_device().free()
else:
    run = plan.run
for _ in range(4):
    run()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(stream)
for _ in range(5):
    run()
e1.record(stream)
torch.cuda.synchronize()
plain_ms = e0.elapsed_time(e1) / 5

STEPS = 3
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(STEPS):
        run()
    torch.cuda.synchronize()

ev = prof.events()
dev = sorted([e for e in ev if e.device_type == torch.autograd.DeviceType.CUDA], key=lambda e: e.time_range.start)
host = sorted([e for e in ev if e.device_type == torch.autograd.DeviceType.CPU], key=lambda e: e.time_range.start)
if rank == 0:
    n_per = len(dev) // STEPS
    last = dev[-n_per:] if n_per else dev
    lines = [f"{which} rows_total={n_all} world={world} strong={strong}: {plain_ms:.3f} ms/step without the profiler; {len(dev)} device activities in {STEPS} steps"]
    t0 = last[0].time_range.start
    prev_end = t0
    busy = 0.0
    for e in last:
        s, t = e.time_range.start, e.time_range.end
        gap = s - prev_end
        lines.append(f"{(s - t0) / 1e3:9.3f} ms  +{(t - s):9.1f} us  gap {gap:8.1f} us  {e.name[:110]}")
        busy += (t - s)
        prev_end = max(prev_end, t)
    lines.append(f"span {(prev_end - t0) / 1e3:.3f} ms, busy {busy / 1e3:.3f} ms")
    # host calls of the last step
    hs = [e for e in host if e.time_range.start >= t0 - 200 and e.time_range.start <= prev_end]
    lines.append("host runtime calls in that window:")
    for e in hs:
        s, t = e.time_range.start, e.time_range.end
        if t - s >= 5 or "Synchronize" in e.name:
            lines.append(f"{(s - t0) / 1e3:9.3f} ms  +{(t - s):9.1f} us  {e.name[:80]}")
    text = "\n".join(lines)
    print(text)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    tag = f"{which}_n{world}{'_strong' if strong else ''}"
    open(os.path.join(ROOT, "gpurun_out", f"timeline_{tag}.txt"), "w").write(text + "\n")
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
