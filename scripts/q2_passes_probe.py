"""Development probe: the Q2 probe side at kernel level - fused bitmap probe against key-range passes + row-bit scan, for an
L2-sized bitmap (250 M keys, 31 MB) and for the bitmap one rank of an 8-GPU weak-scaling run sees (2 B keys, 250 MB)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402

bq = load_package()
from bosql_b200 import synthetic as datagen  # noqa: E402

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000_000
ctx = bq.Context(0)
res = {}


def gen(schema, rows, seed):
    out = {}
    for i, (name, typ, spec) in enumerate(schema):
        out[name] = ctx.alloc(typ, rows).generate(seed=seed, stream=i, **spec)
    ctx.sync()
    return out


def kernel_ms(fn, reps=4):
    fn()
    ctx.profile_read()
    ctx.profile(True)
    for _ in range(reps):
        fn()
    ctx.profile(False)
    k, ms = ctx.profile_read()
    return ms / reps, k // reps


def scan_spec(li, no_keys=False):
    s = bq.ScanSpec()
    s.key = bq.make_slot(li["l.sku"])
    s.a = bq.make_slot(li["l.qty"])
    s.b = bq.make_slot(li["l.price"])
    s.row_begin, s.row_end = 0, n
    s.n_v = 1
    s.v[0] = bq.VExpr(op=bq.V_MUL)
    s.group_mode = bq.GROUP_DENSE
    s.key_min, s.key_max = 0, 99999
    s.n_out = 1
    s.out[0] = bq.AggOut(func=bq.AGG_SUM, v=0)
    return s


for n_orders in (n // 4, 2 * n):
    tag = f"{n_orders * 1e-6:.0f}M_keys"
    od = gen(datagen.orders_schema(n_orders, prefix="o.")[:2], n_orders, 2)
    li = gen(datagen.lineitem_schema(n_orders), n, 3)
    j = ctx.join_build(od["o.order_id"], preds=[bq.make_slot(od["o.status"], [(0, 0, 0)])], unique=True, key_min=1, key_max=n_orders)
    res[tag] = {"bitmap_mb": j.bytes / 1e6}
    s = scan_spec(li)
    s.jkey = bq.make_slot(li["l.order_id"])
    s.join = j.h
    ms, k = kernel_ms(lambda: ctx.scan_aggregate(s).free())
    res[tag]["fused_ms"] = ms
    print(tag, "fused probe", ms, flush=True)
    for hints in ("0", "1"):
        os.environ["BOSQL_PROBE_HINTS"] = hints
        for slice_mb in (32, 48, 64, 80, 96, 128):
            if (slice_mb << 20) * 1.25 >= j.bytes:
                continue
            ms_p, k = kernel_ms(lambda: j.probe_bits(li["l.order_id"], 0, n, slice_bytes=slice_mb << 20).free())
            res[tag][f"passes_slice{slice_mb}_hints{hints}_ms"] = ms_p
            print(tag, "key-range passes, slice", slice_mb, "MB, L2 hints", hints, ":", ms_p, "ms in", k, "launches", flush=True)
    os.environ["BOSQL_PROBE_HINTS"] = "0"
    bits = j.probe_bits(li["l.order_id"], 0, n, slice_bytes=96 << 20)
    s2 = scan_spec(li)
    s2.row_bits = bits.h
    ms, k = kernel_ms(lambda: ctx.scan_aggregate(s2).free())
    res[tag]["rowbits_scan_ms"] = ms
    print(tag, "row-bit scan:", ms, flush=True)
    del s, j, bits, od, li
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "q2_passes_probe.json"), "w"), indent=1)
