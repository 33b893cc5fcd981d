"""One invocation of the key-range probe passes (1 B probe keys over a 2 B-key, 250 MB bitmap, 64 MB slices) for ncu."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402

bq = load_package()
from bosql_b200 import synthetic as datagen  # noqa: E402

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000_000
n_orders = 2 * n
ctx = bq.Context(0)
od = {}
for i, (name, typ, spec) in enumerate(datagen.orders_schema(n_orders, prefix="o.")[:2]):
    od[name] = ctx.alloc(typ, n_orders).generate(seed=2, stream=i, **spec)
key = ctx.alloc(bq.INT64, n).generate(seed=3, stream=0, dist=bq.GEN_UNIFORM, lo=1, hi=n_orders)
ctx.sync()
j = ctx.join_build(od["o.order_id"], preds=[bq.make_slot(od["o.status"], [(0, 0, 0)])], unique=True, key_min=1, key_max=n_orders)
for _ in range(2):
    j.probe_bits(key, 0, n, slice_bytes=64 << 20).free()
ctx.sync()
print("done")
