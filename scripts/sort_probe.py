"""Development probe: one ORDER BY over n synthetic rows, left in HBM (for ncu captures of the radix kernels)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402

bq = load_package()
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
xl = bq.exec_lib()
assert xl.bqx_init(0) == 0
ctx = bq.wrap_context(xl.bqx_context())
k = ctx.alloc(bq.INT64, n).generate(dist=bq.GEN_UNIFORM, seed=7, stream=0, lo=0, hi=(1 << 40))
v = ctx.alloc(bq.DOUBLE, n).generate(dist=bq.GEN_UNIFORM_DIV, seed=7, stream=1, lo=1, hi=(1 << 52), div=4096.0)
ctx.sync()
eng = bq.Engine()
eng.add_table("s", [("k", bq.INT64, k), ("v", bq.DOUBLE, v)], eng.new_dict([]))
plan = eng.plan("SELECT k, v FROM s ORDER BY v DESC")
for _ in range(reps):
    plan.run_device().free()
ctx.sync()
print("done")
