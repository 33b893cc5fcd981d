"""The three kernels behind configuration 4's numbers, once each, for one `ncu --set full` capture:
k_part_scatter at 32 and at 1024 partitions, and the L2-resident partition-major table (k_scan, G_HASH) over the 32-way output.
    ncu --set full --clock-control none -k regex:"k_part_scatter|k_scan" -c 3 -o gpurun_out/prof_r2_c4 python scripts/c4_ncu_probe.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402

bq = load_package()
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 250_000_000
ids = max(16, n // 20)
ctx = bq.Context(0)
k = ctx.alloc(bq.INT64, n).generate(dist=bq.GEN_HASHED, seed=46, stream=0, lo=0, hi=ids - 1, modulus=1 << 61)
v = ctx.alloc(bq.DOUBLE, n).generate(dist=bq.GEN_UNIFORM_DIV, seed=46, stream=1, lo=1, hi=6400, div=64.0)
ctx.sync()
ko, (vo,), off = ctx.partition(k, [v], log2_parts=5)
k10, (v10,), off10 = ctx.partition(k, [v], log2_parts=10)
ctx.sync()
del k10, v10, off10
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
s.out[0] = bq.AggOut(func=bq.AGG_SUM, v=0)
s.out[1] = bq.AggOut(func=bq.AGG_COUNT)
s.out[2] = bq.AggOut(func=bq.AGG_AVG, v=0)
r = ctx.scan_aggregate(s)
print("groups", r.rows)
