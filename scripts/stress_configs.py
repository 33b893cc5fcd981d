"""Configurations 4 and 5 of BASELINE.json through the SQL operator layer, on 1..N GPUs.

    python scripts/stress_configs.py                                   # one GPU
    torchrun --nproc-per-node N scripts/stress_configs.py [--rows R]   # N GPUs, R rows per GPU (weak scaling)

C4  SELECT k, SUM(v) FROM t GROUP BY k           k INT64 sparse (hashed), one distinct key per 20 rows; N > 1: key-hash
                                                 shuffle over NVLink, groups stay on the rank that owns them
C5  SELECT COUNT(*), SUM(p.v * b.w) FROM probe p JOIN build b ON p.k = b.k
                                                 b.k unique 1..B, p.k Zipf(1.1) over the build domain; N > 1: broadcast join
Each statement is timed with CUDA events on the context's stream (max over ranks) after warm-up; one JSON line per config is
printed by rank 0 and written to gpurun_out/stress_n<N>.json.  Results are checked through invariants that hold at any size
(row counts add up; C5's COUNT(*) equals the probe rows because every probe key exists in the build side).
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=float, default=2.5e8, help="rows per GPU (C4 table, C5 probe side)")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--only", default="")
    a = ap.parse_args()
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    bq = load_package()
    from bosql_b200 import synthetic as datagen
    xl = bq.exec_lib()
    if xl.bqx_init(local):
        raise RuntimeError(xl.bqx_last_error().decode())
    ctx = bq.wrap_context(xl.bqx_context())
    stream = torch.cuda.Stream()
    ctx.set_stream(stream.cuda_stream)
    ex = None
    if world > 1:
        from bosql_b200 import distributed as DIST
        ex = DIST.install(xl, device="cuda", keep_sharded=True)
    n = int(a.rows)
    out = []

    def timed(plan, steps):
        for _ in range(2):
            r = plan.run()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        l0 = ctx.launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(stream)
        for _ in range(steps):
            r = plan.run()
        e1.record(stream)
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) / steps * 1e3
        ms = e0.elapsed_time(e1) / steps
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        return r, ms, wall, (ctx.launches - l0) // steps

    def timed_device(plan, steps):
        """The same statement with its output left in HBM (bqx_plan_run_device): what a consumer on the GPU would see."""
        for _ in range(2):
            plan.run_device().free()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            plan.run_device().free()
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        return ms

    def total(x):
        if world == 1:
            return int(x)
        t = torch.tensor([int(x)], dtype=torch.int64, device="cuda")
        dist.all_reduce(t)
        return int(t)

    # ---- C4 -------------------------------------------------------------------------------------------------------
    if a.only in ("", "c4"):
        ids = max(16, n * world // 20)
        k = ctx.alloc(bq.INT64, n).generate(dist=bq.GEN_HASHED, seed=5, stream=0, lo=0, hi=ids - 1, modulus=1 << 61, row0=rank * n)
        v = ctx.alloc(bq.DOUBLE, n).generate(dist=bq.GEN_UNIFORM_DIV, seed=5, stream=1, lo=1, hi=6400, div=64.0, row0=rank * n)
        ctx.sync()
        eng = bq.Engine()
        eng.add_table("t", [("k", bq.INT64, k), ("v", bq.DOUBLE, v)], stats={"k": (0, (1 << 61) - 1, ids)})
        plan = eng.plan("SELECT k, SUM(v), COUNT(*) FROM t GROUP BY k")
        calls0 = dict(ex.calls) if ex else {}
        sent0 = ex.bytes_sent if ex else 0
        r, ms, wall, launches = timed(plan, a.steps)
        ms_dev = timed_device(plan, a.steps)
        groups = total(len(r.cols[0]))
        rows_seen = total(int(r.cols[2].sum()))
        assert rows_seen == n * world, (rows_seen, n * world)
        assert groups <= ids and groups > ids * 0.99, (groups, ids)
        o = {"config": "C4 GROUP BY high-cardinality key", "n_gpus": world, "rows_per_gpu": n, "distinct_keys": groups, "ms_per_step": ms,
             "ms_per_step_result_left_in_hbm": ms_dev, "rows_per_sec_result_left_in_hbm": n * world / (ms_dev * 1e-3),
             "result_bytes_to_host_per_gpu": int(sum(c.nbytes for c in r.cols)), "host_ms_per_step": wall, "rows_per_sec": n * world / (ms * 1e-3), "algorithmic_gbs": (16 * n * world + 24 * groups) / (ms * 1e-3) / 1e9,
             "launches_per_step": int(launches), "sql": "SELECT k, SUM(v), COUNT(*) FROM t GROUP BY k",
             "checked": "counts add up to the input rows; distinct keys within 1 % of the generator's domain"}
        if ex:
            steps_all = a.steps + 2
            # the peer-write shuffle moves its rows inside the partition kernel, not through a callback: count them here
            o["nvlink_bytes_stored_per_gpu_per_step"] = 16 * n * (world - 1) // world
            o["collective_bytes_sent_per_gpu_per_step"] = (ex.bytes_sent - sent0) // steps_all
            o["collectives_per_step"] = {c: (ex.calls[c] - calls0[c]) // steps_all for c in ex.calls}
        out.append(o)
        del plan, eng, r
        k.free()
        v.free()

    # ---- C5 -------------------------------------------------------------------------------------------------------
    if a.only in ("", "c5"):
        nb = max(16, n // 4)                       # build rows per GPU; the domain is 1 .. nb * world
        nb_all = nb * world
        bk = ctx.alloc(bq.INT64, nb).generate(dist=bq.GEN_SEQ, seed=6, stream=0, lo=1, row0=rank * nb)
        bw = ctx.alloc(bq.DOUBLE, nb).generate(dist=bq.GEN_UNIFORM_DIV, seed=6, stream=1, lo=1, hi=64, div=4.0, row0=rank * nb)
        cdf = datagen.zipf_cdf(min(nb_all, 1 << 22), 1.1)      # Zipf(1.1) over the hot head of the key domain
        pk = ctx.alloc(bq.INT64, n).generate(dist=bq.GEN_TABLE, seed=7, stream=0, lo=1, cdf=cdf, row0=rank * n)
        pv = ctx.alloc(bq.DOUBLE, n).generate(dist=bq.GEN_UNIFORM_DIV, seed=7, stream=1, lo=1, hi=64, div=4.0, row0=rank * n)
        ctx.sync()
        eng = bq.Engine()
        eng.add_table("build", [("b.k", bq.INT64, bk), ("b.w", bq.DOUBLE, bw)], stats={"b.k": (1, nb_all, nb_all)})
        eng.add_table("probe", [("p.k", bq.INT64, pk), ("p.v", bq.DOUBLE, pv)], stats={"p.k": (1, nb_all, min(nb_all, 1 << 22))})
        sql = "SELECT COUNT(*), SUM(p.v * b.w) FROM probe p JOIN build b ON p.k = b.k"
        plan = eng.plan(sql)
        sent0 = ex.bytes_sent if ex else 0
        r, ms, wall, launches = timed(plan, a.steps)
        assert int(r.cols[0][0]) == n * world, (int(r.cols[0][0]), n * world)
        o = {"config": "C5 Zipf(1.1) join + aggregate", "n_gpus": world, "probe_rows_per_gpu": n, "build_rows_per_gpu": nb, "ms_per_step": ms,
             "host_ms_per_step": wall, "rows_per_sec": (n + nb) * world / (ms * 1e-3), "algorithmic_gbs": 16 * (n + nb) * world / (ms * 1e-3) / 1e9,
             "launches_per_step": int(launches), "sql": sql, "join": "broadcast (build side all-gathered)" if world > 1 else "direct-address table",
             "checked": "COUNT(*) equals the probe rows (every probe key exists once in the build side)", "sum": float(r.cols[1][0])}
        if ex:
            o["nvlink_bytes_sent_per_gpu_per_step"] = (ex.bytes_sent - sent0) // (a.steps + 2)
        out.append(o)

    if rank == 0:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", f"stress_n{world}.json"), "w") as f:
            json.dump(out, f, indent=1)
        for o in out:
            os.write(real_stdout, (json.dumps(o) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
