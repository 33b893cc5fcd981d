"""Configuration 4's pipeline across GPUs: GROUP BY a high-cardinality key with a key-hash shuffle over NVLink.

    torchrun --nproc-per-node N scripts/c4_shuffle.py [--rows R per GPU] [--ids distinct keys] [--check]

Every rank owns R rows of (k INT64 sparse, v DOUBLE).  Rows are hash-partitioned on the device into one run per rank
(bq_partition), exchanged with one NCCL all-to-all per column, partitioned once more locally so that each table region is
L2-sized, and aggregated (SUM / COUNT / AVG) into a table that stays sharded by key hash.  --check gathers everything to
rank 0 and compares with numpy (small sizes only).
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
    ap.add_argument("--rows", type=float, default=2.5e8)
    ap.add_argument("--ids", type=float, default=0)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--check", action="store_true")
    a = ap.parse_args()
    world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    os.dup2(2, 1) if False else None
    bq = load_package()
    from bosql_b200 import distributed as D
    ctx = bq.Context(local)
    stream = torch.cuda.Stream()
    ctx.set_stream(stream.cuda_stream)
    n = int(a.rows)
    ids = int(a.ids) if a.ids else max(16, n * world // 20)
    with torch.cuda.stream(stream):
        k = ctx.alloc(bq.INT64, n).generate(dist=bq.GEN_HASHED, seed=5, stream=0, lo=0, hi=ids - 1, modulus=1 << 61, row0=rank * n)
        v = ctx.alloc(bq.DOUBLE, n).generate(dist=bq.GEN_UNIFORM_DIV, seed=5, stream=1, lo=1, hi=6400, div=64.0, row0=rank * n)
        ctx.sync()
        log2w = world.bit_length() - 1
        LOG2P = 8

        phases = {}

        def mark(name, t0):
            torch.cuda.synchronize()
            phases[name] = phases.get(name, 0.0) + (time.perf_counter() - t0) * 1e3
            return time.perf_counter()

        def step():
            t0 = time.perf_counter()
            (rk, rv), recv = D.shuffle_by_key(ctx, k, [v])
            t0 = mark("shuffle (partition + all-to-all)", t0)
            m = rk.numel()
            kc, vc = ctx.wrap(bq.INT64, rk.data_ptr(), m), ctx.wrap(bq.DOUBLE, rv.data_ptr(), m)
            pk, (pv,), _off = ctx.partition(kc, [vc], log2_parts=LOG2P)
            t0 = mark("local partition", t0)
            s = bq.ScanSpec()
            s.key = bq.make_slot(pk)
            s.a = bq.make_slot(pv)
            s.row_begin, s.row_end = 0, m
            s.n_v = 1
            s.v[0] = bq.VExpr(op=bq.V_A)
            s.group_mode = bq.GROUP_HASH
            s.ndv_hint = max(1024, int(ids / world * 1.2))
            s.hash_part_log2, s.hash_part_shift = LOG2P, 64 - LOG2P
            s.n_out = 3
            s.out[0] = bq.AggOut(func=bq.AGG_COUNT)
            s.out[1] = bq.AggOut(func=bq.AGG_SUM, v=0)
            s.out[2] = bq.AggOut(func=bq.AGG_AVG, v=0)
            rel = ctx.scan_aggregate(s)
            mark("aggregate + emit", t0)
            return rel, m, (rk, rv, pk, pv)

        rel, m, keep = step()
        # every received key belongs to this rank
        gk = rel.col(0).to_numpy()
        mine = (np.array([bq.kernel_lib().bq_key_hash(int(x)) for x in gk[:2000]], dtype=np.uint64) >> np.uint64(D.SHUFFLE_SHIFT)) & np.uint64(world - 1)
        assert np.all(mine == rank), "a key landed on the wrong rank"
        cnt = rel.col(1).to_numpy()
        tot = torch.tensor([int(cnt.sum()), len(gk)], dtype=torch.int64, device="cuda")
        dist.all_reduce(tot)
        assert int(tot[0]) == n * world, (int(tot[0]), n * world)
        if a.check:
            ks, vs = k.to_numpy(), v.to_numpy()
            allk = [None] * world
            dist.all_gather_object(allk, (ks, vs, gk, cnt, rel.col(2).to_numpy(), rel.col(3).to_numpy()))
            if rank == 0:
                K = np.concatenate([x[0] for x in allk])
                V = np.concatenate([x[1] for x in allk])
                uk, inv = np.unique(K, return_inverse=True)
                gK = np.concatenate([x[2] for x in allk])
                o = np.argsort(gK)
                assert np.array_equal(gK[o], uk)
                assert np.array_equal(np.concatenate([x[3] for x in allk])[o], np.bincount(inv))
                assert np.array_equal(np.concatenate([x[4] for x in allk])[o], np.bincount(inv, weights=V))
                assert np.array_equal(np.concatenate([x[5] for x in allk])[o], np.bincount(inv, weights=V) / np.bincount(inv))
        del rel, keep
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        phases.clear()
        e0.record(stream)
        for _ in range(a.steps):
            r_, m_, keep_ = step()
            del r_, keep_
        e1.record(stream)
        torch.cuda.synchronize()
        dist.barrier()
        ms = torch.tensor([e0.elapsed_time(e1) / a.steps], dtype=torch.float64, device="cuda")
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        out = {"workload": "GROUP BY high-cardinality key, key-hash shuffle (configuration 4)", "n_gpus": world, "rows_per_gpu": n,
               "distinct_keys": int(tot[1]), "ms_per_step": float(ms), "rows_per_sec": n * world / (float(ms) * 1e-3),
               "nvlink_bytes_out_per_gpu": 16 * n * (world - 1) // world, "checked": bool(a.check),
               "phase_ms_rank0": {k_: v_ / a.steps for k_, v_ in phases.items()}, "shuffle_partition_ms_total": D._TIMES}
        print(json.dumps(out))
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", f"c4_shuffle_n{world}.json"), "w") as f:
            json.dump(out, f)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
