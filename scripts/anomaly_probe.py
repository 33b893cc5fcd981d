"""Why does bq_partition take 100+ ms per call in a multi-process run when the same call takes 4 ms alone?  (scratch probe)"""
import ctypes as C
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402

world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
mode = sys.argv[1] if len(sys.argv) > 1 else "nccl"
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
bq = load_package()
ctx = bq.Context(local)
stream = torch.cuda.Stream()
ctx.set_stream(stream.cuda_stream)
K = bq.kernel_lib()
n = 250_000_000
k = ctx.alloc(bq.INT64, n).generate(dist=bq.GEN_HASHED, seed=5, stream=0, lo=0, hi=n // 20, modulus=1 << 61, row0=rank * n)
v = ctx.alloc(bq.DOUBLE, n).generate(dist=bq.GEN_UNIFORM_DIV, seed=5, stream=1, lo=1, hi=6400, div=64.0)
ctx.sync()


def pool():
    r, u = C.c_size_t(), C.c_size_t()
    K.bq_ctx_pool_stats(ctx.h, C.byref(r), C.byref(u))
    return f"pool reserved {r.value / 2**30:.2f} GiB used {u.value / 2**30:.2f} GiB"


def part(tag, shift, log2p):
    ctx.sync()
    t0 = time.perf_counter()
    pk, pp, off = ctx.partition(k, [v], log2_parts=log2p, hash_shift=shift)
    t1 = time.perf_counter()
    ctx.sync()
    t2 = time.perf_counter()
    if rank == 0:
        print(f"{tag:34s} host {1e3 * (t1 - t0):8.2f} ms  +sync {1e3 * (t2 - t1):8.2f} ms   {pool()}", flush=True)
    return pk, pp, off


for i in range(3):
    r = part(f"before any collective #{i}", 60, 4)
    del r
if world > 1:
    t = torch.ones(1 << 20, device="cuda")
    dist.all_reduce(t)
    torch.cuda.synchronize()
    for i in range(3):
        r = part(f"after all_reduce #{i}", 60, 4)
        del r
    with torch.cuda.stream(stream):
        a = torch.empty(1 << 28, dtype=torch.uint8, device="cuda")
        b = torch.empty(1 << 28, dtype=torch.uint8, device="cuda")
        dist.all_to_all_single(b, a)
    torch.cuda.synchronize()
    for i in range(3):
        r = part(f"after all_to_all #{i}", 40, 4)
        del r
    for i in range(3):
        with torch.cuda.stream(stream):
            dist.all_to_all_single(b, a)
        r = part(f"a2a in flight then partition #{i}", 40, 4)
        del r
    dist.destroy_process_group()
