"""NVLink sanity: what do NCCL's collectives reach here for GB-sized buffers?  torchrun --nproc-per-node N scripts/nccl_probe.py"""
import os
import time

import torch
import torch.distributed as dist

world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
if rank == 0:
    print("peer access 0->1:", torch.cuda.can_device_access_peer(0, 1), flush=True)
n = 1 << 30
a = torch.empty(n, dtype=torch.uint8, device="cuda")
b = torch.empty(n, dtype=torch.uint8, device="cuda")
g = torch.empty(n * world, dtype=torch.uint8, device="cuda") if world <= 8 else None


def timeit(name, fn, bytes_moved):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 5
    if rank == 0:
        print(f"{name:28s} {dt * 1e3:8.2f} ms   {bytes_moved / dt / 1e9:8.1f} GB/s per GPU (sent)", flush=True)


timeit("all_to_all_single 1 GiB", lambda: dist.all_to_all_single(b, a), n * (world - 1) / world)
timeit("all_gather_into_tensor 1 GiB", lambda: dist.all_gather_into_tensor(g, a), n * (world - 1))
timeit("broadcast 1 GiB from 0", lambda: dist.broadcast(a, src=0), n)
timeit("all_reduce 1 GiB", lambda: dist.all_reduce(a.view(torch.int32)), 2 * n * (world - 1) / world)
if world >= 2:
    def sr():
        if rank == 0:
            dist.send(a, dst=1)
        elif rank == 1:
            dist.recv(b, src=0)
    timeit("send/recv 0->1 1 GiB", sr, n)
# ---- the shapes of the shuffle: 2 GB per rank, uneven splits, memory that torch did not allocate
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package  # noqa: E402
bq = load_package()
from bosql_b200 import distributed as D  # noqa: E402
m = 2_000_000_000
per = m // world // 8 * 8
sb = [per + (8 * 1000 * ((r + rank) % 3 - 1)) for r in range(world)]
sbt = torch.tensor(sb, dtype=torch.int64, device="cuda")
rbt = torch.empty_like(sbt)
dist.all_to_all_single(rbt, sbt)
rb = rbt.tolist()
a2 = torch.empty(sum(sb), dtype=torch.uint8, device="cuda")
b2 = torch.empty(sum(rb), dtype=torch.uint8, device="cuda")
timeit("a2a 2 GB uneven, torch memory", lambda: dist.all_to_all_single(b2, a2, output_split_sizes=rb, input_split_sizes=sb), sum(sb) - sb[rank])
ctx = bq.Context(local)
ca, cb = ctx.alloc(bq.INT64, sum(sb) // 8 + 1), ctx.alloc(bq.INT64, sum(rb) // 8 + 1)
ctx.sync()
ta = torch.as_tensor(D._CudaArray(ca.ptr, sum(sb), "|u1"), device="cuda")
tb = torch.as_tensor(D._CudaArray(cb.ptr, sum(rb), "|u1"), device="cuda")
assert ta.data_ptr() == ca.ptr and tb.data_ptr() == cb.ptr, "as_tensor copied"
timeit("a2a 2 GB uneven, bq memory", lambda: dist.all_to_all_single(tb, ta, output_split_sizes=rb, input_split_sizes=sb), sum(sb) - sb[rank])
st = torch.cuda.Stream()
def on_stream():
    with torch.cuda.stream(torch.cuda.ExternalStream(st.cuda_stream)):
        dist.all_to_all_single(tb, ta, output_split_sizes=rb, input_split_sizes=sb)
timeit("  ... on an external stream", on_stream, sum(sb) - sb[rank])
ex = D.Exchange(device="cuda")
import ctypes as C
i64 = lambda xs: (C.c_int64 * len(xs))(*xs)
def via_callback():
    ex.table.all_to_all_v(None, ca.ptr, i64(sb), cb.ptr, i64(rb), st.cuda_stream)
timeit("  ... through the C callback", via_callback, sum(sb) - sb[rank])


def isolated(name, fn):
    ts = []
    for _ in range(6):
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        ts.append((time.perf_counter() - t0) * 1e3)
    if rank == 0:
        print(f"{name:36s} isolated: " + " ".join(f"{t:6.2f}" for t in ts) + " ms", flush=True)


isolated("a2a 2 GB via callback", via_callback)
isolated("a2a 2 GB torch memory", lambda: dist.all_to_all_single(b2, a2, output_split_sizes=rb, input_split_sizes=sb))
# dirty the send buffer with a kernel first (as the partition pass does), then exchange
def dirty_then():
    ta.add_(1)
    ex.table.all_to_all_v(None, ca.ptr, i64(sb), cb.ptr, i64(rb), 0)
isolated("write send buffer, then a2a", dirty_then)
def host_then():
    t = torch.tensor([1, 2], dtype=torch.int64, device="cuda")
    g2 = torch.empty(2 * world, dtype=torch.int64, device="cuda")
    dist.all_gather_into_tensor(g2, t)
    g2.cpu()
    ex.table.all_to_all_v(None, ca.ptr, i64(sb), cb.ptr, i64(rb), st.cuda_stream)
isolated("host gather, then a2a", host_then)
dist.destroy_process_group()
