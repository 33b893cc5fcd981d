"""bosql_b200.distributed — the exchange step of partitioned aggregates (one process per GPU, torch.distributed).

Scans, selections and broadcast (bitmap) joins partition by row range: every rank runs the fused kernel on its own rows
and nothing crosses NVLink until the partial aggregate states (a few KB for Q1, 3 MB for Q2's 100 k groups) are
exchanged with ONE all-gather per state column and merged in rank order (so results do not depend on arrival order).
The same code runs over NCCL on GPUs and over gloo on CPU tensors (tests/test_distributed_cpu.py).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def gather_partials(cols, capacity: int, group=None):
    """cols: this rank's partial-state columns (1-D tensors of equal length r <= capacity; the count column zero means
    "no group").  Returns, per column, a tensor of world*capacity rows: rank 0's rows (zero-padded), then rank 1's, ..."""
    world = dist.get_world_size(group)
    out = []
    for c in cols:
        r = c.numel()
        if r > capacity:
            raise ValueError(f"partial state has {r} rows, capacity is {capacity}")
        pad = torch.zeros(capacity, dtype=c.dtype, device=c.device)
        if r:
            pad[:r].copy_(c)
        g = torch.empty(capacity * world, dtype=c.dtype, device=c.device)
        dist.all_gather_into_tensor(g, pad, group=group) if c.is_cuda else _gather_cpu(g, pad, world, group)
        out.append(g)
    return out


def _gather_cpu(g, pad, world, group):
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    torch.cat(parts, out=g)


def or_reduce_bitmap(words: torch.Tensor, group=None):
    """Union of per-rank join bitmaps built from disjoint build-side shards.  Bits set by different ranks never
    coincide (the BITMAP table requires unique keys), so the integer sum of the words IS their bitwise OR — which lets
    NCCL's all-reduce (no OR operator) do it in place."""
    dist.all_reduce(words, op=dist.ReduceOp.SUM, group=group)
    return words


def gather_partials_packed(cols, capacity: int, group=None):
    """Same exchange in ONE collective: the columns are packed (widest element first, so every column stays naturally
    aligned) into one byte buffer of capacity rows, all-gathered once, and handed back as per-rank views
    [rank][column] (no copies).  Padding rows are zero (count 0 = no group)."""
    world = dist.get_world_size(group)
    order = sorted(range(len(cols)), key=lambda i: -cols[i].element_size())
    offs, off = {}, 0
    for i in order:
        offs[i] = off
        off += cols[i].element_size() * capacity
    block = (off + 15) // 16 * 16
    dev = cols[0].device
    buf = torch.zeros(block, dtype=torch.uint8, device=dev)
    for i, c in enumerate(cols):
        r = c.numel()
        if r > capacity:
            raise ValueError(f"partial state has {r} rows, capacity is {capacity}")
        if r:
            buf[offs[i]:offs[i] + r * c.element_size()].copy_(c.contiguous().view(torch.uint8))
    g = torch.empty(block * world, dtype=torch.uint8, device=dev)
    if g.is_cuda:
        dist.all_gather_into_tensor(g, buf, group=group)
    else:
        _gather_cpu(g, buf, world, group)
    views = [[g[rk * block + offs[i]: rk * block + offs[i] + capacity * cols[i].element_size()].view(cols[i].dtype)
              for i in range(len(cols))] for rk in range(world)]
    return g, views


# ---- key-hash shuffle: partition on the device, exchange with one all-to-all per column -----------------------------
class _CudaArray:
    """__cuda_array_interface__ over a raw device pointer, so torch can view a bq column without copying."""

    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


_TYPESTR = {0: ("<i8", torch.int64), 1: ("<f8", torch.float64), 2: ("<i4", torch.int32), 3: ("<i4", torch.int32)}
_TIMES = {}
SHUFFLE_SHIFT = 40      # ranks are chosen by hash bits [40, 40 + log2(world)): disjoint from the L2-partition bits (top)
                        # and from the slot bits (bottom) used by the local tables


def as_tensor(col):
    """A torch view of a device-resident bq column (STRING / DATE32 as int32 bit patterns)."""
    ts, dt = _TYPESTR[col.type]
    n = col.n
    if n == 0:
        return torch.empty(0, dtype=dt, device="cuda")
    return torch.as_tensor(_CudaArray(col.ptr, n, ts), device="cuda")


def exchange_counts(send_counts, group=None):
    """send_counts[r] = rows this rank sends to rank r  ->  recv_counts[r] = rows rank r sends to this rank."""
    recv = torch.empty_like(send_counts)
    if send_counts.is_cuda:
        dist.all_to_all_single(recv, send_counts, group=group)
    else:                                   # gloo has no all_to_all_single on every build: gather the matrix instead
        world = dist.get_world_size(group)
        rows = [torch.empty_like(send_counts) for _ in range(world)]
        dist.all_gather(rows, send_counts, group=group)
        recv = torch.stack(rows)[:, dist.get_rank(group)].contiguous()
    return recv


def shuffle_by_key(ctx, key, payload, group=None):
    """Hash-partition (key, payload...) on the device into one run per rank (bq_partition), then exchange the runs with one
    NCCL all-to-all per column.  Returns the received columns as torch tensors (rows whose key hashes to this rank) and the
    per-peer receive counts.  world must be a power of two."""
    world = dist.get_world_size(group)
    log2w = world.bit_length() - 1
    if (1 << log2w) != world:
        raise ValueError("shuffle_by_key needs a power-of-two world size")
    import time as _t
    _t0 = _t.perf_counter()
    pk, pp, off = ctx.partition(key, payload, log2_parts=log2w, hash_shift=SHUFFLE_SHIFT)
    offs = torch.from_numpy(off.to_numpy()).to(torch.int64)
    _TIMES["partition"] = _TIMES.get("partition", 0.0) + (_t.perf_counter() - _t0) * 1e3
    send = (offs[1:] - offs[:-1]).cuda()
    recv = exchange_counts(send, group)
    send_l, recv_l = send.tolist(), recv.tolist()
    out = []
    for col in [pk] + list(pp):
        src = as_tensor(col)
        dst = torch.empty(sum(recv_l), dtype=src.dtype, device="cuda")
        dist.all_to_all_single(dst, src, output_split_sizes=recv_l, input_split_sizes=send_l, group=group)
        out.append(dst)
    torch.cuda.current_stream().synchronize()
    return out, recv_l
