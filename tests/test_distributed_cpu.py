"""CPU, world_size 2 and 3, gloo: the exchange table the C++ operator layer calls (bosql_b200.distributed.Exchange).

First test: the exchange step of a row-range-partitioned aggregate exactly as `gather_partials` (bo-sql_b200/host/
exchange.cpp) drives it - each rank's partial state [key, count, sum0, sum1] packed into one zero-padded block of fixed
capacity, ONE all_gather callback, per-rank views, merge in rank order (numpy restatements of the kernels on both ends) - must
equal the single-process aggregate of the whole table.  On GPUs the same callbacks run over NCCL with the kernels producing and
consuming the blocks (tests/test_distributed_gpu.py, bench.py --gpus N).  Second test: every callback of the table against
known answers, over host pointers."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import datagen


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _partial(cols, lo, hi):
    """[key, count, sum0, sum1] of Q1 over rows [lo,hi): what bq_scan_partial returns, restated in numpy."""
    st, d, t = cols["status"][lo:hi], cols["order_date"][lo:hi], cols["total"][lo:hi]
    m = (st == 0) & (d >= 20240101) & (d <= 20240331)
    keys, inv = np.unique(d[m], return_inverse=True)
    cnt = np.bincount(inv, minlength=len(keys)).astype(np.int64)
    s0 = np.bincount(inv, weights=t[m], minlength=len(keys))
    return keys.astype(np.int32), cnt, s0, np.zeros_like(s0)


def _merge(g_key, g_cnt, g_s0):
    """bq_agg_finish restated: equal keys combined, rows with count 0 are padding."""
    live = g_cnt != 0
    keys, inv = np.unique(g_key[live], return_inverse=True)
    cnt = np.zeros(len(keys), dtype=np.int64)
    s0 = np.zeros(len(keys))
    np.add.at(cnt, inv, g_cnt[live])
    np.add.at(s0, inv, g_s0[live])        # gathered in rank order -> added in rank order
    return keys, cnt, s0


def _align16(b):
    return (b + 15) // 16 * 16


def _worker(rank, world, port, n, q):
    import ctypes as C
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from __graft_entry__ import load_package
        load_package()
        from bosql_b200 import distributed as D
        ex = D.Exchange(device="cpu")
        tab = datagen.host_table(datagen.orders_schema(n), n, seed=11)
        cols = {name: a for name, _t, a in tab}
        lo, hi = rank * n // world, (rank + 1) * n // world
        part = _partial(cols, lo, hi)
        cap = 20241228 - 20240101 + 1
        # the block layout of gather_partials: columns back to back, each padded to 16 bytes, `cap` rows each, zero filled
        widths = [c.dtype.itemsize for c in part]
        offs = np.concatenate([[0], np.cumsum([_align16(w * cap) for w in widths])]).astype(int)
        block = int(offs[-1])
        send = np.zeros(block, dtype=np.uint8)
        for c, o in zip(part, offs):
            send[o:o + c.nbytes] = np.frombuffer(c.tobytes(), dtype=np.uint8)
        recv = np.zeros(block * world, dtype=np.uint8)
        assert ex.table.all_gather(None, send.ctypes.data_as(C.c_void_p), recv.ctypes.data_as(C.c_void_p), block, None) == 0, ex.error
        g = [np.concatenate([recv[r * block + o: r * block + o + w * cap].view(c.dtype) for r in range(world)])
             for c, o, w in zip(part, offs, widths)]
        # rank r's rows sit in [r*cap, r*cap + len(part_r)) in rank order, padding rows have count 0
        mine = slice(rank * cap, rank * cap + len(part[0]))
        assert np.array_equal(g[0][mine], part[0]) and np.array_equal(g[1][mine], part[1])
        keys, cnt, s0 = _merge(g[0], g[1], g[2])
        # bitmap union: disjoint bits from each rank, summed
        words = np.zeros(64, dtype=np.int32)
        words[rank::world] = 1 << rank
        assert ex.table.all_reduce_sum_u32(None, words.ctypes.data_as(C.c_void_p), 64, None) == 0, ex.error
        q.put((rank, keys, cnt, s0, words))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_partitioned_aggregate_world2():
    n, world = 40_003, 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=150) for _ in range(world)]
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    tab = datagen.host_table(datagen.orders_schema(n), n, seed=11)
    cols = {name: a for name, _t, a in tab}
    wk, wc, ws, _ = _partial(cols, 0, n)
    for rank, keys, cnt, s0, words in results:
        assert np.array_equal(keys, wk) and np.array_equal(cnt, wc)
        assert np.allclose(s0, ws, rtol=1e-12, atol=0)
        want = np.zeros(64, dtype=np.int32)
        for r in range(world):
            want[r::world] = 1 << r
        assert np.array_equal(words, want)


# ---- the operator layer's exchange table (bqx_exchange) over gloo with host pointers ---------------------------------
def _exchange_worker(rank, world, port, q):
    import ctypes as C
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from __graft_entry__ import load_package
        load_package()
        from bosql_b200 import distributed as D
        ex = D.Exchange(device="cpu")
        t = ex.table
        ptr = lambda a: a.ctypes.data_as(C.c_void_p)
        i64 = lambda xs: (C.c_int64 * len(xs))(*xs)

        # all_gather: fixed-size blocks in rank order
        mine = np.arange(16, dtype=np.uint8) + 16 * rank
        got = np.zeros(16 * world, dtype=np.uint8)
        assert t.all_gather(None, ptr(mine), ptr(got), 16, None) == 0, ex.error
        assert np.array_equal(got, np.arange(16 * world, dtype=np.uint8))

        # all_gather_v: rank r contributes 8*(r+1) bytes, one rank may contribute nothing
        sizes = [0 if r == 1 else 8 * (r + 1) for r in range(world)]
        mine = np.full(max(sizes[rank], 1), rank + 1, dtype=np.uint8)
        got = np.zeros(sum(sizes), dtype=np.uint8)
        assert t.all_gather_v(None, ptr(mine), ptr(got), i64(sizes), None) == 0, ex.error
        assert np.array_equal(got, np.concatenate([np.full(s, r + 1, dtype=np.uint8) for r, s in enumerate(sizes)]))

        # all_to_all_v: rank s sends (s + d + 1) int64 of value 100*s + d to rank d
        send_rows = [rank + d + 1 for d in range(world)]
        recv_rows = [s + rank + 1 for s in range(world)]
        send = np.concatenate([np.full(n, 100 * rank + d, dtype=np.int64) for d, n in enumerate(send_rows)])
        recv = np.zeros(sum(recv_rows), dtype=np.int64)
        assert t.all_to_all_v(None, ptr(send), i64([8 * n for n in send_rows]), ptr(recv), i64([8 * n for n in recv_rows]), None) == 0, ex.error
        assert np.array_equal(recv, np.concatenate([np.full(n, 100 * s + rank, dtype=np.int64) for s, n in enumerate(recv_rows)]))

        # all_reduce_sum_u32 on disjoint bits == bitwise OR, including the top bit
        words = np.zeros(64, dtype=np.uint32)
        words[rank::world] = 0x80000001
        words[63] = np.uint32(1) << np.uint32(31 - rank)
        assert t.all_reduce_sum_u32(None, ptr(words), 64, None) == 0, ex.error
        ref = np.zeros(64, dtype=np.uint64)
        for r in range(world):
            w = np.zeros(64, dtype=np.uint64)
            w[r::world] = 0x80000001
            w[63] = 1 << (31 - r)
            ref += w
        assert np.array_equal(words, (ref & 0xFFFFFFFF).astype(np.uint32))

        # host_all_gather_i64
        out = (C.c_int64 * (2 * world))()
        assert t.host_all_gather_i64(None, i64([rank, -rank]), 2, out) == 0, ex.error
        assert list(out) == [v for r in range(world) for v in (r, -r)]
        assert ex.calls["all_gather"] == 1 and ex.calls["all_to_all_v"] == 1
        q.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        import traceback
        q.put((rank, "FAIL " + traceback.format_exc()[-1500:]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_exchange_table_over_gloo(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_exchange_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] == "ok" for r in res), res


def test_set_exchange_validates():
    import ctypes as C
    from __graft_entry__ import load_package
    bq = load_package()
    from bosql_b200 import distributed as D
    L = bq.exec_lib()
    assert L.bqx_set_exchange(None) == 0                       # back to single-GPU execution
    t = D.ExchangeTable()                                      # world 0: treated as "no exchange"
    assert L.bqx_set_exchange(C.byref(t)) == 0
    t.world, t.rank = 2, 5
    assert L.bqx_set_exchange(C.byref(t)) != 0 and b"rank" in L.bqx_last_error()
    t.rank = 1                                                 # missing callbacks
    assert L.bqx_set_exchange(C.byref(t)) != 0 and b"collective" in L.bqx_last_error()
    assert L.bqx_set_exchange(None) == 0
