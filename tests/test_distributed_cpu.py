"""CPU, world_size 2, gloo: the exchange step of a row-range-partitioned aggregate (bosql_b200.distributed).

Each rank aggregates its own row range (numpy restatement of the fused kernel's partial state), the partial states are
all-gathered and merged in rank order; the result must equal the single-process aggregate of the whole table.  On GPUs the
same functions run over NCCL with the kernels producing / consuming the states (bench.py --gpus N)."""
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


def _worker(rank, world, port, n, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from __graft_entry__ import load_package
        bq = load_package()
        from bosql_b200 import distributed as D
        tab = datagen.host_table(datagen.orders_schema(n), n, seed=11)
        cols = {name: a for name, _t, a in tab}
        lo, hi = rank * n // world, (rank + 1) * n // world
        part = _partial(cols, lo, hi)
        cap = 20241228 - 20240101 + 1
        gathered = D.gather_partials([torch.from_numpy(np.ascontiguousarray(c)) for c in part], cap)
        g = [x.numpy() for x in gathered]
        assert all(len(x) == cap * world for x in g)
        # rank r's rows sit in [r*cap, r*cap + len(part_r)) in rank order
        mine = slice(rank * cap, rank * cap + len(part[0]))
        assert np.array_equal(g[0][mine], part[0]) and np.array_equal(g[1][mine], part[1])
        keys, cnt, s0 = _merge(g[0], g[1], g[2])
        # the packed variant (one collective) must deliver the same rows, as per-rank views
        _buf, views = D.gather_partials_packed([torch.from_numpy(np.ascontiguousarray(c)) for c in part], cap)
        for rk in range(world):
            for ci in range(4):
                assert np.array_equal(views[rk][ci].numpy(), g[ci][rk * cap:(rk + 1) * cap])
        # bitmap union: disjoint bits from each rank
        words = torch.zeros(64, dtype=torch.int32)
        words[rank::world] = 1 << rank
        D.or_reduce_bitmap(words)
        # the count exchange of the key-hash shuffle: rank r tells every peer how many rows it will send
        send = torch.tensor([10 * rank + p for p in range(world)], dtype=torch.int64)
        recv = D.exchange_counts(send)
        assert recv.tolist() == [10 * p + rank for p in range(world)]
        q.put((rank, keys, cnt, s0, words.numpy()))
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
