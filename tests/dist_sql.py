"""Worker for the multi-GPU operator-layer tests: every rank loads a ROW SHARD of the same synthetic tables, installs the
exchange table (bosql_b200.distributed.install) and runs the same SQL; results must equal the oracle's on the whole table.

Launched two ways:
  * tests/test_distributed_gpu.py spawns `world` processes that SHARE cuda:0 over a gloo group (device buffers staged
    through the host) - this is what a single-GPU box can run;
  * `python -m torch.distributed.run --nproc-per-node N tests/dist_sql.py` - one GPU per rank over NCCL.
The checker is the numpy restatement of the reference (oracle/oracle.py); nothing here reads /root/reference.
"""
import os
import sys
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_ORDERS = 60000
N_LINES = 200000
N_SKU = 500
BIG_ROWS = 6_000_000          # high-cardinality GROUP BY: ~4.7 M distinct keys -> the shuffle path (> 96 MB of state)
BIG_KEYS = 5_000_000


def shard(n, rank, world, skew):
    """Contiguous row ranges; skew=True gives rank 0 most rows and the LAST rank none (empty-shard coverage)."""
    if not skew:
        return rank * n // world, (rank + 1) * n // world
    cuts = [0] + [int(n * (0.6 + 0.4 * r / max(1, world - 1))) for r in range(world - 1)] + [n]
    cuts[-2] = n                      # last rank: empty
    return cuts[rank], cuts[rank + 1]


def tables():
    from oracle import datagen
    orders = datagen.host_table(datagen.orders_schema(N_ORDERS, prefix="o."), N_ORDERS, seed=21)
    lines = datagen.host_table(datagen.lineitem_schema(N_ORDERS, N_SKU), N_LINES, seed=22)
    return orders, lines


QUERIES = [
    # (name, sql, ordered_by, needs) -- needs: which catalog statistics are supplied ("stats" / "nostats" / "stale")
    ("q1_dense", "SELECT o.order_date, SUM(o.total) AS revenue FROM orders o WHERE o.status = 'COMPLETE' AND o.order_date >= 20240101 "
                 "AND o.order_date <= 20240331 GROUP BY o.order_date ORDER BY o.order_date", [(0, True)]),
    ("global", "SELECT COUNT(*), SUM(o.total), AVG(o.total) FROM orders o WHERE o.total > 250", None),
    ("global_empty", "SELECT COUNT(*), SUM(o.total) FROM orders o WHERE o.total > 99999999", None),
    ("two_keys", "SELECT o.status, o.order_date, COUNT(*) AS n FROM orders o WHERE o.order_date <= 20240215 GROUP BY o.status, o.order_date", None),
    ("avg_int", "SELECT l.sku, AVG(l.qty), SUM(l.qty), COUNT(*) FROM lineitem l GROUP BY l.sku ORDER BY l.sku", [(0, True)]),
    ("three_sums", "SELECT l.sku, SUM(l.qty), SUM(l.price), SUM(l.qty * l.price), COUNT(*) FROM lineitem l WHERE l.qty > 10 GROUP BY l.sku", None),
    ("q2_bitmap", "SELECT l.sku, SUM(l.qty * l.price) AS rev FROM lineitem l JOIN orders o ON l.order_id = o.order_id "
                  "WHERE o.status = 'COMPLETE' GROUP BY l.sku ORDER BY rev DESC LIMIT 20", [(1, False)]),
    ("join_payload", "SELECT o.status, COUNT(*) AS n, SUM(l.qty) AS q FROM lineitem l JOIN orders o ON l.order_id = o.order_id "
                     "GROUP BY o.status ORDER BY o.status", [(0, True)]),
    ("join_rows", "SELECT l.order_id, l.sku, o.status FROM lineitem l JOIN orders o ON l.order_id = o.order_id WHERE l.qty > 49", "sharded"),
    ("filter_rows", "SELECT o.order_id, o.total FROM orders o WHERE o.total > 990", "sharded"),
    # ORDER BY / LIMIT over sharded rows: local top-k, one all-gather, final top-k - the same complete answer on every rank
    ("topk_rows", "SELECT o.order_id, o.total FROM orders o WHERE o.status = 'PENDING' ORDER BY o.order_id DESC LIMIT 37", [(0, False)]),
    ("sort_rows", "SELECT l.order_id, l.qty FROM lineitem l WHERE l.qty > 49 AND l.sku < 40 ORDER BY l.order_id", [(0, True)]),
    ("limit_rows", "SELECT o.order_id FROM orders o WHERE o.total > 500 LIMIT 25", [(0, True)]),
    ("topk_join", "SELECT l.order_id, l.sku, o.status FROM lineitem l JOIN orders o ON l.order_id = o.order_id WHERE l.qty > 49 "
                  "ORDER BY l.order_id DESC, l.sku DESC LIMIT 11", [(0, False), (1, False)]),
]


class _Log(list):
    """results list that also traces every outcome to $BOSQL_DIST_LOG.<rank> (a hung collective leaves no other clue)."""

    def __init__(self, rank):
        super().__init__()
        base = os.environ.get("BOSQL_DIST_LOG")
        self.f = open(f"{base}.{rank}", "a") if base else None

    def append(self, item):
        super().append(item)
        if self.f:
            self.f.write(f"{item[0]}: {item[1][-500:]}\n")
            self.f.flush()


def run_rank(rank, world, backend, results):
    """Returns nothing; appends (name, 'ok' | message) to results."""
    import torch
    import torch.distributed as dist
    from __graft_entry__ import load_package
    from oracle import datagen, oracle as ORA
    from tests.parity import assert_same_rows

    bq = load_package()
    from bosql_b200 import distributed as D
    local = int(os.environ.get("LOCAL_RANK", "0")) if backend == "nccl" else 0
    torch.cuda.set_device(local)
    xl = bq.exec_lib()
    if xl.bqx_init(local):
        raise RuntimeError(xl.bqx_last_error().decode())
    # BOSQL_TEST_EXCHANGE=native: the library's own NCCL collectives (bqx_comm_init; one GPU per rank) instead of the
    # torch.distributed callback table
    native = os.environ.get("BOSQL_TEST_EXCHANGE") == "native"
    ex = D.install_native(xl) if native else D.install(xl, device="cuda")

    def reinstall(keep_sharded=False):
        if native:
            ex.keep_sharded(keep_sharded)
            return ex
        return D.install(xl, device="cuda", keep_sharded=keep_sharded)
    orders, lines = tables()
    ora = ORA.Oracle()
    sdict = ora.new_dict(datagen.STATUS_DICT)              # one dictionary shared by both tables, as in the engine
    ora.add_table("orders", orders, sdict)
    ora.add_table("lineitem", lines, sdict)

    def engine(skew, stats):
        eng = bq.Engine()
        d = eng.new_dict(datagen.STATUS_DICT)
        lo, hi = shard(N_ORDERS, rank, world, skew)
        llo, lhi = shard(N_LINES, rank, world, skew)
        ost = lst = None
        if stats == "stats":        # statistics describe the WHOLE table
            ost = {"o.order_id": (1, N_ORDERS, N_ORDERS), "o.order_date": (20240101, 20241228, 336)}
            lst = {"l.sku": (0, N_SKU - 1, N_SKU), "l.order_id": (1, N_ORDERS, N_ORDERS)}
        elif stats == "stale":      # bounds that miss part of the data: the kernels must notice, every rank must retry
            ost = {"o.order_id": (1, N_ORDERS, N_ORDERS), "o.order_date": (20240101, 20240120, 20)}
            lst = {"l.sku": (0, 99, 100), "l.order_id": (1, N_ORDERS, N_ORDERS)}
        eng.add_table("orders", [(n, t, np.ascontiguousarray(a[lo:hi])) for n, t, a in orders], d, stats=ost)
        eng.add_table("lineitem", [(n, t, np.ascontiguousarray(a[llo:lhi])) for n, t, a in lines], d, stats=lst)
        return eng

    def gather_rows(cols):
        """Sharded results: concatenate every rank's rows (host side, test only)."""
        objs = [None] * world
        dist.all_gather_object(objs, [c.copy() for c in cols])
        return [np.concatenate([o[i] for o in objs]) for i in range(len(cols))]

    for skew in (False, True):
        for stats in ("stats", "nostats", "stale"):
            eng = engine(skew, stats)
            for name, sql, order in QUERIES:
                tag = f"{name}[{'skew' if skew else 'even'},{stats}]"
                try:
                    got = eng.query(sql)
                    want = ora.query(sql)
                    cols = got.cols
                    if order == "sharded":
                        cols = gather_rows(cols)
                        order = None
                    assert_same_rows(cols, want.cols, ordered_by=order, what=tag)
                    results.append((tag, "ok"))
                except Exception:  # noqa: BLE001
                    results.append((tag, traceback.format_exc()[-1200:]))
            del eng

    # ---- errors must surface on EVERY rank (a rank that threw alone would leave the others inside a collective) ------
    eng = engine(False, "stats")
    try:
        eng.query("SELECT SUM(l.qty / (l.order_id - 7)) FROM lineitem l")      # one row of one shard divides by zero
        results.append(("div_by_zero", "no error raised"))
    except Exception as e:  # noqa: BLE001
        results.append(("div_by_zero", "ok" if "Division by zero" in str(e) else f"wrong error: {e}"))
    try:
        got = eng.query("SELECT COUNT(*) FROM orders o")                           # and the engine is still usable
        results.append(("after_error", "ok" if int(got.cols[0][0]) == N_ORDERS else f"count {got.cols[0]}"))
    except Exception:  # noqa: BLE001
        results.append(("after_error", traceback.format_exc()[-800:]))
    del eng

    # ---- co-partitioned join with skewed probe keys: both sides shuffled by hash(join key); the heavy hitters' probe rows stay
    # where they are and their build rows are replicated (forced here: at this size the planner would broadcast) ---------------
    try:
        nb, npr = 20000, 400000
        rows = np.arange(npr, dtype=np.uint64)
        u = datagen.row_hash(9, 0, rows)
        pk = (datagen.row_hash(9, 1, rows) % np.uint64(nb)).astype(np.int64) + 1
        pk[(u % np.uint64(100)) < 30] = 7                       # 30 % of the probe rows hit one key, 15 % another
        pk[((u % np.uint64(100)) >= 30) & ((u % np.uint64(100)) < 45)] = 13
        pv = (datagen.row_hash(9, 2, rows) % np.uint64(4096)).astype(np.float64) / 32.0
        bk = np.arange(1, nb + 1, dtype=np.int64)
        bw = (datagen.row_hash(9, 3, np.arange(nb, dtype=np.uint64)) % np.uint64(256)).astype(np.float64) / 8.0
        bg = (np.arange(nb) % 5).astype(np.int64)
        ora2 = ORA.Oracle()
        ora2.add_table("probe", [("p.k", bq.INT64, pk), ("p.v", bq.DOUBLE, pv)])
        ora2.add_table("build", [("b.k", bq.INT64, bk), ("b.w", bq.DOUBLE, bw), ("b.g", bq.INT64, bg)])
        joins = [("SELECT COUNT(*), SUM(p.v * b.w) FROM probe p JOIN build b ON p.k = b.k", None),
                 ("SELECT b.g, COUNT(*) AS n, SUM(p.v) AS s FROM probe p JOIN build b ON p.k = b.k GROUP BY b.g ORDER BY b.g", [(0, True)]),
                 ("SELECT p.k, COUNT(*) AS n, SUM(b.w) AS s FROM probe p JOIN build b ON p.k = b.k GROUP BY p.k ORDER BY n DESC LIMIT 2", [(1, False)])]
        for skew in (False, True):
            plo, phi = shard(npr, rank, world, skew)
            blo, bhi = shard(nb, rank, world, skew)
            for mode in ("shuffle", "broadcast"):
                os.environ["BOSQL_JOIN"] = mode
                for stats in (True, False):
                    reinstall()
                    eng = bq.Engine()
                    eng.add_table("probe", [("p.k", bq.INT64, np.ascontiguousarray(pk[plo:phi])), ("p.v", bq.DOUBLE, np.ascontiguousarray(pv[plo:phi]))],
                                  stats={"p.k": (1, nb, nb)} if stats else None)
                    eng.add_table("build", [("b.k", bq.INT64, np.ascontiguousarray(bk[blo:bhi])), ("b.w", bq.DOUBLE, np.ascontiguousarray(bw[blo:bhi])),
                                            ("b.g", bq.INT64, np.ascontiguousarray(bg[blo:bhi]))], stats={"b.k": (1, nb, nb)} if stats else None)
                    for sql, order in joins:
                        tag = f"skew_join[{mode},{'skew' if skew else 'even'},{'stats' if stats else 'nostats'}] {sql[7:30]}"
                        try:
                            got, want = eng.query(sql), ora2.query(sql)
                            assert_same_rows(got.cols, want.cols, ordered_by=order, what=tag)
                            results.append((tag, "ok"))
                        except Exception:  # noqa: BLE001
                            results.append((tag, traceback.format_exc()[-1200:]))
                    del eng
        os.environ.pop("BOSQL_JOIN", None)
    except Exception:  # noqa: BLE001
        results.append(("skew_join", traceback.format_exc()[-1500:]))

    # ---- high-cardinality GROUP BY: rows are shuffled by key hash, every rank aggregates the keys it owns ------------
    try:
        rows = np.arange(BIG_ROWS, dtype=np.uint64)
        k = (datagen.row_hash(5, 0, rows) % np.uint64(BIG_KEYS)).astype(np.int64) * 7 - 3_000_000
        v = ((datagen.row_hash(5, 1, rows) % np.uint64(1 << 20)).astype(np.float64)) / 64.0       # dyadic: sums are exact
        lo, hi = shard(BIG_ROWS, rank, world, False)
        uk, inv = np.unique(k, return_inverse=True)
        cnt = np.bincount(inv, minlength=len(uk)).astype(np.int64)
        sm = np.bincount(inv, weights=v, minlength=len(uk))
        # "peer": the partition kernel writes straight into the owning rank's buffer (CUDA IPC); "collective": partition into
        # a send buffer, then the host's all-to-all.  Same rows, same owners.
        for mode in ("peer", "collective"):
            os.environ["BOSQL_SHUFFLE"] = mode
            for keep in (True, False):
                cur = reinstall(keep)
                eng = bq.Engine()
                eng.add_table("t", [("k", bq.INT64, np.ascontiguousarray(k[lo:hi])), ("v", bq.DOUBLE, np.ascontiguousarray(v[lo:hi]))])
                before = dict(cur.calls)
                got = eng.query("SELECT k, SUM(v), COUNT(*) FROM t GROUP BY k")
                moved = cur.calls["all_to_all_v"] - before["all_to_all_v"]
                assert moved == (2 if mode == "collective" else 0), f"{mode}: {moved} all-to-all calls"
                cols = gather_rows(got.cols) if keep else got.cols
                order = np.argsort(cols[0], kind="stable")
                assert np.array_equal(cols[0][order], uk), "group keys differ"
                assert np.array_equal(cols[2][order], cnt), "counts differ"
                assert np.array_equal(cols[1][order], sm), "sums differ (dyadic values: must be exact)"
                if keep:
                    owned = len(got.cols[0])
                    assert 0 < owned < len(uk), "keep_sharded: a rank should own a strict subset of the groups (was the shuffle taken?)"
                results.append((f"shuffle_groupby[{mode},keep_sharded={keep}]", "ok"))
                del eng
        os.environ.pop("BOSQL_SHUFFLE", None)
    except Exception:  # noqa: BLE001
        results.append(("shuffle_groupby", traceback.format_exc()[-1500:]))
    # ---- ORDER BY without LIMIT, range-partitioned: rank r ends up with the r-th key range of the ordered result ----------
    try:
        reinstall()
        os.environ["BOSQL_SORT"] = "range"
        for skew in (False, True):
            eng = engine(skew, "stats")
            for sql, order in (("SELECT l.order_id, l.sku, l.qty FROM lineitem l WHERE l.qty > 40 ORDER BY l.order_id, l.sku, l.qty", [(0, True), (1, True), (2, True)]),
                               ("SELECT o.order_id, o.total FROM orders o ORDER BY o.total DESC, o.order_id", [(1, False), (0, True)]),
                               ("SELECT l.sku, l.price * 2 AS p2 FROM lineitem l WHERE l.sku < 3 ORDER BY p2 DESC", [(1, False)])):
                tag = f"range_sort[{'skew' if skew else 'even'}] {sql[7:40]}"
                try:
                    got, want = eng.query(sql), ora.query(sql)
                    parts = [None] * world
                    dist.all_gather_object(parts, [c.copy() for c in got.cols])
                    # concatenation in RANK order must be the ordered result
                    cols = [np.concatenate([pp[i] for pp in parts]) for i in range(len(got.cols))]
                    assert_same_rows(cols, want.cols, ordered_by=order, what=tag)
                    if world > 1 and not skew:
                        assert sum(1 for pp in parts if len(pp[0])) > 1, "every row ended up on one rank: was the range partition taken?"
                    results.append((tag, "ok"))
                except Exception:  # noqa: BLE001
                    results.append((tag, traceback.format_exc()[-1200:]))
            del eng
        os.environ.pop("BOSQL_SORT", None)
    except Exception:  # noqa: BLE001
        results.append(("range_sort", traceback.format_exc()[-1500:]))
    if not native:
        D.uninstall(xl)
    if ex.error:
        results.append(("exchange_callbacks", ex.error))


def main():
    import torch.distributed as dist
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl")
    rank, world = dist.get_rank(), dist.get_world_size()
    results = _Log(rank)
    try:
        run_rank(rank, world, "nccl", results)
    finally:
        bad = [r for r in results if r[1] != "ok"]
        print(f"[rank {rank}] {len(results) - len(bad)} ok, {len(bad)} failed", flush=True)
        for name, msg in bad:
            print(f"[rank {rank}] FAIL {name}: {msg}", flush=True)
        dist.destroy_process_group()
    sys.exit(1 if bad or not results else 0)


if __name__ == "__main__":
    main()
