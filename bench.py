#!/usr/bin/env python
"""bench.py — the hot path's headline measurement (BASELINE.json: rows/s and HBM GB/s for Q1/Q2, beside the CPU reference).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--rows R] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Headline (`value`): Q1 revenue-by-day — WHERE status = 'COMPLETE' AND order_date >= .. AND order_date <= ..
GROUP BY order_date SUM(total) ORDER BY order_date — over a synthetic orders table of R rows per GPU (default 1e9)
that is resident in HBM when the timed region starts.  One step = one complete query through the operator layer
(plan.run(): open / next* / close, result rows back on the host).  Inputs (16 GB) are far larger than the 126 MB L2, so
no flush is needed between steps.  Weak scaling: every rank owns its own R-row shard of an N*R-row table (rows are
generated from the global row index), partial aggregates are exchanged with one NCCL all-gather and merged.

The same JSON line carries: `e2e` (same query with HOST-resident pinned columns, host->device copy inside every step),
`roofline` (the fused scan kernel's algorithmic bytes / its CUDA-event duration, against MEASURED_PEAKS.json),
`cpu_baseline` (the compiled reference executor on a bounded sample, one thread), `q2` (the join query, same treatment),
`clocks`, `gpu_launches`.  `--impl reference` times the reference's own CPU executor (oracle/_ref) on the same query.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

Q1_SQL = ("SELECT order_date, SUM(total) AS revenue FROM orders WHERE status = 'COMPLETE' AND order_date >= 20240101 "
          "AND order_date <= 20240131 GROUP BY order_date ORDER BY order_date")
Q2_SQL = ("SELECT l.sku, SUM(l.qty * l.price) AS rev FROM lineitem l JOIN orders o ON l.order_id = o.order_id "
          "WHERE o.status = 'COMPLETE' GROUP BY l.sku ORDER BY rev DESC LIMIT 20")
Q1_BYTES_PER_ROW = 16          # status 4 + order_date 4 + total 8   
Q2_BYTES_PER_PROBE_ROW = 32    # l.order_id 8 + l.sku 8 + l.qty 8 + l.price 8
Q2_BYTES_PER_BUILD_ROW = 12    # o.order_id 8 + o.status 4
SEED = 20240101
N_SKU = 100_000


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# The contract is ONE JSON line on stdout.  Libraries (NCCL prints its version banner) also write to fd 1, so the real
# stdout is set aside for the result line and everything else is sent to stderr.
_RESULT_FD = os.dup(1)
os.dup2(2, 1)


def emit(obj):
    os.write(_RESULT_FD, (json.dumps(obj) + "\n").encode())


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


class ClockSampler:
    """SM clock and throttle reasons while the timed region runs: NVML polled every 5 ms from a thread (the timed region of
    the default run is ~50 ms), or `nvidia-smi -lms 50` when NVML cannot be loaded."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NVML_REASONS = (("hw_slowdown", 0x8), ("sw_thermal_slowdown", 0x20), ("hw_thermal_slowdown", 0x40), ("sw_power_cap", 0x4))

    def __init__(self, device):
        self.device, self.rows, self.proc = device, [], None
        self.t0 = self.t1 = None
        self.nvml = None
        self.samples = []            # (time, sm_mhz, reasons bitmask)
        self.max_mhz = None
        self._stop = False

    def start(self):
        """Started BEFORE the warm-up so that sampling is already running when the timed region begins."""
        try:
            import pynvml
            pynvml.nvmlInit()
            index = self.device
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            if vis:
                index = int(vis.split(",")[self.device])
            h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            self.nvml = (pynvml, h)
            threading.Thread(target=self._poll, daemon=True).start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                                          "-i", str(self.device)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _poll(self):
        pynvml, h = self.nvml
        while not self._stop:
            try:
                self.samples.append((time.time(), float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)),
                                     int(pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h))))
            except Exception:
                pass
            time.sleep(0.005)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def mark_begin(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def stop(self):
        if self.nvml:
            self._stop = True
            inside = [s for s in self.samples if self.t0 is not None and self.t0 <= s[0] <= (self.t1 or s[0])]
            window = "timed region"
            if not inside:
                inside, window = self.samples[-4:], "around the timed region"
            reasons = sorted({name for _, _, bits in inside for name, mask in self.NVML_REASONS if bits & mask})
            sm = [s[1] for s in inside]
            return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": self.max_mhz, "reasons": reasons,
                    "samples": len(sm), "window": window, "source": "NVML, 5 ms period"}
        if self.proc:
            time.sleep(0.12)
            self.proc.terminate()
        inside = [r for t, r in self.rows if self.t0 is not None and self.t0 <= t <= (self.t1 or t) + 0.06]
        window = "timed region"
        if not inside:                       # region shorter than the sampling period: the nearest samples around it
            inside = [r for _, r in self.rows][-4:]
            window = "around the timed region (it is shorter than the 50 ms sampling period)"

        def num(x):
            try:
                return float(x)
            except ValueError:
                return None
        sm = [num(r[1]) for r in inside if len(r) > 8 and num(r[1]) is not None]
        mx = [num(r[2]) for r in inside if len(r) > 8 and num(r[2]) is not None]
        reasons = set()
        for r in inside:
            if len(r) > 8:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "window": window, "source": "nvidia-smi -lms 50"}


# ---------------------------------------------------------------------------------------------------------------------
def run_reference(args):
    """The reference's own CPU executor (oracle/_ref: its unmodified sources, compiled) on the same query, one thread —
    it is single-threaded by design (README.md:3 of the reference).  Each step runs Q1 over a bounded sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import datagen, ref_engine
    if not ref_engine.available():
        emit({"impl": "reference", "unavailable": "oracle/_ref/libbosql_ref.so was not built (needs /root/reference at build time)"})
        return
    sample = int(min(args.rows, args.ref_rows))
    schema = datagen.orders_schema(sample)
    tab = datagen.host_table(schema, sample, SEED)
    eng = ref_engine.RefEngine()
    eng.add_table("orders", tab, eng.new_dict(datagen.STATUS_DICT))
    for _ in range(args.warmup):
        eng.query(Q1_SQL)
    secs = []
    t0 = time.perf_counter()
    for _ in range(args.steps):
        secs.append(eng.query(Q1_SQL).seconds)
    wall = time.perf_counter() - t0
    per_step = wall / args.steps
    value = sample / per_step
    cores = 1
    out = {
        "impl": "reference", "metric": "q1_rows_per_sec", "value": value, "unit": "rows/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": per_step * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "Q1 revenue-by-day (status = 'COMPLETE' AND order_date range, GROUP BY order_date SUM(total), ORDER BY)",
                   "rows_per_step": sample, "sql": Q1_SQL, "executor": "oracle/_ref: reference sources compiled unmodified, 1 thread"},
        "cpu_baseline": {"value": value, "unit": "rows/s", "cores": cores, "kind": "reference",
                         "sample": f"{sample} rows of the synthetic orders table per step (open..close {statistics.median(secs):.3f}s median)",
                         "host_cores_available": os.cpu_count()},
        "e2e": {"value": value, "unit": "rows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(out)


# ---------------------------------------------------------------------------------------------------------------------
class CudaArray:
    """__cuda_array_interface__ over a raw device pointer so torch can view a bq column without copying."""

    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    from __graft_entry__ import load_package

    bq = load_package()
    from bosql_b200 import synthetic as datagen        # workload definitions only; the oracle is imported in the cpu_baseline leg
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    xl = bq.exec_lib()
    if xl.bqx_init(local):
        raise RuntimeError(xl.bqx_last_error().decode())
    ctx = bq.wrap_context(xl.bqx_context())
    stream = torch.cuda.Stream()
    ctx.set_stream(stream.cuda_stream)
    rows = int(args.rows)
    row0 = rank * rows                      # this rank's shard of the N*R-row table
    peak, peak_src = measured_peak()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed(step, steps, warmup, profile=False):
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        for _ in range(warmup):
            step()
        barrier()
        sampler.mark_begin()
        l0 = ctx.launches
        if profile:
            ctx.profile_read()
            ctx.profile(True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            step()
        e1.record(stream)
        barrier()
        sampler.mark_end()
        ms = max_over_ranks(e0.elapsed_time(e1))
        kern = None
        if profile:
            ctx.profile(False)
            kern = ctx.profile_read()
        clocks = sampler.stop() if rank == 0 else None
        return ms / steps, ctx.launches - l0, kern, clocks

    # ---- synthetic tables, generated in HBM --------------------------------------------------------------------------
    def gen_table(schema, n, seed, r0):
        cols = {}
        for i, (name, typ, spec) in enumerate(schema):
            cols[name] = ctx.alloc(typ, n).generate(seed=seed, stream=i, row0=r0, **spec)
        ctx.sync()
        return cols

    t0 = time.perf_counter()
    orders = gen_table(datagen.orders_schema(rows), rows, SEED, row0)
    log(f"[rank {rank}] generated orders ({rows} rows) in {time.perf_counter() - t0:.2f}s")
    DATE_MIN, DATE_MAX = 20240101, 20241228

    eng = bq.Engine()
    sdict = eng.new_dict(datagen.STATUS_DICT)
    q1_cols = [("status", bq.STRING, orders["status"]), ("order_date", bq.DATE32, orders["order_date"]),
               ("total", bq.DOUBLE, orders["total"])]
    eng.add_table("orders", q1_cols, sdict, stats={"order_date": (DATE_MIN, DATE_MAX, 336), "total": (1.0, 1000.0, 99901)})
    q1_plan = eng.plan(Q1_SQL)
    result = {}

    # One code path for every N: the SQL statement through the operator layer.  With N > 1 every rank holds a row shard
    # of `orders` (statistics describe the whole table) and the installed exchange makes HashAggregate all-gather and
    # merge the partial states in rank order (bo-sql_b200/host/exchange.cpp); every rank ends up with the full result.
    exchange = None
    if world > 1:
        from bosql_b200 import distributed as DIST
        exchange = DIST.install(xl, device="cuda")

    def q1_step():
        result["q1"] = q1_plan.run()

    ms_step, launches, kern, clocks = timed(q1_step, args.steps, max(3, args.warmup), profile=True)
    total_rows = rows * world
    value = total_rows / (ms_step * 1e-3)
    k_launches, k_ms = kern
    k_avg_ms = k_ms / max(1, k_launches)
    achieved = Q1_BYTES_PER_ROW * rows / (k_avg_ms * 1e-3) / 1e9 if k_launches else None
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "q1_scan_traffic.json")))["dram_bytes_per_launch"]
    except Exception:
        pass
    out = {
        "metric": "q1_rows_per_sec", "value": value, "unit": "rows/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(3, args.warmup), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "Q1 revenue-by-day (status = 'COMPLETE' AND order_date range, GROUP BY order_date SUM(total), ORDER BY) "
                               f"on {rows} synthetic orders rows per GPU",
                   "rows_per_gpu": rows, "sql": Q1_SQL, "parallelism": f"row-range x{world}",
                   "l2": "inputs (16 GB per GPU) exceed the 126 MB L2; no flush needed", "seed": SEED,
                   "algorithmic_bytes_per_row": Q1_BYTES_PER_ROW},
        "gbs_whole_query": Q1_BYTES_PER_ROW * total_rows / (ms_step * 1e-3) / 1e9,
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": "bq::k_scan (fused scan+selection+dense GROUP BY)", "achieved": achieved,
                     "peak": peak, "unit": "GB/s", "frac": (achieved / peak) if achieved else None, "peak_source": peak_src,
                     "traffic": traffic, "launches_timed": int(k_launches), "avg_launch_ms": k_avg_ms,
                     "algorithmic_bytes_per_launch": Q1_BYTES_PER_ROW * rows},
    }

    # ---- parity + cpu_baseline on a bounded sample (rank 0, N = 1 only) -------------------------------------------
    if world == 1 and not args.no_cpu:
        from oracle import ref_engine
        from tests.parity import assert_same_rows
        sample = int(min(rows, args.ref_rows))
        host = [(name, typ, col.to_numpy(0, sample)) for name, typ, col in q1_cols]
        geng = bq.Engine()
        geng.add_table("orders", host, geng.new_dict(datagen.STATUS_DICT))
        got = geng.query(Q1_SQL)
        if ref_engine.available():
            reng = ref_engine.RefEngine()
            reng.add_table("orders", host, reng.new_dict(datagen.STATUS_DICT))
            runs = [reng.query(Q1_SQL) for _ in range(3)]
            sec = statistics.median(r.seconds for r in runs)
            assert_same_rows(got.cols, runs[0].cols, ordered_by=[(0, True)], what="bench Q1 sample vs reference")
            out["cpu_baseline"] = {"value": sample / sec, "unit": "rows/s", "cores": 1, "kind": "reference",
                                   "sample": f"first {sample} rows of the same table, reference open..close median of 3 = {sec:.3f}s",
                                   "parity_on_sample": "ok (keys exact, SUM within 1e-12)"}
        else:
            out["cpu_baseline"] = {"value": None, "unit": "rows/s", "cores": 1, "kind": "reference", "sample": "oracle/_ref not built"}
        del geng, host

    # ---- e2e: the same query with HOST-resident columns; every step pays the host->device copy ----------------------
    if not args.no_e2e:
        e2e_rows = rows
        widths = {bq.STRING: 4, bq.DATE32: 4, bq.DOUBLE: 8, bq.INT64: 8}
        bufs = []
        try:
            host_cols = []
            for name, typ, col in q1_cols:
                p = ctx.host_alloc(widths[typ] * e2e_rows)
                bufs.append(p)
                import ctypes as C
                if bq.kernel_lib().bq_col_read(ctx.h, col.h, 0, e2e_rows, C.c_void_p(p)):
                    raise RuntimeError(bq.kernel_lib().bq_last_error().decode())
                host_cols.append((name, typ, (p, e2e_rows)))
            heng = bq.Engine()
            heng.add_table("orders", host_cols, heng.new_dict(datagen.STATUS_DICT),
                           stats={"order_date": (DATE_MIN, DATE_MAX, 336)})
            hplan = heng.plan(Q1_SQL)
            res = {}

            def e2e_step():
                heng.evict_device("orders")
                res["r"] = hplan.run()
            e2e_steps = max(1, min(args.steps, args.e2e_steps))
            ms_e2e, _, _, _ = timed(e2e_step, e2e_steps, 1)
            d2h = sum(c.nbytes for c in res["r"].cols)
            out["e2e"] = {"value": e2e_rows * world / (ms_e2e * 1e-3), "unit": "rows/s",
                          "h2d_bytes_per_step": Q1_BYTES_PER_ROW * e2e_rows, "d2h_bytes_per_step": int(d2h),
                          "ms_per_step": ms_e2e, "steps": e2e_steps,
                          "path": "Engine over pinned host columns (bqx_table_add_borrowed_column); mirrors evicted before every step"}
            del hplan, heng
        except Exception as e:  # noqa: BLE001
            out["e2e"] = {"value": None, "unit": "rows/s", "error": str(e)[:200]}
        finally:
            for p in bufs:
                ctx.host_free(p)

    # ---- Q2 (N = 1): lineitem(R) JOIN orders(R/4) ------------------------------------------------------------------------
    if not args.no_q2:
        try:
            del q1_plan, eng
            for c in ("total", "order_date"):
                orders[c].free()
            n_orders = max(1, rows // 4)                  # per GPU; both tables are sharded by row range
            n_orders_all = n_orders * world
            o2 = gen_table(datagen.orders_schema(n_orders_all, prefix="o.")[:2], n_orders, SEED + 1, rank * n_orders)
            li = gen_table(datagen.lineitem_schema(n_orders_all, N_SKU), rows, SEED + 2, row0)
            e2 = bq.Engine()
            d2 = e2.new_dict(datagen.STATUS_DICT)
            e2.add_table("orders", [("o.order_id", bq.INT64, o2["o.order_id"]), ("o.status", bq.STRING, o2["o.status"])], d2,
                         stats={"o.order_id": (1, n_orders_all, n_orders_all)})
            e2.add_table("lineitem", [(n, t, li[n]) for n, t, _ in datagen.lineitem_schema(n_orders_all, N_SKU)], d2,
                         stats={"l.sku": (0, N_SKU - 1, N_SKU), "l.order_id": (1, n_orders_all, n_orders_all)})
            p2 = e2.plan(Q2_SQL)
            r2 = {}

            def q2_step():
                r2["r"] = p2.run()
            ms2, l2, kern2, _ = timed(q2_step, max(3, args.steps // 2), 3, profile=True)
            q2_bytes = Q2_BYTES_PER_PROBE_ROW * rows + Q2_BYTES_PER_BUILD_ROW * n_orders
            k2n, k2ms = kern2
            k2avg = k2ms / max(1, k2n)
            out["q2"] = {"metric": "q2_rows_per_sec", "value": (rows + n_orders) * world / (ms2 * 1e-3), "unit": "rows/s (probe + build)",
                         "ms_per_step": ms2, "gbs_whole_query": q2_bytes * world / (ms2 * 1e-3) / 1e9, "gpu_launches": int(l2),
                         "rows": {"lineitem_per_gpu": rows, "orders_per_gpu": n_orders, "sku": N_SKU}, "sql": Q2_SQL,
                         "join": "bitmap over the global o.order_id domain" + (", per-rank bitmaps summed over NVLink" if world > 1 else ""),
                         "roofline": {"bound": "hbm", "kernel": "bq::k_scan (probe + GROUP BY sku)",
                                      "achieved": Q2_BYTES_PER_PROBE_ROW * rows / (k2avg * 1e-3) / 1e9 if k2n else None,
                                      "peak": peak, "unit": "GB/s", "avg_launch_ms": k2avg,
                                      "frac": (Q2_BYTES_PER_PROBE_ROW * rows / (k2avg * 1e-3) / 1e9 / peak) if k2n else None},
                         "top": [int(x) for x in r2["r"].cols[0][:5]]}
        except Exception as e:  # noqa: BLE001
            out["q2"] = {"error": str(e)[:300]}

    if exchange is not None:
        out["exchange"] = {"calls": exchange.calls, "bytes_sent_per_rank": int(exchange.bytes_sent), "error": exchange.error}
    if rank == 0:
        emit(out)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--rows", type=float, default=1e9, help="rows per GPU")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--ref-rows", type=float, default=2e7, help="sample size for the CPU reference")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-q2", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
