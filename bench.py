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


def bind_host_memory_near_gpu(device_index):
    """Pinned host buffers of the end-to-end leg should live on the NUMA node the GPU hangs off: with eight ranks copying
    16 GB each at once, buffers that all landed on one socket share that socket's memory controllers and the inter-socket
    link (round 1: 290 ms per step on one GPU, 693 ms on eight).  Sets this process's memory policy to PREFER the GPU's node
    and its CPU affinity to that node's cores; returns what it did (for the JSON line) or why it could not."""
    import ctypes
    try:
        import pynvml
        pynvml.nvmlInit()
        index = device_index
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            index = int(vis.split(",")[device_index])
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(index)).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        dom, rest = bus.split(":", 1)
        path = f"/sys/bus/pci/devices/{int(dom, 16):04x}:{rest.lower()}/numa_node"
        node = int(open(path).read().strip())
        if node < 0:
            return {"bound": False, "why": f"{path} reports no NUMA node"}
        mask = ctypes.c_ulong(1 << node)
        libc = ctypes.CDLL("libc.so.6", use_errno=True)
        rc = libc.syscall(238, 1, ctypes.byref(mask), ctypes.c_ulong(64))          # set_mempolicy(MPOL_PREFERRED, {node})
        if rc != 0:
            return {"bound": False, "why": f"set_mempolicy failed (errno {ctypes.get_errno()})"}
        cpus = []
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.extend(range(int(lo), int(hi or lo) + 1))
        try:
            os.sched_setaffinity(0, set(cpus) & os.sched_getaffinity(0) or os.sched_getaffinity(0))
        except OSError:
            pass
        return {"bound": True, "numa_node": node, "gpu_pci": bus, "policy": "MPOL_PREFERRED + CPU affinity to the node"}
    except Exception as e:  # noqa: BLE001
        return {"bound": False, "why": str(e)[:120]}


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
C2_TYPES = ("INT64", "DOUBLE", "STRING", "DATE32")
TRAFFIC_FILES = {"q1": "q1_scan_traffic.json", "q2": "q2_scan_traffic.json"}


def stamped_traffic(which, lib_path):
    """DRAM bytes per launch from the committed ncu capture - only while the kernel's SASS still hashes to the value
    recorded with that capture (scripts/sass_hash.py); a changed kernel drops the number instead of repeating it."""
    try:
        rec = json.load(open(os.path.join(ROOT, "profiles", TRAFFIC_FILES[which])))
    except Exception:  # noqa: BLE001
        return None, "no ncu capture committed for this kernel"
    want = rec.get("sass_sha256")
    if not want:
        return None, "capture carries no SASS stamp"
    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    from sass_hash import kernel_sass_hash
    have = kernel_sass_hash(lib_path, rec["kernel_regex"])
    if have != want:
        return None, f"stale: kernel SASS changed since the capture ({rec.get('source')})"
    return rec["dram_bytes_per_launch"], f"ncu --set full, {rec.get('source')}; SASS stamp matches the built library"


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    from __graft_entry__ import load_package

    bq = load_package()
    from bosql_b200 import synthetic as datagen        # workload definitions only; the oracle is imported in the checker legs
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    host_numa = bind_host_memory_near_gpu(local) if not args.no_numa else {"bound": False, "why": "--no-numa"}
    host_cpus = None
    if world > 1 and not args.no_pin:
        # Every rank's main thread spins in cudaStreamSynchronize while its kernel runs and NCCL keeps a proxy thread per
        # communicator: on a box with two host cores per GPU the ranks get in each other's way between steps (measured: the
        # device timeline of a Q1 step is 2.26 ms at N = 8, the step loop 2.33-2.58 ms).  Give every rank its own cores.
        try:
            avail = sorted(os.sched_getaffinity(0))
            per = max(1, len(avail) // world)
            mine = avail[local * per:(local + 1) * per] or avail
            os.sched_setaffinity(0, set(mine))
            host_cpus = {"pinned_to": mine, "of": len(avail)}
        except Exception as e:  # noqa: BLE001
            host_cpus = {"pinned_to": None, "why": str(e)[:80]}
    xl = bq.exec_lib()
    if xl.bqx_init(local):
        raise RuntimeError(xl.bqx_last_error().decode())
    ctx = bq.wrap_context(xl.bqx_context())
    stream = torch.cuda.Stream()
    ctx.set_stream(stream.cuda_stream)
    rows = int(args.rows)
    peak, peak_src = measured_peak()
    NVLINK_PEAK = 770.0            # GB/s per direction per GPU, measured peer copy on this pool (B200_PROFILING.md; 900 nominal)
    t_start = time.perf_counter()

    # One code path for every N: the SQL statement through the operator layer.  With N > 1 every rank holds a row shard of
    # each table (statistics describe the whole table) and the library's own NCCL exchange (bq_comm_*, no Python between a
    # plan and its collectives) runs at the plan's exchange points; every rank ends up with the full result.
    exchange = None
    if world > 1:
        from bosql_b200 import distributed as DIST
        exchange = DIST.install_native(xl)

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

    def sum_over_ranks(x):
        if world == 1:
            return int(x)
        t = torch.tensor([int(x)], dtype=torch.int64, device="cuda")
        dist.all_reduce(t)
        return int(t.item())

    def timed(step, steps, warmup, profile=False, clocks=False):
        """W warm-up steps, then K steps between barriers, CUDA events on the launching stream, max over ranks."""
        sampler = ClockSampler(local) if (clocks and rank == 0) else None
        if sampler:
            sampler.start()
        for _ in range(warmup):
            step()
        barrier()
        if sampler:
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
        if sampler:
            sampler.mark_end()
        ms = max_over_ranks(e0.elapsed_time(e1))
        kern = None
        if profile:
            ctx.profile(False)
            kern = ctx.profile_read()
        return ms / steps, (ctx.launches - l0) // steps, kern, (sampler.stop() if sampler else None)

    def gen_table(schema, n, seed, r0):
        cols = {}
        for i, (name, typ, spec) in enumerate(schema):
            cols[name] = ctx.alloc(typ, n).generate(seed=seed, stream=i, row0=r0, **spec)
        ctx.sync()
        return cols

    def free_table(cols):
        for c in cols.values():
            c.free()

    def exchange_counters():
        return (dict(exchange.calls), exchange.bytes_sent) if exchange else ({}, 0)

    def exchange_delta(before, runs):
        if not exchange:
            return None
        calls0, sent0 = before
        calls1, sent1 = exchange_counters()
        return {"collectives_per_step": {k: (calls1[k] - calls0.get(k, 0)) // runs for k in calls1 if calls1[k] != calls0.get(k, 0)},
                "collective_bytes_sent_per_gpu_per_step": (sent1 - sent0) // runs}

    DATE_MIN, DATE_MAX = 20240101, 20241228

    # ---- workloads: (tables, statistics, SQL) as functions of the global size, so that the full-size run, the strong-
    # scaling run and the small parity sample are THE SAME statements over the same generator -----------------------------
    def q1_tables(n_all, n_local, r0):
        o = gen_table(datagen.orders_schema(n_all), n_local, SEED, r0)
        cols = [("status", bq.STRING, o["status"]), ("order_date", bq.DATE32, o["order_date"]), ("total", bq.DOUBLE, o["total"])]
        o["order_id"].free()
        return {"orders": (cols, {"order_date": (DATE_MIN, DATE_MAX, 336), "total": (1.0, 1000.0, 99901)})}

    def q2_tables(n_line_all, n_line, line_r0, n_ord_all, n_ord, ord_r0, key_stride=1):
        """key_stride > 1: order ids 1, 1 + s, 1 + 2s, ... - a sparse domain, so the join needs a real hash table"""
        o = gen_table(datagen.orders_schema(n_ord_all, prefix="o.", key_stride=key_stride)[:2], n_ord, SEED + 1, ord_r0)
        li = gen_table(datagen.lineitem_schema(n_ord_all, N_SKU, key_stride=key_stride), n_line, SEED + 2, line_r0)
        key_max = 1 + (n_ord_all - 1) * key_stride
        return {"orders": ([("o.order_id", bq.INT64, o["o.order_id"]), ("o.status", bq.STRING, o["o.status"])],
                           {"o.order_id": (1, key_max, n_ord_all)}),
                "lineitem": ([(n, t, li[n]) for n, t, _ in datagen.lineitem_schema(n_ord_all, N_SKU)],
                             {"l.sku": (0, N_SKU - 1, N_SKU), "l.order_id": (1, key_max, n_ord_all)})}

    def c2_tables(n_all, n_local, r0):
        s = gen_table(datagen.sweep_schema_c2(), n_local, SEED + 3, r0)
        return {"sweep": ([(n, t, s[n]) for n, t, _ in datagen.sweep_schema_c2()], {})}

    def c4_tables(n_all, n_local, r0, ids=None):
        ids = ids or max(16, n_all // 20)                # 100 M distinct keys over 2 B rows
        k = ctx.alloc(bq.INT64, n_local).generate(dist=bq.GEN_HASHED, seed=SEED + 4, stream=0, lo=0, hi=ids - 1, modulus=1 << 61, row0=r0)
        v = ctx.alloc(bq.DOUBLE, n_local).generate(dist=bq.GEN_UNIFORM_DIV, seed=SEED + 4, stream=1, lo=1, hi=6400, div=64.0, row0=r0)
        ctx.sync()
        return {"t": ([("k", bq.INT64, k), ("v", bq.DOUBLE, v)], {"k": (0, (1 << 61) - 1, ids)})}

    def c5_tables(n_probe_all, n_probe, probe_r0, n_build_all, n_build, build_r0):
        bk = ctx.alloc(bq.INT64, n_build).generate(dist=bq.GEN_SEQ, seed=SEED + 5, stream=0, lo=1, row0=build_r0)
        bw = ctx.alloc(bq.DOUBLE, n_build).generate(dist=bq.GEN_UNIFORM_DIV, seed=SEED + 5, stream=1, lo=1, hi=64, div=4.0, row0=build_r0)
        cdf, starts = datagen.zipf_buckets(n_build_all, 1.1)           # Zipf(1.1) over the WHOLE build-key domain
        pk = ctx.alloc(bq.INT64, n_probe).generate(dist=bq.GEN_BUCKETS, seed=SEED + 6, stream=0, lo=1, cdf=cdf, starts=starts, row0=probe_r0)
        pv = ctx.alloc(bq.DOUBLE, n_probe).generate(dist=bq.GEN_UNIFORM_DIV, seed=SEED + 6, stream=1, lo=1, hi=64, div=4.0, row0=probe_r0)
        ctx.sync()
        return {"build": ([("b.k", bq.INT64, bk), ("b.w", bq.DOUBLE, bw)], {"b.k": (1, n_build_all, n_build_all)}),
                "probe": ([("p.k", bq.INT64, pk), ("p.v", bq.DOUBLE, pv)], {"p.k": (1, n_build_all, n_build_all)})}

    C4_SQL = "SELECT k, SUM(v), COUNT(*), AVG(v) FROM t GROUP BY k"
    C5_SQL = "SELECT COUNT(*), SUM(p.v * b.w) FROM probe p JOIN build b ON p.k = b.k"

    def c2_queries():
        """(type, nominal selectivity, SQL): COUNT(*) + SUM(v) under one predicate on a column of each physical type."""
        qs = []
        for sel in (0.01, 0.10, 0.50, 0.90, 0.99):
            qs.append(("INT64", sel, f"SELECT COUNT(*), SUM(v) FROM sweep WHERE c_i64 < {int(sel * 1_000_000)}"))
            qs.append(("DOUBLE", sel, f"SELECT COUNT(*), SUM(v) FROM sweep WHERE c_f64 < {int(sel * 10_000)}"))
        for sel, pred in ((1 / 64, "c_str = 's5'"), (0.125, "c_str = 's2'"), (0.5, "c_str = 's0'"), (0.875, "c_str != 's2'"), (63 / 64, "c_str != 's5'")):
            qs.append(("STRING", sel, f"SELECT COUNT(*), SUM(v) FROM sweep WHERE {pred}"))
        for sel, lit in ((1 / 120, 20150201), (0.1, 20160101), (0.5, 20200101), (0.9, 20240101), (119 / 120, 20241201)):
            qs.append(("DATE32", sel, f"SELECT COUNT(*), SUM(v) FROM sweep WHERE c_date < {lit}"))
        return qs

    def make_engine(tables, dicts=None):
        eng = bq.Engine()
        d = eng.new_dict(dicts if dicts is not None else datagen.STATUS_DICT)
        for name, (cols, stats) in tables.items():
            eng.add_table(name, cols, d, stats=stats)
        return eng

    def drop(tables):
        for cols, _ in tables.values():
            for _n, _t, c in cols:
                c.free()

    out = {}
    errors = {}

    def section(name, fn):
        try:
            t0 = time.perf_counter()
            fn()
            log(f"[rank {rank}] {name}: {time.perf_counter() - t0:.1f}s")
        except Exception as e:  # noqa: BLE001
            import traceback
            errors[name] = str(e)[:300]
            log(f"[rank {rank}] {name} FAILED: {traceback.format_exc()[-1500:]}")
        barrier()

    # =================================================================================================================
    # Q1 (the headline): weak scaling, R rows per GPU
    # =================================================================================================================
    steps, warmup = args.steps, max(3, args.warmup)
    q1 = {}

    def run_q1():
        tabs = q1_tables(rows * world, rows, rank * rows)
        q1["tabs"] = tabs
        eng = make_engine(tabs)
        plan = eng.plan(Q1_SQL)
        res = {}

        def step():
            res["r"] = plan.run()
        before = exchange_counters()
        ms_step, launches, kern, clocks = timed(step, steps, warmup, profile=True, clocks=True)
        total_rows = rows * world
        k_launches, k_ms = kern
        k_avg_ms = k_ms / max(1, k_launches)
        achieved = Q1_BYTES_PER_ROW * rows / (k_avg_ms * 1e-3) / 1e9 if k_launches else None
        traffic, traffic_note = stamped_traffic("q1", bq.KERNEL_LIB) if rank == 0 else (None, None)
        out.update({
            "metric": "q1_rows_per_sec", "value": total_rows / (ms_step * 1e-3), "unit": "rows/s", "n_gpus": world, "steps": steps,
            "warmup": warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "Q1 revenue-by-day (status = 'COMPLETE' AND order_date range, GROUP BY order_date SUM(total), ORDER BY) "
                                   f"on {rows} synthetic orders rows per GPU",
                       "rows_per_gpu": rows, "sql": Q1_SQL, "parallelism": f"row-range x{world}",
                       "l2": "inputs (16 GB per GPU) exceed the 126 MB L2; no flush needed", "seed": SEED,
                       "algorithmic_bytes_per_row": Q1_BYTES_PER_ROW,
                       "exchange": "native NCCL inside libbosql_b200.so (bq_comm_*)" if world > 1 else "none (one GPU)"},
            "gbs_whole_query": Q1_BYTES_PER_ROW * total_rows / (ms_step * 1e-3) / 1e9,
            "gpu_launches": int(launches) * steps,
            "host_cpus": host_cpus,
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "bq::k_scan (fused scan+selection+dense GROUP BY)", "achieved": achieved,
                         "peak": peak, "unit": "GB/s", "frac": (achieved / peak) if achieved else None, "peak_source": peak_src,
                         "traffic": traffic, "traffic_source": traffic_note, "launches_timed": int(k_launches), "avg_launch_ms": k_avg_ms,
                         "algorithmic_bytes_per_launch": Q1_BYTES_PER_ROW * rows},
        })
        if exchange:
            out["exchange"] = exchange_delta(before, steps + warmup)
        q1["rows_result"] = res["r"].rows
        del plan, eng

    section("q1", run_q1)

    # ---- e2e: the same query with HOST-resident columns; every step pays the host->device copy ----------------------------
    def run_q1_e2e():
        import ctypes as C
        widths = {bq.STRING: 4, bq.DATE32: 4, bq.DOUBLE: 8, bq.INT64: 8}
        cols, _ = q1["tabs"]["orders"]
        bufs, host_cols = [], []
        try:
            for name, typ, col in cols:
                p = ctx.host_alloc(widths[typ] * rows)
                bufs.append(p)
                if bq.kernel_lib().bq_col_read(ctx.h, col.h, 0, rows, C.c_void_p(p)):
                    raise RuntimeError(bq.kernel_lib().bq_last_error().decode())
                host_cols.append((name, typ, (p, rows)))
            heng = bq.Engine()
            heng.add_table("orders", host_cols, heng.new_dict(datagen.STATUS_DICT), stats={"order_date": (DATE_MIN, DATE_MAX, 336)})
            hplan = heng.plan(Q1_SQL)
            res = {}

            def e2e_step():
                heng.evict_device("orders")
                res["r"] = hplan.run()
            e2e_steps = max(1, min(steps, args.e2e_steps))
            ms_e2e, _, _, _ = timed(e2e_step, e2e_steps, 1)
            d2h = sum(c.nbytes for c in res["r"].cols)
            out["e2e"] = {"value": rows * world / (ms_e2e * 1e-3), "unit": "rows/s", "h2d_bytes_per_step": Q1_BYTES_PER_ROW * rows,
                          "d2h_bytes_per_step": int(d2h), "ms_per_step": ms_e2e, "steps": e2e_steps,
                          "path": "Engine over pinned host columns (bqx_table_add_borrowed_column); mirrors evicted before every step",
                          "host_memory": host_numa}
            del hplan, heng
        finally:
            for p in bufs:
                ctx.host_free(p)

    if not args.no_e2e:
        section("q1_e2e", run_q1_e2e)
        if "q1_e2e" in errors:
            out["e2e"] = {"value": None, "unit": "rows/s", "error": errors["q1_e2e"]}
    if "tabs" in q1:
        drop(q1.pop("tabs"))

    # =================================================================================================================
    # Q2: lineitem JOIN orders.  weak = R lineitem + R/4 orders rows PER GPU (the key domain, and with it the join bitmap,
    # grows with N); strong = R + R/4 rows IN TOTAL, sharded N ways (N > 1 only)
    # =================================================================================================================
    def run_q2(key, n_line_all, n_line, line_r0, n_ord_all, n_ord, ord_r0, e2e=False, key_stride=1):
        tabs = q2_tables(n_line_all, n_line, line_r0, n_ord_all, n_ord, ord_r0, key_stride)
        try:
            eng = make_engine(tabs)
            plan = eng.plan(Q2_SQL)
            res = {}

            def step():
                res["r"] = plan.run()
            before = exchange_counters()
            q2_steps = max(3, steps // 2)
            ms2, l2, kern2, _ = timed(step, q2_steps, 3, profile=True)
            q2_bytes = Q2_BYTES_PER_PROBE_ROW * n_line + Q2_BYTES_PER_BUILD_ROW * n_ord
            k2n, k2ms = kern2
            probe_ms = k2ms / q2_steps              # all probe-side launches of one step (key-range passes + fused scan)
            bitmap_mb = (n_ord_all + 7) // 8 / 1e6
            traffic, traffic_note = (stamped_traffic("q2", bq.KERNEL_LIB) if (rank == 0 and key == "q2" and world == 1) else (None, "captured at N = 1 only"))
            rec = {"metric": "q2_rows_per_sec", "value": (n_line_all + n_ord_all) / (ms2 * 1e-3), "unit": "rows/s (probe + build)",
                   "ms_per_step": ms2, "steps": q2_steps, "gbs_whole_query": q2_bytes * world / (ms2 * 1e-3) / 1e9, "gpu_launches": int(l2) * q2_steps,
                   "rows": {"lineitem_per_gpu": n_line, "orders_per_gpu": n_ord, "lineitem_total": n_line_all, "orders_total": n_ord_all, "sku": N_SKU},
                   "sql": Q2_SQL,
                   "join": ("open-addressing hash table (sparse key domain: 16-byte slots, capacity 2 x build rows)" if key_stride > 1 else
                            f"bitmap over the global o.order_id domain ({bitmap_mb:.0f} MB)" +
                            (", per-rank bitmaps summed over NVLink and verified by popcount" if world > 1 else "") +
                            (", probed in key-range passes" if k2n // q2_steps > 1 else ", probed inside the fused scan")),
                   "roofline": {"bound": "hbm", "kernel": "probe side: bq::k_probe_bits passes + bq::k_scan (GROUP BY sku)" if k2n // q2_steps > 1
                                else "bq::k_scan (probe + GROUP BY sku)",
                                "achieved": Q2_BYTES_PER_PROBE_ROW * n_line / (probe_ms * 1e-3) / 1e9 if k2n else None,
                                "peak": peak, "unit": "GB/s", "probe_ms_per_step": probe_ms, "probe_launches_per_step": k2n // q2_steps,
                                "frac": (Q2_BYTES_PER_PROBE_ROW * n_line / (probe_ms * 1e-3) / 1e9 / peak) if k2n else None,
                                "algorithmic_bytes_per_step": Q2_BYTES_PER_PROBE_ROW * n_line, "traffic": traffic, "traffic_source": traffic_note},
                   "roofline_whole_query": {"achieved": q2_bytes / (ms2 * 1e-3) / 1e9, "frac": q2_bytes / (ms2 * 1e-3) / 1e9 / peak,
                                            "frac_of_8TBs": q2_bytes / (ms2 * 1e-3) / 1e9 / 8000.0},
                   "top": [int(x) for x in res["r"].cols[0][:5]]}
            if exchange:
                rec["exchange"] = exchange_delta(before, q2_steps + 3)
            out[key] = rec
            del plan, eng
            if e2e:
                run_q2_e2e(tabs, n_line, n_ord, n_ord_all, rec)
        finally:
            drop(tabs)

    def run_q2_e2e(tabs, n_line, n_ord, n_ord_all, rec):
        """Q2 from pinned HOST columns: 32 B x lineitem + 12 B x orders cross PCIe inside every step."""
        import ctypes as C
        widths = {bq.STRING: 4, bq.DATE32: 4, bq.DOUBLE: 8, bq.INT64: 8}
        bufs = []
        try:
            host_tabs = {}
            h2d = 0
            for tname, (cols, stats) in tabs.items():
                hc = []
                for name, typ, col in cols:
                    n = col.n
                    p = ctx.host_alloc(widths[typ] * n)
                    bufs.append(p)
                    if bq.kernel_lib().bq_col_read(ctx.h, col.h, 0, n, C.c_void_p(p)):
                        raise RuntimeError(bq.kernel_lib().bq_last_error().decode())
                    hc.append((name, typ, (p, n)))
                    h2d += widths[typ] * n
                host_tabs[tname] = (hc, stats)
            heng = make_engine(host_tabs)
            hplan = heng.plan(Q2_SQL)
            res = {}

            def step():
                heng.evict_device("orders")
                heng.evict_device("lineitem")
                res["r"] = hplan.run()
            ms, _, _, _ = timed(step, 2, 1)
            rec["e2e"] = {"value": (n_line + n_ord) / (ms * 1e-3), "unit": "rows/s (probe + build)", "h2d_bytes_per_step": int(h2d),
                          "d2h_bytes_per_step": int(sum(c.nbytes for c in res["r"].cols)), "ms_per_step": ms, "steps": 2,
                          "path": "Engine over pinned host columns; both tables' mirrors evicted before every step"}
            del hplan, heng
        except Exception as e:  # noqa: BLE001
            rec["e2e"] = {"value": None, "error": str(e)[:200]}
        finally:
            for p in bufs:
                ctx.host_free(p)

    if not args.no_q2:
        n_ord = max(1, rows // 4)
        section("q2", lambda: run_q2("q2", rows * world, rows, rank * rows, n_ord * world, n_ord, rank * n_ord,
                                     e2e=(world == 1 and not args.no_e2e)))

    # ---- the same query over SPARSE order ids: no bitmap, no direct-address table - the open-addressing hash join (N = 1) ----
    if not args.no_q2 and world == 1 and not args.no_stress:
        n_ord = max(1, rows // 4)
        section("q2_hash", lambda: run_q2("q2_hash", rows, rows, 0, n_ord, n_ord, 0, key_stride=7919))

    # ---- strong scaling: the 1 B-row tables IN TOTAL, sharded N ways --------------------------------------------------------
    def run_q1_strong():
        n_local = rows // world
        tabs = q1_tables(rows, n_local, rank * n_local)
        try:
            eng = make_engine(tabs)
            plan = eng.plan(Q1_SQL)
            ms, launches, kern, _ = timed(lambda: plan.run(), steps, warmup, profile=True)
            kn, kms = kern
            out["q1_strong"] = {"metric": "q1_rows_per_sec", "scaling": "strong", "value": n_local * world / (ms * 1e-3), "unit": "rows/s",
                                "rows_total": n_local * world, "rows_per_gpu": n_local, "ms_per_step": ms, "steps": steps,
                                "kernel_ms": kms / max(1, kn), "gbs_whole_query": Q1_BYTES_PER_ROW * n_local * world / (ms * 1e-3) / 1e9}
            del plan, eng
        finally:
            drop(tabs)

    if world > 1 and not args.no_strong:
        section("q1_strong", run_q1_strong)
        if not args.no_q2:
            nl, no = rows // world, max(1, rows // 4 // world)
            section("q2_strong", lambda: run_q2("q2_strong", nl * world, nl, rank * nl, no * world, no, rank * no))
            if "q2_strong" in out:
                out["q2_strong"]["scaling"] = "strong"

    # =================================================================================================================
    # C2: filter-only sweep, COUNT(*) + SUM(v) under one predicate per physical type, 1 % .. 99 %
    # =================================================================================================================
    def run_c2():
        tabs = c2_tables(rows * world, rows, rank * rows)
        try:
            eng = make_engine(tabs, dicts=datagen.C2_STR_DICT)
            per_type = {t: [] for t in C2_TYPES}
            for typ, sel, sql in c2_queries():
                plan = eng.plan(sql)
                res = {}

                def step():
                    res["r"] = plan.run()
                ms, launches, kern, _ = timed(step, 5, 2, profile=True)
                kn, kms = kern
                width = 16 if typ in ("INT64", "DOUBLE") else 12
                cnt = int(res["r"].cols[0][0]) if res["r"].rows else 0         # a global aggregate: every rank holds the merged count
                per_type[typ].append({"selectivity_nominal": round(sel, 4), "selectivity_measured": cnt / (rows * world), "ms_per_step": ms,
                                      "kernel_ms": kms / max(1, kn), "rows_per_sec": rows * world / (ms * 1e-3),
                                      "gbs": width * rows * world / (ms * 1e-3) / 1e9,
                                      "kernel_frac_of_peak": width * rows / (kms / max(1, kn) * 1e-3) / 1e9 / peak, "sql": sql})
                del plan
            worst = min(p["kernel_frac_of_peak"] for v in per_type.values() for p in v)
            out["c2"] = {"config": "filter-only scan selectivity sweep, COUNT(*) + SUM(v)", "rows_per_gpu": rows, "n_gpus": world,
                         "algorithmic_bytes_per_row": {"INT64": 16, "DOUBLE": 16, "STRING": 12, "DATE32": 12},
                         "sweep": per_type, "roofline": {"bound": "hbm", "kernel": "bq::k_scan (global aggregate)", "worst_kernel_frac": worst,
                                                         "peak": peak, "unit": "GB/s"}}
            del eng
        finally:
            drop(tabs)

    if not args.no_stress:
        section("c2", run_c2)

    # =================================================================================================================
    # C4: high-cardinality GROUP BY (100 M distinct keys over 2 B rows on 8 GPUs = 250 M rows per GPU), key-hash shuffle
    # =================================================================================================================
    def run_c4():
        n = max(1024, rows // 4)
        tabs = c4_tables(n * world, n, rank * n)
        try:
            if exchange:
                exchange.keep_sharded(True)
            eng = make_engine(tabs)
            plan = eng.plan(C4_SQL)
            before = exchange_counters()

            def step_dev():
                plan.run_device().free()
            ms_dev, launches, _, _ = timed(step_dev, 5, 2)
            res = {}

            def step_host():
                res["r"] = plan.run()
            ms_host, _, _, _ = timed(step_host, 2, 1)
            r = res["r"]
            groups = sum_over_ranks(r.rows)
            rows_seen = sum_over_ranks(int(r.cols[2].sum()))
            ids = max(16, n * world // 20)
            alg = 16 * n + 32 * (groups // world)
            rec = {"config": "C4 high-cardinality GROUP BY: SUM / COUNT / AVG per key", "n_gpus": world, "rows_per_gpu": n, "distinct_keys": groups,
                   "sql": C4_SQL, "ms_per_step": ms_dev, "rows_per_sec": n * world / (ms_dev * 1e-3),
                   "result": "groups left in HBM, each rank holding the keys it owns" if world > 1 else "groups left in HBM",
                   "ms_per_step_result_paged_to_host": ms_host, "result_bytes_to_host_per_gpu": int(sum(c.nbytes for c in r.cols)),
                   "gpu_launches_per_step": int(launches),
                   "roofline": {"bound": "hbm", "achieved": alg / (ms_dev * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                                "frac": alg / (ms_dev * 1e-3) / 1e9 / peak, "algorithmic_bytes_per_gpu_per_step": alg,
                                "what_bounds_it": "not HBM: after a 32-way hash partition (47 % of DRAM peak) the partition-major group table "
                                                  "lives in L2 and its reductions run at 88 % of the L2 slices' throughput "
                                                  "(profiles/ncu_r2_c4_kernels.json; shared-memory tables measured slower, profiles/README.md)"},
                   "invariants_at_full_size": "ok" if (rows_seen == n * world and ids * 0.99 < groups <= ids) else
                                              f"FAILED: rows {rows_seen} of {n * world}, groups {groups} of {ids}"}
            if world > 1:
                moved = 16 * n * (world - 1) // world
                rec["nvlink"] = {"bytes_stored_per_gpu_per_step": moved, "note": "rows are written straight into the owning rank's buffer by the partition kernel",
                                 "gbs_if_the_whole_step_were_the_shuffle": moved / (ms_dev * 1e-3) / 1e9, "peak": NVLINK_PEAK,
                                 "floor_ms": moved / NVLINK_PEAK / 1e6}
                rec["exchange"] = exchange_delta(before, 10)
            out["c4"] = rec
            del plan, eng, r, res
        finally:
            if exchange:
                exchange.keep_sharded(False)
            drop(tabs)

    if not args.no_stress:
        section("c4", run_c4)

    # =================================================================================================================
    # C5: Zipf(1.1) equi-join, 2 B probe x 500 M build rows on 8 GPUs = 250 M x 62.5 M per GPU
    # =================================================================================================================
    def run_c5():
        n, nb = max(1024, rows // 4), max(256, rows // 16)
        tabs = c5_tables(n * world, n, rank * n, nb * world, nb, rank * nb)
        try:
            eng = make_engine(tabs)
            plan = eng.plan(C5_SQL)
            res = {}

            def step():
                res["r"] = plan.run()
            before = exchange_counters()
            ms, launches, _, _ = timed(step, 5, 2)
            r = res["r"]
            alg = 16 * (n + nb)
            rec = {"config": "C5 skewed equi-join: probe key Zipf(1.1) over the whole build domain", "n_gpus": world, "probe_rows_per_gpu": n,
                   "build_rows_per_gpu": nb, "sql": C5_SQL, "ms_per_step": ms, "rows_per_sec": (n + nb) * world / (ms * 1e-3),
                   "gpu_launches_per_step": int(launches),
                   "roofline": {"bound": "hbm", "achieved": alg / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s", "frac": alg / (ms * 1e-3) / 1e9 / peak,
                                "algorithmic_bytes_per_gpu_per_step": alg},
                   "invariants_at_full_size": "ok" if int(r.cols[0][0]) == n * world else f"FAILED: COUNT(*) {int(r.cols[0][0])} of {n * world}"}
            if exchange:
                d = exchange_delta(before, 7)
                rec["exchange"] = d
                sent = d["collective_bytes_sent_per_gpu_per_step"]
                rec["nvlink"] = {"collective_bytes_sent_per_gpu_per_step": sent, "peak": NVLINK_PEAK,
                                 "note": "broadcast join: the build side is all-gathered; a co-partitioned join would move "
                                         f"{16 * (n + nb) * (world - 1) // world} bytes per GPU instead"}
            out["c5"] = rec
            del plan, eng
        finally:
            drop(tabs)

    if not args.no_stress:
        section("c5", run_c5)

    # =================================================================================================================
    # ORDER BY <double> DESC without LIMIT (BASELINE.md lists it as a measured reference row: 0.52 Mrows/s): rows / 10 rows
    # per GPU (10^8), result left in HBM.  N = 1: LSD radix sort of (64-bit key, row id) pairs, then one gather per column.
    # N > 1: sampled splitters, all-to-all by key range, local sort - rank r ends up holding the r-th range.
    # =================================================================================================================
    SORT_SQL = "SELECT k, v FROM s ORDER BY v DESC"

    def sort_tables(n_all, n_local, r0):
        k = ctx.alloc(bq.INT64, n_local).generate(dist=bq.GEN_UNIFORM, seed=SEED + 7, stream=0, lo=0, hi=(1 << 40), row0=r0)
        v = ctx.alloc(bq.DOUBLE, n_local).generate(dist=bq.GEN_UNIFORM_DIV, seed=SEED + 7, stream=1, lo=1, hi=(1 << 52), div=4096.0, row0=r0)
        ctx.sync()
        return {"s": ([("k", bq.INT64, k), ("v", bq.DOUBLE, v)], {})}

    def run_sort():
        n = max(4096, rows // 10)
        tabs = sort_tables(n * world, n, rank * n)
        try:
            eng = make_engine(tabs)
            plan = eng.plan(SORT_SQL)
            keep = {}

            def step():
                if "r" in keep:
                    keep["r"].free()
                keep["r"] = plan.run_device()
            before = exchange_counters()
            ms, launches, kern, _ = timed(step, 5, 2, profile=True)
            rel = keep["r"]
            mine = rel.rows
            total = sum_over_ranks(mine)
            # invariants at full size: every row is somewhere, each rank's rows descend, and the ranks' ranges descend too
            head = rel.col(1).to_numpy(0, min(mine, 1 << 20)) if mine else np.empty(0)
            tail = rel.col(1).to_numpy(max(0, mine - (1 << 20)), min(mine, 1 << 20)) if mine else np.empty(0)
            ok = total == n * world and bool(np.all(head[:-1] >= head[1:])) and bool(np.all(tail[:-1] >= tail[1:]))
            if world > 1:
                ends = torch.tensor([float(head[0]) if mine else float("inf"), float(tail[-1]) if mine else float("inf")], dtype=torch.float64, device="cuda")
                allends = [torch.empty_like(ends) for _ in range(world)]
                dist.all_gather(allends, ends)
                e = [t.tolist() for t in allends]
                ok = ok and all(e[r][1] >= e[r + 1][0] for r in range(world - 1))
            kn, kms = kern
            passes = kn // 5 if kn else 0
            alg = (16 + 16) * n + passes * 24 * n          # keys made + columns gathered once, 24 B per row per radix pass
            rec = {"config": "ORDER BY a DOUBLE column DESC, no LIMIT; result left in HBM", "n_gpus": world, "rows_per_gpu": n, "sql": SORT_SQL,
                   "ms_per_step": ms, "rows_per_sec": n * world / (ms * 1e-3), "gpu_launches_per_step": int(launches),
                   "radix_passes_per_step": passes, "radix_ms_per_step": kms / 5 if kn else None,
                   "roofline": {"bound": "hbm", "achieved": alg / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s", "frac": alg / (ms * 1e-3) / 1e9 / peak,
                                "algorithmic_bytes_per_gpu_per_step": alg,
                                "note": "24 B per row per 8-bit pass (key + row id read and written) + 32 B per row to make the keys and gather two columns"},
                   "rows_after_the_exchange_on_rank_0": mine,
                   "invariants_at_full_size": "ok" if ok else f"FAILED: {total} of {n * world} rows, or an order violation"}
            if exchange:
                rec["exchange"] = exchange_delta(before, 7)
            out["sort"] = rec
            keep["r"].free()
            del plan, eng
        finally:
            drop(tabs)

    if not args.no_stress:
        section("sort", run_sort)

    # =================================================================================================================
    # Parity on samples, at EVERY N: the same statements over small tables of the same generator, sharded over the ranks
    # and run through the same (NCCL) path; rank 0 regenerates the whole small table on the host, runs the compiled
    # reference on it and compares.  At N = 1 the reference's timings are the cpu_baseline of each configuration.
    # =================================================================================================================
    def run_parity():
        have_ref = False
        if rank == 0:
            from oracle import datagen as odg, ref_engine
            from tests.parity import assert_same_rows
            have_ref = ref_engine.available()
        S = int(min(rows, args.ref_rows))

        def shard(n):
            lo = rank * n // world
            return lo, (rank + 1) * n // world - lo

        def compare(name, sqls, gpu_tables, host_tables_fn, dicts, sample_desc, sample_rows, ordered=None, env=None):
            """gpu_tables: this rank's shard in HBM; host_tables_fn(): the whole sample as numpy (rank 0 only)."""
            rec = {"sample": sample_desc}
            old_env = {}
            for k, v in (env or {}).items():
                old_env[k] = os.environ.get(k)
                os.environ[k] = v
            try:
                eng = make_engine(gpu_tables, dicts=dicts)
                got = [eng.query(sql) for sql in sqls]
                del eng
            finally:
                for k, v in old_env.items():
                    if v is None:
                        os.environ.pop(k, None)
                    else:
                        os.environ[k] = v
                drop(gpu_tables)
            if rank == 0:
                if not have_ref:
                    rec["parity_on_sample"] = "not checked: oracle/_ref was not built"
                else:
                    reng = ref_engine.RefEngine()
                    d = reng.new_dict(dicts if dicts is not None else odg.STATUS_DICT)
                    for tname, cols in host_tables_fn().items():
                        reng.add_table(tname, cols, d)
                    secs = 0.0
                    for sql, g in zip(sqls, got):
                        w = reng.query(sql)
                        secs += w.seconds
                        assert_same_rows(g.cols, w.cols, ordered_by=ordered, what=f"bench {name} sample vs reference: {sql[:60]}")
                    rec["parity_on_sample"] = (f"ok: {len(sqls)} statement(s), N = {world}, integers / keys / order exact, DOUBLE SUM / AVG within 1e-12, "
                                               "against the compiled reference (oracle/_ref)")
                    rec["reference_seconds"] = secs
                    if world == 1:
                        rec["cpu_baseline"] = {"value": sample_rows * len(sqls) / secs, "unit": "rows/s", "cores": 1, "kind": "reference",
                                               "sample": sample_desc + f"; reference open..close {secs:.2f}s for {len(sqls)} statement(s)"}
            return rec

        results = {}
        # Q1
        lo, n = shard(S)
        results["q1"] = compare("Q1", [Q1_SQL], q1_tables(S, n, lo),
                                lambda: {"orders": [c for c in odg.host_table(odg.orders_schema(S), S, SEED) if c[0] != "order_id"]},
                                None, f"{S} orders rows of the same generator", S, ordered=[(0, True)])
        # Q2: fused probe, and the key-range passes forced by a tiny slice size
        SL, SO = max(1024, S // 2), max(256, S // 8)
        lo, n = shard(SL)
        olo, on = shard(SO)

        def q2_host():
            return {"orders": odg.host_table(odg.orders_schema(SO, prefix="o.")[:2], SO, SEED + 1),
                    "lineitem": odg.host_table(odg.lineitem_schema(SO, N_SKU), SL, SEED + 2)}
        results["q2"] = compare("Q2", [Q2_SQL], q2_tables(SL, n, lo, SO, on, olo), q2_host, None,
                                f"{SL} lineitem x {SO} orders rows of the same generator", SL + SO, ordered=[(1, False)])
        if rank == 0 and "cpu_baseline" in results["q2"]:
            results["q2"]["cpu_baseline"]["note"] = "rows = probe + build"
        r2 = compare("Q2 key-range passes", [Q2_SQL], q2_tables(SL, n, lo, SO, on, olo), q2_host, None, "same sample, bitmap cut into 16 KB slices",
                     SL + SO, ordered=[(1, False)], env={"BOSQL_BITMAP_SLICE_KB": "16"})
        results["q2"]["parity_on_sample_key_range_passes"] = r2.get("parity_on_sample")
        if world == 1:
            def q2h_host():
                return {"orders": odg.host_table(odg.orders_schema(SO, prefix="o.", key_stride=7919)[:2], SO, SEED + 1),
                        "lineitem": odg.host_table(odg.lineitem_schema(SO, N_SKU, key_stride=7919), SL, SEED + 2)}
            results["q2_hash"] = compare("Q2 hash join", [Q2_SQL], q2_tables(SL, n, lo, SO, on, olo, key_stride=7919), q2h_host, None,
                                         f"{SL} lineitem x {SO} orders rows, sparse order ids", SL + SO, ordered=[(1, False)])
        # C2: one statement per type and selectivity
        S2 = max(1024, S // 2)
        lo, n = shard(S2)
        results["c2"] = compare("C2", [q[2] for q in c2_queries()], c2_tables(S2, n, lo),
                                lambda: {"sweep": odg.host_table(odg.sweep_schema_c2(), S2, SEED + 3)}, datagen.C2_STR_DICT,
                                f"{S2} rows of the sweep table", S2)
        # C4: 3.3 M distinct keys, so that the N > 1 run takes the key-hash shuffle (more than 96 MB of group state)
        S4 = max(1024, min(S, 4_000_000))
        ids4 = max(16, S4 * 5 // 6)
        lo, n = shard(S4)

        def c4_host():
            return {"t": [("k", bq.INT64, odg.generate(bq.INT64, S4, odg.GEN_HASHED, SEED + 4, 0, lo=0, hi=ids4 - 1, modulus=1 << 61)),
                          ("v", bq.DOUBLE, odg.generate(bq.DOUBLE, S4, odg.GEN_UNIFORM_DIV, SEED + 4, 1, lo=1, hi=6400, div=64.0))]}
        results["c4"] = compare("C4", [C4_SQL], c4_tables(S4, n, lo, ids=ids4), c4_host, None, f"{S4} rows, up to {ids4} distinct keys", S4)
        # C5
        SP, SB = max(1024, S // 4), max(256, S // 16)
        lo, n = shard(SP)
        blo, bn = shard(SB)

        def c5_host():
            cdf, starts = odg.zipf_buckets(SB, 1.1)
            return {"build": [("b.k", bq.INT64, odg.generate(bq.INT64, SB, odg.GEN_SEQ, SEED + 5, 0, lo=1)),
                              ("b.w", bq.DOUBLE, odg.generate(bq.DOUBLE, SB, odg.GEN_UNIFORM_DIV, SEED + 5, 1, lo=1, hi=64, div=4.0))],
                    "probe": [("p.k", bq.INT64, odg.generate(bq.INT64, SP, odg.GEN_BUCKETS, SEED + 6, 0, lo=1, cdf=cdf, starts=starts)),
                              ("p.v", bq.DOUBLE, odg.generate(bq.DOUBLE, SP, odg.GEN_UNIFORM_DIV, SEED + 6, 1, lo=1, hi=64, div=4.0))]}
        results["c5"] = compare("C5", [C5_SQL], c5_tables(SP, n, lo, SB, bn, blo), c5_host, None, f"{SP} probe x {SB} build rows, Zipf(1.1) probe keys", SP + SB)
        if world > 1:
            r5 = compare("C5 co-partitioned", [C5_SQL], c5_tables(SP, n, lo, SB, bn, blo), c5_host, None, "same sample, join forced to the key-hash shuffle",
                         SP + SB, env={"BOSQL_JOIN": "shuffle"})
            results["c5"]["parity_on_sample_shuffle_join"] = r5.get("parity_on_sample")
        # ORDER BY DESC over a DOUBLE column (the whole sample ends up on every rank: it is below the range-sort threshold)
        SS = max(1024, min(S, 2_000_000))
        lo, n = shard(SS)

        def sort_host():
            return {"s": [("k", bq.INT64, odg.generate(bq.INT64, SS, odg.GEN_UNIFORM, SEED + 7, 0, lo=0, hi=(1 << 40))),
                          ("v", bq.DOUBLE, odg.generate(bq.DOUBLE, SS, odg.GEN_UNIFORM_DIV, SEED + 7, 1, lo=1, hi=(1 << 52), div=4096.0))]}
        results["sort"] = compare("ORDER BY", [SORT_SQL], sort_tables(SS, n, lo), sort_host, None, f"{SS} rows of the same generator", SS, ordered=[(1, False)])
        return results

    if not args.no_cpu:
        parity = {}

        def go():
            parity.update(run_parity())
        section("parity", go)
        if rank == 0:
            for key, rec in parity.items():
                target = out if key == "q1" else out.get(key)
                if target is None:
                    out[key] = target = {}
                for k, v in rec.items():
                    if key == "q1" and k == "sample":
                        continue
                    target[k] = v
            if "cpu_baseline" not in out and world == 1:
                out["cpu_baseline"] = {"value": None, "unit": "rows/s", "cores": 1, "kind": "reference", "sample": "oracle/_ref not built or the check failed"}
            if "cpu_baseline" in out and "parity_on_sample" in out:
                out["cpu_baseline"]["parity_on_sample"] = out["parity_on_sample"]

    if errors:
        out["errors"] = errors
    out["bench_wall_s"] = time.perf_counter() - t_start
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
    ap.add_argument("--no-pin", action="store_true", help="N > 1: do not give every rank its own share of the host cores")
    ap.add_argument("--no-numa", action="store_true", help="leave the host buffers of the end-to-end leg wherever the kernel puts them")
    ap.add_argument("--no-q2", action="store_true")
    ap.add_argument("--no-cpu", action="store_true", help="skip the parity samples / cpu_baseline (reference executor on rank 0)")
    ap.add_argument("--no-stress", action="store_true", help="skip the C2 / C4 / C5 configurations")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling variants (N > 1)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
