"""bosql_b200 — ctypes bindings of the B200-native bo-sql operator hot path.

Two in-tree shared libraries (built by ``bo-sql_b200/Makefile``, see ``__graft_entry__.build``):

* ``libbosql_b200.so``      hand-written sm_100a kernels behind the kernel-layer C ABI (``include/bosql_b200.h``)
* ``libbosql_b200_exec.so`` the C++ mirror of the reference's operator interface
  (``include/exec/operator.hpp:17-218`` of bolu-atx/bo-sql) behind ``include/bosql_b200_exec.h``

This package is only a binding: it holds no arithmetic of its own and there is NO CPU fallback — importing it
without the built libraries raises, and creating a context without a CUDA device raises.

The directory name contains a hyphen (it is the name the project was given), so the package is loaded under
the module name ``bosql_b200`` by ``__graft_entry__.load_package()``.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
KERNEL_LIB = os.path.join(HERE, "libbosql_b200.so")
EXEC_LIB = os.path.join(HERE, "libbosql_b200_exec.so")

INT64, DOUBLE, STRING, DATE32 = 0, 1, 2, 3
NP_DTYPES = {INT64: np.int64, DOUBLE: np.float64, STRING: np.uint32, DATE32: np.int32}
TYPE_NAMES = {INT64: "INT64", DOUBLE: "DOUBLE", STRING: "STRING", DATE32: "DATE32"}

GEN_SEQ, GEN_UNIFORM, GEN_UNIFORM_DIV, GEN_DATE, GEN_TABLE, GEN_HASHED, GEN_BUCKETS = range(7)
V_NONE, V_A, V_B, V_MUL, V_ADD, V_SUB, V_DIV = range(7)
L_A, L_B, L_IMM = range(3)      # left operand of a binary aggregate argument
R_B, R_A, R_IMM = range(3)      # right operand (zero values = A op B)
GROUP_NONE, GROUP_DENSE, GROUP_HASH = range(3)
AGG_COUNT, AGG_SUM, AGG_AVG = range(3)
JOIN_AUTO, JOIN_BITMAP, JOIN_DIRECT, JOIN_HASH = range(4)

# opcodes of include/bosql_b200.h
OP = dict(COL=0, IMM_I=1, IMM_F=2, I2F=3, I2F_2=4, F2I=5, SX32=6, ZX32=7,
          ADD_I=8, SUB_I=9, MUL_I=10, DIV_I=11, ADD_F=12, SUB_F=13, MUL_F=14, DIV_F=15,
          EQ_I=16, NE_I=17, LT_I=18, LE_I=19, GT_I=20, GE_I=21,
          EQ_F=22, NE_F=23, LT_F=24, LE_F=25, GT_F=26, GE_F=27,
          TRUTHY_I=28, TRUTHY_F=29, TRUTHY_I_2=30, TRUTHY_F_2=31, AND=32, OR=33)


class BqError(RuntimeError):
    """Any failure reported through the C ABI (the reference throws std::runtime_error)."""


class GenSpec(C.Structure):
    _fields_ = [("dist", C.c_int), ("seed", C.c_uint64), ("stream", C.c_uint64), ("lo", C.c_int64), ("hi", C.c_int64),
                ("div", C.c_double), ("base_year", C.c_int32), ("n_years", C.c_int32), ("cdf", C.c_void_p),
                ("n_cdf", C.c_size_t), ("modulus", C.c_uint64), ("starts", C.c_void_p)]


class Range(C.Structure):
    _fields_ = [("lo", C.c_int64), ("hi", C.c_int64), ("neg", C.c_int32), ("pad", C.c_int32)]


class Slot(C.Structure):
    _fields_ = [("col", C.c_void_p), ("n_ranges", C.c_int32), ("from_build", C.c_int32), ("r", Range * 2)]


class VExpr(C.Structure):
    _fields_ = [("op", C.c_int32), ("l_src", C.c_int32), ("r_src", C.c_int32), ("imm_is_f", C.c_int32),
                ("imm_i", C.c_int64), ("imm_f", C.c_double)]


class AggOut(C.Structure):
    _fields_ = [("func", C.c_int32), ("v", C.c_int32), ("as_int", C.c_int32), ("pad", C.c_int32)]


class ScanSpec(C.Structure):
    _fields_ = [("key", Slot), ("a", Slot), ("b", Slot), ("pred", Slot * 3), ("jkey", Slot),
                ("row_begin", C.c_size_t), ("row_end", C.c_size_t), ("mask", C.c_void_p),
                ("n_v", C.c_int32), ("v", VExpr * 2), ("group_mode", C.c_int32),
                ("key_min", C.c_int64), ("key_max", C.c_int64), ("ndv_hint", C.c_size_t),
                ("join", C.c_void_p), ("row_bits", C.c_void_p), ("n_out", C.c_int32), ("out", AggOut * 8),
                ("hash_part_log2", C.c_int32), ("hash_part_shift", C.c_int32)]


class SelectSpec(C.Structure):
    _fields_ = [("pred", Slot * 4), ("mask", C.c_void_p), ("row_begin", C.c_size_t), ("row_end", C.c_size_t)]


class JoinSpec(C.Structure):
    _fields_ = [("key", C.c_void_p), ("pred", Slot * 3), ("mask", C.c_void_p),
                ("row_begin", C.c_size_t), ("row_end", C.c_size_t),
                ("kind", C.c_int32), ("need_rows", C.c_int32), ("unique", C.c_int32), ("pad", C.c_int32),
                ("key_min", C.c_int64), ("key_max", C.c_int64)]


class _Imm(C.Union):
    _fields_ = [("i", C.c_int64), ("f", C.c_double)]


class Insn(C.Structure):
    _fields_ = [("op", C.c_int32), ("arg", C.c_int32), ("imm", _Imm)]


_klib = None


def kernel_lib():
    """libbosql_b200.so, loaded once.  Missing library = hard error (no fallback of any kind)."""
    global _klib
    if _klib is not None:
        return _klib
    if not os.path.exists(KERNEL_LIB):
        raise BqError(f"{KERNEL_LIB} is missing: build it with `make -C bo-sql_b200` (or __graft_entry__.build()); "
                      "the hot path has no CPU fallback")
    L = C.CDLL(KERNEL_LIB, mode=C.RTLD_GLOBAL)
    vp, sz, i64 = C.c_void_p, C.c_size_t, C.c_int64
    P = C.POINTER
    sig = {
        "bq_ctx_create": ([C.c_int, P(vp)], C.c_int),
        "bq_ctx_destroy": ([vp], None),
        "bq_ctx_set_stream": ([vp, vp], C.c_int),
        "bq_ctx_sync": ([vp], C.c_int),
        "bq_ctx_stream": ([vp], vp),
        "bq_ctx_pool_stats": ([vp, P(sz), P(sz)], C.c_int),
        "bq_copy_bytes": ([vp, vp, vp, sz], C.c_int),
        "bq_zero_bytes": ([vp, vp, sz], C.c_int),
        "bq_ctx_info": ([vp, P(C.c_int), P(sz), P(sz)], C.c_int),
        "bq_ctx_launches": ([vp], C.c_uint64),
        "bq_ctx_profile": ([vp, C.c_int], C.c_int),
        "bq_ctx_profile_read": ([vp, P(C.c_uint64), P(C.c_double)], C.c_int),
        "bq_col_wrap": ([vp, C.c_int, vp, sz, P(vp)], C.c_int),
        "bq_last_error": ([], C.c_char_p),
        "bq_col_alloc": ([vp, C.c_int, sz, P(vp)], C.c_int),
        "bq_col_upload": ([vp, C.c_int, vp, sz, P(vp)], C.c_int),
        "bq_col_write": ([vp, vp, sz, vp, sz], C.c_int),
        "bq_col_read": ([vp, vp, sz, sz, vp], C.c_int),
        "bq_col_read_async": ([vp, vp, sz, sz, vp], C.c_int),
        "bq_col_owns": ([vp], C.c_int),
        "bq_col_free": ([vp, vp], None),
        "bq_col_size": ([vp], sz),
        "bq_col_type": ([vp], C.c_int),
        "bq_col_ptr": ([vp], vp),
        "bq_col_set_stats": ([vp, i64, i64, sz], C.c_int),
        "bq_col_invalidate_stats": ([vp], None),
        "bq_col_minmax": ([vp, vp, P(i64), P(i64)], C.c_int),
        "bq_f64_key": ([C.c_double], i64),
        "bq_f64_from_key": ([i64], C.c_double),
        "bq_host_alloc": ([sz, P(vp)], C.c_int),
        "bq_host_free": ([vp], None),
        "bq_col_generate": ([vp, vp, P(GenSpec), C.c_uint64], C.c_int),
        "bq_eval": ([vp, P(Insn), C.c_int, P(vp), C.c_int, sz, sz, C.c_int, P(vp)], C.c_int),
        "bq_scan_aggregate": ([vp, P(ScanSpec), P(vp)], C.c_int),
        "bq_scan_partial": ([vp, P(ScanSpec), P(vp)], C.c_int),
        "bq_agg_finish": ([vp, P(vp), C.c_int, C.c_int, C.c_int, P(AggOut), C.c_int, P(vp)], C.c_int),
        "bq_scan_state": ([vp, P(ScanSpec), P(vp)], C.c_int),
        "bq_agg_state_dense": ([vp, P(vp), P(sz)], C.c_int),
        "bq_agg_state_fold": ([vp, vp, vp, C.c_int], C.c_int),
        "bq_agg_state_emit": ([vp, vp, P(AggOut), C.c_int, P(vp)], C.c_int),
        "bq_agg_state_free": ([vp], None),
        "bq_partition": ([vp, vp, P(vp), C.c_int, sz, sz, C.c_int, C.c_int, P(vp), P(vp), P(vp)], C.c_int),
        "bq_group_tables_plan": ([sz, C.c_int, P(C.c_int), P(C.c_int)], C.c_int),
        "bq_partition_aggregate": ([vp, vp, P(vp), C.c_int, vp, C.c_int, C.c_int, P(AggOut), C.c_int, P(vp)], C.c_int),
        "bq_partition_count": ([vp, vp, sz, sz, C.c_int, C.c_int, P(i64), P(vp)], C.c_int),
        "bq_partition_count_hot": ([vp, vp, sz, sz, C.c_int, C.c_int, P(i64), C.c_int, P(i64), P(vp)], C.c_int),
        "bq_partition_scatter": ([vp, vp, P(vp), C.c_int, P(vp), P(vp), P(vp)], C.c_int),
        "bq_part_plan_free": ([vp], None),
        "bq_col_alloc_shared": ([vp, C.c_int, sz, P(vp)], C.c_int),
        "bq_col_ipc_export": ([vp, vp, vp], C.c_int),
        "bq_ipc_open": ([vp, vp, P(vp)], C.c_int),
        "bq_ctx_ipc_mappings": ([vp], sz),
        "bq_key_hash": ([i64], C.c_uint64),
        "bq_comm_unique_id": ([vp], C.c_int),
        "bq_comm_init": ([vp, C.c_int, C.c_int, vp], C.c_int),
        "bq_comm_destroy": ([vp], None),
        "bq_comm_world": ([vp], C.c_int),
        "bq_comm_rank": ([vp], C.c_int),
        "bq_comm_all_gather": ([vp, vp, vp, sz], C.c_int),
        "bq_comm_all_gather_v": ([vp, vp, vp, P(i64)], C.c_int),
        "bq_comm_all_to_all_v": ([vp, vp, P(i64), vp, P(i64)], C.c_int),
        "bq_comm_all_reduce_sum_u32": ([vp, vp, sz], C.c_int),
        "bq_comm_host_all_gather_i64": ([vp, P(i64), C.c_int32, P(i64)], C.c_int),
        "bq_comm_stats": ([vp, P(C.c_uint64), P(C.c_uint64)], C.c_int),
        "bq_select": ([vp, P(SelectSpec), P(vp)], C.c_int),
        "bq_gather": ([vp, vp, vp, P(vp)], C.c_int),
        "bq_slice": ([vp, vp, sz, sz, P(vp)], C.c_int),
        "bq_join_build": ([vp, P(JoinSpec), P(vp)], C.c_int),
        "bq_join_free": ([vp, vp], None),
        "bq_join_kind": ([vp], C.c_int),
        "bq_join_bytes": ([vp], sz),
        "bq_join_bitmap_ptr": ([vp, P(sz)], vp),
        "bq_join_build_rows": ([vp], sz),
        "bq_join_bitmap_popcount": ([vp, vp, P(C.c_uint64)], C.c_int),
        "bq_join_build_bitmap_nosync": ([vp, P(JoinSpec), P(vp)], C.c_int),
        "bq_join_bitmap_verdict": ([vp, vp, P(C.c_uint64), P(C.c_uint64), P(C.c_int)], C.c_int),
        "bq_join_probe": ([vp, vp, vp, vp, sz, sz, P(vp), P(vp)], C.c_int),
        "bq_join_probe_bits": ([vp, vp, vp, sz, sz, sz, P(vp)], C.c_int),
        "bq_rel_sort": ([vp, vp, C.c_int, P(C.c_int), P(C.c_int), i64, P(vp)], C.c_int),
        "bq_rel_create": ([vp, P(vp), C.c_int, P(vp)], C.c_int),
        "bq_rel_rows": ([vp], sz),
        "bq_rel_cols": ([vp], C.c_int),
        "bq_rel_col": ([vp, C.c_int], vp),
        "bq_rel_free": ([vp, vp], None),
        "bq_rel_release": ([vp, P(vp)], None),
    }
    for name, (args, res) in sig.items():
        fn = getattr(L, name)          # AttributeError here = the library does not export what the header declares
        fn.argtypes = args
        fn.restype = res
    L._bq_signatures = sig
    _klib = L
    return L


def _check(rc):
    if rc:
        raise BqError(kernel_lib().bq_last_error().decode())


def f64_key(v: float) -> int:
    return kernel_lib().bq_f64_key(float(v))


class Column:
    """A device-resident column (handle owned unless borrowed from a relation)."""

    def __init__(self, ctx, handle, owned=True):
        self.ctx, self.h, self.owned = ctx, handle, owned

    @property
    def n(self):
        return kernel_lib().bq_col_size(self.h)

    @property
    def type(self):
        return kernel_lib().bq_col_type(self.h)

    @property
    def ptr(self):
        return kernel_lib().bq_col_ptr(self.h)

    def to_numpy(self, offset=0, n=None):
        n = self.n - offset if n is None else n
        out = np.empty(n, dtype=NP_DTYPES[self.type])
        if n:
            _check(kernel_lib().bq_col_read(self.ctx.h, self.h, offset, n, out.ctypes.data_as(C.c_void_p)))
        return out

    def generate(self, dist, seed, stream, lo=0, hi=0, div=1.0, base_year=2024, n_years=1, cdf=None, modulus=0, row0=0, starts=None):
        s = GenSpec(dist=dist, seed=seed, stream=stream, lo=lo, hi=hi, div=div, base_year=base_year, n_years=n_years,
                    cdf=None, n_cdf=0, modulus=modulus, starts=None)
        keep = keep2 = None
        if cdf is not None:
            keep = np.ascontiguousarray(cdf, dtype=np.uint64)
            s.cdf = keep.ctypes.data_as(C.c_void_p)
            s.n_cdf = keep.size
        if starts is not None:
            keep2 = np.ascontiguousarray(starts, dtype=np.uint64)
            assert keep2.size == s.n_cdf + 1, "GEN_BUCKETS: one start per bucket plus the end"
            s.starts = keep2.ctypes.data_as(C.c_void_p)
        _check(kernel_lib().bq_col_generate(self.ctx.h, self.h, C.byref(s), row0))
        return self

    def set_stats(self, lo_key, hi_key, ndv=0):
        _check(kernel_lib().bq_col_set_stats(self.h, int(lo_key), int(hi_key), int(ndv)))
        return self

    def minmax(self):
        lo, hi = C.c_int64(), C.c_int64()
        _check(kernel_lib().bq_col_minmax(self.ctx.h, self.h, C.byref(lo), C.byref(hi)))
        return lo.value, hi.value

    def free(self):
        if self.h and self.owned:
            kernel_lib().bq_col_free(self.ctx.h, self.h)
        self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Relation:
    """A device-resident result relation; columns are borrowed views."""

    def __init__(self, ctx, handle):
        self.ctx, self.h = ctx, handle

    @property
    def rows(self):
        return kernel_lib().bq_rel_rows(self.h)

    @property
    def ncols(self):
        return kernel_lib().bq_rel_cols(self.h)

    def col(self, i):
        return Column(self.ctx, kernel_lib().bq_rel_col(self.h, i), owned=False)

    def to_numpy(self):
        return [self.col(i).to_numpy() for i in range(self.ncols)]

    def free(self):
        if self.h:
            kernel_lib().bq_rel_free(self.ctx.h, self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Join:
    def __init__(self, ctx, handle):
        self.ctx, self.h = ctx, handle

    @property
    def kind(self):
        return kernel_lib().bq_join_kind(self.h)

    @property
    def bytes(self):
        return kernel_lib().bq_join_bytes(self.h)

    def bitmap(self):
        n = C.c_size_t()
        p = kernel_lib().bq_join_bitmap_ptr(self.h, C.byref(n))
        return p, n.value

    @property
    def build_rows(self):
        return kernel_lib().bq_join_build_rows(self.h)

    def popcount(self):
        out = C.c_uint64(0)
        _check(kernel_lib().bq_join_bitmap_popcount(self.ctx.h, self.h, C.byref(out)))
        return int(out.value)

    def verdict(self):
        """(bits set, rows inserted, flags) of a bitmap built with join_build_bitmap_nosync (after any merge of the words)."""
        bits, ins, flags = C.c_uint64(0), C.c_uint64(0), C.c_int(0)
        _check(kernel_lib().bq_join_bitmap_verdict(self.ctx.h, self.h, C.byref(bits), C.byref(ins), C.byref(flags)))
        return int(bits.value), int(ins.value), int(flags.value)

    def probe_bits(self, probe_key, row_begin=0, row_end=None, slice_bytes=48 << 20):
        """Per-row match bits (uint32 column, bit i = row i) computed in key-range passes over an L2-sized bitmap slice."""
        out = C.c_void_p()
        end = probe_key.n if row_end is None else row_end
        _check(kernel_lib().bq_join_probe_bits(self.ctx.h, self.h, probe_key.h, row_begin, end, slice_bytes, C.byref(out)))
        return Column(self.ctx, out.value)

    def free(self):
        if self.h:
            kernel_lib().bq_join_free(self.ctx.h, self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def make_slot(col=None, ranges=(), from_build=False) -> Slot:
    """ranges: iterable of (lo_key, hi_key, neg) on the column's integer key (f64_key for DOUBLE)."""
    s = Slot()
    s.col = col.h if col is not None else None
    s.n_ranges = len(ranges)
    s.from_build = 1 if from_build else 0
    for i, (lo, hi, neg) in enumerate(ranges):
        s.r[i] = Range(int(lo), int(hi), int(bool(neg)), 0)
    return s


class Context:
    """One per process per GPU (bq_ctx)."""

    def __init__(self, device=0):
        self.L = kernel_lib()
        h = C.c_void_p()
        _check(self.L.bq_ctx_create(device, C.byref(h)))
        self.h = h
        self.device = device

    def close(self):
        if self.h:
            self.L.bq_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream_ptr):
        _check(self.L.bq_ctx_set_stream(self.h, C.c_void_p(cuda_stream_ptr) if cuda_stream_ptr else None))

    def sync(self):
        _check(self.L.bq_ctx_sync(self.h))

    def info(self):
        sm, fr, tot = C.c_int(), C.c_size_t(), C.c_size_t()
        _check(self.L.bq_ctx_info(self.h, C.byref(sm), C.byref(fr), C.byref(tot)))
        return sm.value, fr.value, tot.value

    @property
    def launches(self):
        return self.L.bq_ctx_launches(self.h)

    def profile(self, enable: bool):
        _check(self.L.bq_ctx_profile(self.h, int(enable)))

    def profile_read(self):
        """(launches, total_ms) of the fused scan kernel since the last read."""
        n, ms = C.c_uint64(), C.c_double()
        _check(self.L.bq_ctx_profile_read(self.h, C.byref(n), C.byref(ms)))
        return n.value, ms.value

    def wrap(self, typ, device_ptr, n) -> Column:
        """A non-owning Column over device memory managed elsewhere (e.g. a torch tensor)."""
        h = C.c_void_p()
        _check(self.L.bq_col_wrap(self.h, typ, C.c_void_p(device_ptr), n, C.byref(h)))
        return Column(self, h)

    # ---- columns
    def alloc(self, typ, n) -> Column:
        h = C.c_void_p()
        _check(self.L.bq_col_alloc(self.h, typ, n, C.byref(h)))
        return Column(self, h)

    def upload(self, typ, arr) -> Column:
        a = np.ascontiguousarray(arr, dtype=NP_DTYPES[typ])
        h = C.c_void_p()
        _check(self.L.bq_col_upload(self.h, typ, a.ctypes.data_as(C.c_void_p), a.size, C.byref(h)))
        return Column(self, h)

    def write(self, col, offset, host_ptr, n):
        _check(self.L.bq_col_write(self.h, col.h, offset, C.c_void_p(host_ptr), n))

    def host_alloc(self, nbytes):
        p = C.c_void_p()
        _check(self.L.bq_host_alloc(nbytes, C.byref(p)))
        return p.value

    def host_free(self, p):
        self.L.bq_host_free(C.c_void_p(p))

    # ---- pipelines
    def scan_aggregate(self, spec: ScanSpec, partial=False) -> Relation:
        h = C.c_void_p()
        fn = self.L.bq_scan_partial if partial else self.L.bq_scan_aggregate
        _check(fn(self.h, C.byref(spec), C.byref(h)))
        return Relation(self, h)

    def group_tables_plan(self, ndv_hint, n_args):
        """(log2_parts, splits) of the shared-memory-table GROUP BY for `ndv_hint` groups, or None when it does not apply."""
        lp, sp = C.c_int(), C.c_int()
        if not self.L.bq_group_tables_plan(int(ndv_hint), n_args, C.byref(lp), C.byref(sp)):
            return None
        return lp.value, sp.value

    def partition_aggregate(self, key, args, offsets, log2_parts, splits, outs) -> Relation:
        """GROUP BY over rows ordered by partition(): one shared-memory table per (partition, split)."""
        a = (C.c_void_p * max(1, len(args)))(*[c.h for c in args])
        o = (AggOut * max(1, len(outs)))(*outs)
        h = C.c_void_p()
        _check(self.L.bq_partition_aggregate(self.h, key.h, a, len(args), offsets.h, log2_parts, splits, o, len(outs), C.byref(h)))
        return Relation(self, h)

    def agg_finish(self, parts, has_key, key_type, outs) -> Relation:
        arr = (C.c_void_p * len(parts))(*[p.h for p in parts])
        o = (AggOut * len(outs))(*outs)
        h = C.c_void_p()
        _check(self.L.bq_agg_finish(self.h, arr, len(parts), int(has_key), key_type, o, len(outs), C.byref(h)))
        return Relation(self, h)

    def select(self, preds=(), mask=None, row_begin=0, row_end=0) -> Column:
        s = SelectSpec()
        for i, p in enumerate(preds):
            s.pred[i] = p
        s.mask = mask.h if mask is not None else None
        s.row_begin, s.row_end = row_begin, row_end
        h = C.c_void_p()
        _check(self.L.bq_select(self.h, C.byref(s), C.byref(h)))
        return Column(self, h)

    def partition(self, key, payload=(), log2_parts=8, hash_shift=None, row_begin=0, row_end=None):
        """Rows reordered by partition = (hash(key) >> hash_shift) & (2^log2_parts - 1).  Returns (key, [payload], offsets)."""
        hash_shift = 64 - log2_parts if hash_shift is None else hash_shift
        pay = (C.c_void_p * max(1, len(payload)))(*[c.h for c in payload])
        ok, off = C.c_void_p(), C.c_void_p()
        op = (C.c_void_p * 2)()
        _check(self.L.bq_partition(self.h, key.h, pay, len(payload), row_begin, key.n if row_end is None else row_end,
                                   log2_parts, hash_shift, C.byref(ok), op, C.byref(off)))
        return Column(self, ok), [Column(self, C.c_void_p(op[i])) for i in range(len(payload))], Column(self, off)

    def gather(self, col, rowids) -> Column:
        h = C.c_void_p()
        _check(self.L.bq_gather(self.h, col.h, rowids.h, C.byref(h)))
        return Column(self, h)

    def slice(self, col, begin, end) -> Column:
        h = C.c_void_p()
        _check(self.L.bq_slice(self.h, col.h, begin, end, C.byref(h)))
        return Column(self, h)

    def eval(self, prog, cols, row_begin, row_end, out_type) -> Column:
        """prog: list of (opname, arg, imm)"""
        arr = (Insn * len(prog))()
        for i, (op, arg, imm) in enumerate(prog):
            arr[i].op = OP[op] if isinstance(op, str) else op
            arr[i].arg = arg
            if isinstance(imm, float):
                arr[i].imm.f = imm
            else:
                arr[i].imm.i = int(imm)
        cs = (C.c_void_p * max(1, len(cols)))(*[c.h for c in cols])
        h = C.c_void_p()
        _check(self.L.bq_eval(self.h, arr, len(prog), cs, len(cols), row_begin, row_end, out_type, C.byref(h)))
        return Column(self, h)

    def join_build(self, key, preds=(), mask=None, row_begin=0, row_end=None, kind=JOIN_AUTO, need_rows=False,
                   unique=False, key_min=0, key_max=-1) -> Join:
        s = JoinSpec()
        s.key = key.h
        for i, p in enumerate(preds):
            s.pred[i] = p
        s.mask = mask.h if mask is not None else None
        s.row_begin = row_begin
        s.row_end = key.n if row_end is None else row_end
        s.kind, s.need_rows, s.unique = kind, int(need_rows), int(unique)
        s.key_min, s.key_max = key_min, key_max
        h = C.c_void_p()
        _check(self.L.bq_join_build(self.h, C.byref(s), C.byref(h)))
        return Join(self, h)

    def join_build_bitmap_nosync(self, key, preds=(), row_begin=0, row_end=None, key_min=0, key_max=-1) -> Join:
        """The bitmap build ranks use before they merge their bitmaps: nothing is read back; Join.verdict() does that."""
        s = JoinSpec()
        s.key = key.h
        for i, p in enumerate(preds):
            s.pred[i] = p
        s.row_begin = row_begin
        s.row_end = key.n if row_end is None else row_end
        s.kind = JOIN_BITMAP
        s.key_min, s.key_max = key_min, key_max
        h = C.c_void_p()
        _check(self.L.bq_join_build_bitmap_nosync(self.h, C.byref(s), C.byref(h)))
        return Join(self, h)

    def join_probe(self, join, probe_key, rowids=None, row_begin=0, row_end=None):
        a, b = C.c_void_p(), C.c_void_p()
        _check(self.L.bq_join_probe(self.h, join.h, probe_key.h, rowids.h if rowids is not None else None,
                                    row_begin, probe_key.n if row_end is None else row_end, C.byref(a), C.byref(b)))
        return Column(self, a), Column(self, b)

    def rel_create(self, cols) -> Relation:
        """Takes ownership of the columns."""
        arr = (C.c_void_p * len(cols))(*[c.h for c in cols])
        h = C.c_void_p()
        _check(self.L.bq_rel_create(self.h, arr, len(cols), C.byref(h)))
        for c in cols:
            c.owned = False
        return Relation(self, h)

    def rel_sort(self, rel, key_cols, asc, limit=-1) -> Relation:
        k = (C.c_int * max(1, len(key_cols)))(*key_cols)
        a = (C.c_int * max(1, len(asc)))(*[int(x) for x in asc])
        h = C.c_void_p()
        _check(self.L.bq_rel_sort(self.h, rel.h, len(key_cols), k, a, limit, C.byref(h)))
        return Relation(self, h)


def wrap_context(handle) -> "Context":
    """A non-owning Context over an existing bq_ctx (e.g. the operator layer's: engine.exec_lib().bqx_context())."""
    c = Context.__new__(Context)
    c.L = kernel_lib()
    c.h = C.c_void_p(handle)
    c.device = -1
    c.close = lambda: None
    return c


from .engine import PARSE_ANY_CASE, PARSE_BETWEEN, PARSE_DECIMALS, PARSE_NEGATIVE, Engine, exec_lib  # noqa: E402,F401
