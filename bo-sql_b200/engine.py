"""bosql_b200.engine — ctypes front end of the operator layer (libbosql_b200_exec.so, include/bosql_b200_exec.h).

Same surface as oracle/ref_engine.RefEngine (tables from numpy arrays, SQL in, typed columns out), so a test drives the
reference executor and the GPU operators with one SQL string on identical tables.  This module only marshals: planning and
execution happen in the C++ operator layer and the CUDA kernels.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field

import numpy as np

from . import EXEC_LIB, NP_DTYPES, BqError, kernel_lib

PARSE_BETWEEN, PARSE_DECIMALS, PARSE_NEGATIVE, PARSE_ANY_CASE = 1, 2, 4, 8

_xlib = None


def exec_lib():
    global _xlib
    if _xlib is not None:
        return _xlib
    kernel_lib()        # libbosql_b200.so first (RTLD_GLOBAL), the operator layer links against it
    if not os.path.exists(EXEC_LIB):
        raise BqError(f"{EXEC_LIB} is missing: build it with `make -C bo-sql_b200`")
    L = C.CDLL(EXEC_LIB)
    vp, cp, sz = C.c_void_p, C.c_char_p, C.c_size_t
    P = C.POINTER
    sig = {
        "bqx_last_error": ([], cp),
        "bqx_init": ([C.c_int], C.c_int),
        "bqx_context": ([], vp),
        "bqx_dict_create": ([], vp),
        "bqx_dict_destroy": ([vp], None),
        "bqx_dict_get_or_add": ([vp, cp], C.c_uint32),
        "bqx_dict_size": ([vp], sz),
        "bqx_dict_get": ([vp, C.c_uint32], cp),
        "bqx_catalog_create": ([], vp),
        "bqx_catalog_destroy": ([vp], None),
        "bqx_table_create": ([cp, vp], vp),
        "bqx_table_add_column": ([vp, cp, C.c_int, vp, sz], C.c_int),
        "bqx_table_add_borrowed_column": ([vp, cp, C.c_int, vp, sz], C.c_int),
        "bqx_catalog_evict_device": ([vp, cp], C.c_int),
        "bqx_catalog_load_csv": ([vp, cp, cp], C.c_int),
        "bqx_catalog_table_info": ([vp, cp, P(sz), P(sz)], C.c_int),
        "bqx_catalog_column_info": ([vp, cp, sz, P(cp), P(C.c_int), P(vp), P(C.c_int64), P(C.c_int64), P(C.c_double),
                                    P(C.c_double), P(sz)], C.c_int),
        "bqx_catalog_dict_size": ([vp, cp], sz),
        "bqx_catalog_dict_get": ([vp, cp, C.c_uint32], cp),
        "bqx_table_add_device_column": ([vp, cp, vp, C.c_int], C.c_int),
        "bqx_table_set_stats": ([vp, cp, C.c_int64, C.c_int64, C.c_double, C.c_double, sz], C.c_int),
        "bqx_catalog_register": ([vp, vp], C.c_int),
        "bqx_plan_create": ([vp, cp, C.c_uint, P(vp)], C.c_int),
        "bqx_plan_destroy": ([vp], None),
        "bqx_plan_columns": ([vp], sz),
        "bqx_plan_column_name": ([vp, sz], cp),
        "bqx_plan_column_type": ([vp, sz], C.c_int),
        "bqx_plan_has_dict": ([vp], C.c_int),
        "bqx_plan_dict_get": ([vp, C.c_uint32], cp),
        "bqx_plan_root_kind": ([vp], cp),
        "bqx_plan_open": ([vp], C.c_int),
        "bqx_plan_next": ([vp, P(vp), sz, P(sz), P(C.c_int)], C.c_int),
        "bqx_plan_close": ([vp], C.c_int),
        "bqx_plan_run": ([vp, P(vp)], C.c_int),
        "bqx_plan_run_device": ([vp, P(vp)], C.c_int),
        "bqx_result_rows": ([vp], sz),
        "bqx_result_cols": ([vp], sz),
        "bqx_result_seconds": ([vp], C.c_double),
        "bqx_result_data": ([vp, sz], vp),
        "bqx_result_free": ([vp], None),
        "bqx_explain": ([cp, C.c_uint, C.c_char_p, sz], C.c_int),
        "bqx_set_exchange": ([vp], C.c_int),
        "bqx_comm_unique_id": ([vp], C.c_int),
        "bqx_comm_init": ([C.c_int, C.c_int, vp, C.c_int], C.c_int),
        "bqx_comm_init_file": ([cp, C.c_int, C.c_int, C.c_int], C.c_int),
        "bqx_comm_stats": ([P(C.c_uint64), P(C.c_uint64)], C.c_int),
        "bqx_exchange_keep_sharded": ([C.c_int], C.c_int),
    }
    for name, (args, res) in sig.items():
        fn = getattr(L, name)
        fn.argtypes = args
        fn.restype = res
    L._bqx_signatures = sig
    _xlib = L
    return L


def _check(rc):
    if rc:
        raise BqError(exec_lib().bqx_last_error().decode())


@dataclass
class Result:
    names: list
    types: list
    cols: list
    rows: int
    seconds: float = 0.0
    has_dict: bool = False
    dict_strings: list = field(default_factory=list)

    def col(self, name):
        return self.cols[self.names.index(name)]


class Dict:
    def __init__(self, strings=()):
        self.L = exec_lib()
        self.h = self.L.bqx_dict_create()
        for s in strings:
            self.get_or_add(s)

    def get_or_add(self, s: str) -> int:
        return self.L.bqx_dict_get_or_add(self.h, s.encode())

    def strings(self):
        return [self.L.bqx_dict_get(self.h, i).decode() for i in range(self.L.bqx_dict_size(self.h))]

    def __del__(self):
        try:
            self.L.bqx_dict_destroy(self.h)
        except Exception:
            pass


class _ResultOwner:
    """Frees a bqx_result when the last numpy view of it is gone."""

    def __init__(self, lib, handle):
        self.L, self.h = lib, handle

    def __del__(self):
        try:
            self.L.bqx_result_free(self.h)
        except Exception:
            pass


class Plan:
    """A planned statement: the root Operator behind open / next / close."""

    def __init__(self, engine, handle):
        self.L, self.h, self.engine = exec_lib(), handle, engine
        n = self.L.bqx_plan_columns(self.h)
        self.names = [self.L.bqx_plan_column_name(self.h, i).decode() for i in range(n)]
        self.types = [self.L.bqx_plan_column_type(self.h, i) for i in range(n)]
        self.root_kind = self.L.bqx_plan_root_kind(self.h).decode()

    def __del__(self):
        try:
            self.L.bqx_plan_destroy(self.h)
        except Exception:
            pass

    @property
    def has_dict(self):
        return bool(self.L.bqx_plan_has_dict(self.h))

    def dict_strings(self):
        out, i = [], 0
        while True:
            s = self.L.bqx_plan_dict_get(self.h, i)
            if s is None:
                return out
            out.append(s.decode())
            i += 1

    def open(self):
        _check(self.L.bqx_plan_open(self.h))

    def next(self):
        """One ExecBatch as a list of numpy arrays (copies), or None at end of stream."""
        n = len(self.types)
        ptrs = (C.c_void_p * max(1, n))()
        rows, eos = C.c_size_t(), C.c_int()
        _check(self.L.bqx_plan_next(self.h, ptrs, n, C.byref(rows), C.byref(eos)))
        if eos.value:
            return None
        out = []
        for i, t in enumerate(self.types):
            dt = np.dtype(NP_DTYPES[t])
            buf = (C.c_char * (dt.itemsize * rows.value)).from_address(ptrs[i])
            out.append(np.frombuffer(buf, dtype=dt).copy())
        return out

    def close(self):
        _check(self.L.bqx_plan_close(self.h))

    def run_device(self):
        """Runs the plan and leaves its output in HBM: a bosql_b200.Relation (kernel-layer handle) the caller owns."""
        from . import Relation, wrap_context
        rel = C.c_void_p()
        _check(self.L.bqx_plan_run_device(self.h, C.byref(rel)))
        return Relation(wrap_context(self.L.bqx_context()), rel)

    def run(self, copy=None) -> Result:
        """open .. next* .. close.  Small results are copied into numpy arrays; large ones (>= 1 MB per column, or copy=False)
        are numpy VIEWS of the C result, which is released when the last array goes away."""
        r = C.c_void_p()
        _check(self.L.bqx_plan_run(self.h, C.byref(r)))
        owner = _ResultOwner(self.L, r)
        rows = self.L.bqx_result_rows(r)
        cols = []
        for i, t in enumerate(self.types):
            dt = np.dtype(NP_DTYPES[t])
            if not rows:
                cols.append(np.empty(0, dtype=dt))
                continue
            nbytes = dt.itemsize * rows
            buf = (C.c_char * nbytes).from_address(self.L.bqx_result_data(r, i))
            view = np.frombuffer(buf, dtype=dt)
            if copy is True or (copy is None and nbytes < (1 << 20)):
                cols.append(view.copy())
            else:
                buf._owner = owner                # the ctypes array is the view's base: it keeps the C result alive
                cols.append(view)
        return Result(list(self.names), list(self.types), cols, rows, self.L.bqx_result_seconds(r), self.has_dict,
                      self.dict_strings() if self.has_dict else [])


class Engine:
    """Catalog + planner + GPU operators (the product), with RefEngine's surface."""

    def __init__(self, device=None):
        self.L = exec_lib()
        if device is not None:
            _check(self.L.bqx_init(device))
        self.cat = self.L.bqx_catalog_create()
        self._keep = []

    def close(self):
        if self.cat:
            self.L.bqx_catalog_destroy(self.cat)
            self.cat = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def new_dict(self, strings=()):
        d = Dict(strings)
        self._keep.append(d)
        return d

    def add_table(self, name, columns, dictionary=None, stats=None):
        """columns: list of (name, type, numpy array | bosql_b200.Column).  stats: {col: (min, max, ndv)}."""
        t = self.L.bqx_table_create(name.encode(), dictionary.h if dictionary else None)
        for cname, typ, data in columns:
            if hasattr(data, "h") and hasattr(data, "ctx"):        # device-resident column, kept alive by the caller
                self._keep.append(data)
                _check(self.L.bqx_table_add_device_column(t, cname.encode(), data.h, 0))
            elif isinstance(data, tuple):                          # (host pointer, rows): caller-owned (pinned) memory
                _check(self.L.bqx_table_add_borrowed_column(t, cname.encode(), typ, C.c_void_p(data[0]), data[1]))
            else:
                a = np.ascontiguousarray(data, dtype=NP_DTYPES[typ])
                _check(self.L.bqx_table_add_column(t, cname.encode(), typ, a.ctypes.data_as(C.c_void_p), a.size))
        for cname, (lo, hi, ndv) in (stats or {}).items():
            is_f = isinstance(lo, float)
            _check(self.L.bqx_table_set_stats(t, cname.encode(), 0 if is_f else int(lo), 0 if is_f else int(hi),
                                              float(lo) if is_f else 0.0, float(hi) if is_f else 0.0, int(ndv)))
        _check(self.L.bqx_catalog_register(self.cat, t))

    def load_csv(self, path: str, name: str = "table"):
        """load_csv + register (the reference CLI registers the file as "table")."""
        _check(self.L.bqx_catalog_load_csv(self.cat, path.encode(), name.encode()))

    def table_columns(self, name):
        """(col_name, type, data copy, min, max, ndv) per column — what the loader inferred."""
        rows, ncols = C.c_size_t(), C.c_size_t()
        _check(self.L.bqx_catalog_table_info(self.cat, name.encode(), C.byref(rows), C.byref(ncols)))
        out = []
        for i in range(ncols.value):
            cn, ty, data = C.c_char_p(), C.c_int(), C.c_void_p()
            mi, ma, mf, xf, ndv = C.c_int64(), C.c_int64(), C.c_double(), C.c_double(), C.c_size_t()
            _check(self.L.bqx_catalog_column_info(self.cat, name.encode(), i, C.byref(cn), C.byref(ty), C.byref(data), C.byref(mi),
                                                  C.byref(ma), C.byref(mf), C.byref(xf), C.byref(ndv)))
            dt = np.dtype(NP_DTYPES[ty.value])
            if rows.value:
                buf = (C.c_char * (dt.itemsize * rows.value)).from_address(data.value)
                arr = np.frombuffer(buf, dtype=dt).copy()
            else:
                arr = np.empty(0, dtype=dt)
            lo, hi = (mf.value, xf.value) if ty.value == 1 else (mi.value, ma.value)
            out.append((cn.value.decode(), ty.value, arr, lo, hi, ndv.value))
        return out

    def table_dict(self, name):
        n = self.L.bqx_catalog_dict_size(self.cat, name.encode())
        return [self.L.bqx_catalog_dict_get(self.cat, name.encode(), i).decode() for i in range(n)]

    def evict_device(self, table: str):
        """Drop the HBM mirrors of `table`'s host columns (they are uploaded again by the next query)."""
        _check(self.L.bqx_catalog_evict_device(self.cat, table.encode()))

    def plan(self, sql: str, parse_flags=0) -> Plan:
        h = C.c_void_p()
        _check(self.L.bqx_plan_create(self.cat, sql.encode(), parse_flags, C.byref(h)))
        return Plan(self, h)

    def query(self, sql: str, parse_flags=0) -> Result:
        return self.plan(sql, parse_flags).run()

    def explain(self, sql: str, parse_flags=0) -> str:
        buf = C.create_string_buffer(8192)
        _check(self.L.bqx_explain(sql.encode(), parse_flags, buf, 8192))
        return buf.value.decode()
