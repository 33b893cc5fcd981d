"""oracle/ref_engine.py — TEST INFRASTRUCTURE, not product code.

ctypes front end to ``oracle/_ref/libbosql_ref.so``: the UNMODIFIED reference executor
(bolu-atx/bo-sql) compiled from /root/reference by ``oracle/Makefile`` behind the C shim
``oracle/ref_shim.cpp``.  It is the strongest oracle this repo has: the reference's own
parse_sql -> build_logical_plan -> build_physical_plan -> open/next/close
(src/cli/main.cpp:40-57) run on tables built from numpy arrays, returning raw typed columns.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / ``--impl reference`` legs
may import this module.  It never reads /root/reference at run time — only the prebuilt .so.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_ref", "libbosql_ref.so")
BQ_REF_CLI = os.path.join(HERE, "_ref", "bq_ref")

# TypeId ordinals, include/types.h:17
INT64, DOUBLE, STRING, DATE32 = 0, 1, 2, 3
NP_DTYPES = {INT64: np.int64, DOUBLE: np.float64, STRING: np.uint32, DATE32: np.int32}


def available() -> bool:
    return os.path.exists(LIB_PATH)


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError(f"{LIB_PATH} missing: run `make -C oracle` where /root/reference exists")
        L = C.CDLL(LIB_PATH)
        vp, cp, sz = C.c_void_p, C.c_char_p, C.c_size_t
        L.ref_last_error.restype = cp
        L.ref_dict_create.restype = vp
        L.ref_dict_destroy.argtypes = [vp]
        L.ref_dict_get_or_add.argtypes = [vp, cp]
        L.ref_dict_get_or_add.restype = C.c_uint
        L.ref_dict_size.argtypes = [vp]
        L.ref_dict_size.restype = sz
        L.ref_dict_get.argtypes = [vp, C.c_uint]
        L.ref_dict_get.restype = cp
        L.ref_catalog_create.restype = vp
        L.ref_catalog_destroy.argtypes = [vp]
        L.ref_table_create.argtypes = [cp, vp]
        L.ref_table_create.restype = vp
        L.ref_table_add_column.argtypes = [vp, cp, C.c_int, vp, sz]
        L.ref_catalog_register.argtypes = [vp, vp]
        L.ref_catalog_load_csv.argtypes = [vp, cp, cp]
        L.ref_catalog_table_info.argtypes = [vp, cp, C.POINTER(sz), C.POINTER(sz)]
        L.ref_catalog_column_info.argtypes = [
            vp, cp, sz, C.POINTER(cp), C.POINTER(C.c_int), C.POINTER(vp),
            C.POINTER(C.c_longlong), C.POINTER(C.c_longlong),
            C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(sz)]
        L.ref_catalog_dict_size.argtypes = [vp, cp]
        L.ref_catalog_dict_size.restype = sz
        L.ref_catalog_dict_get.argtypes = [vp, cp, C.c_uint]
        L.ref_catalog_dict_get.restype = cp
        L.ref_query.argtypes = [vp, cp]
        L.ref_query.restype = vp
        for f in ("ref_result_rows", "ref_result_cols"):
            getattr(L, f).argtypes = [vp]
            getattr(L, f).restype = sz
        L.ref_result_seconds.argtypes = [vp]
        L.ref_result_seconds.restype = C.c_double
        L.ref_result_name.argtypes = [vp, sz]
        L.ref_result_name.restype = cp
        L.ref_result_type.argtypes = [vp, sz]
        L.ref_result_data.argtypes = [vp, sz]
        L.ref_result_data.restype = vp
        L.ref_result_has_dict.argtypes = [vp]
        L.ref_result_dict_get.argtypes = [vp, C.c_uint]
        L.ref_result_dict_get.restype = cp
        L.ref_result_free.argtypes = [vp]
        L.ref_explain.argtypes = [cp, C.c_char_p, sz]
        _lib = L
    return _lib


@dataclass
class Result:
    """Typed query result: one numpy array per output column, in emit order."""
    names: list
    types: list
    cols: list
    rows: int
    seconds: float = 0.0
    has_dict: bool = False
    dict_strings: list = field(default_factory=list)

    def col(self, name):
        return self.cols[self.names.index(name)]


class RefDict:
    """A Dictionary (include/storage/dictionary.h:11) that several tables may share."""

    def __init__(self, strings=()):
        self.h = lib().ref_dict_create()
        for s in strings:
            self.get_or_add(s)

    def get_or_add(self, s: str) -> int:
        return lib().ref_dict_get_or_add(self.h, s.encode())

    def strings(self):
        L = lib()
        return [L.ref_dict_get(self.h, i).decode() for i in range(L.ref_dict_size(self.h))]


class RefEngine:
    """The compiled reference behind the same table/query surface as the product's Engine."""

    def __init__(self):
        self.L = lib()
        self.cat = self.L.ref_catalog_create()
        self._dicts = []

    def close(self):
        if self.cat:
            self.L.ref_catalog_destroy(self.cat)
            self.cat = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _err(self):
        return RuntimeError(self.L.ref_last_error().decode())

    def new_dict(self, strings=()):
        d = RefDict(strings)
        self._dicts.append(d)
        return d

    def add_table(self, name, columns, dictionary=None):
        """columns: list of (name, type_ordinal, numpy array); STRING columns carry uint32 ids."""
        t = self.L.ref_table_create(name.encode(), dictionary.h if dictionary else None)
        for cname, typ, arr in columns:
            a = np.ascontiguousarray(arr, dtype=NP_DTYPES[typ])
            if self.L.ref_table_add_column(t, cname.encode(), typ, a.ctypes.data_as(C.c_void_p), a.size):
                raise self._err()
        if self.L.ref_catalog_register(self.cat, t):
            raise self._err()

    def load_csv(self, path, name="table"):
        if self.L.ref_catalog_load_csv(self.cat, path.encode(), name.encode()):
            raise self._err()

    def table_columns(self, name):
        """(col_name, type, data copy, min, max, ndv) per column — what load_csv inferred."""
        rows, ncols = C.c_size_t(), C.c_size_t()
        if self.L.ref_catalog_table_info(self.cat, name.encode(), C.byref(rows), C.byref(ncols)):
            raise self._err()
        out = []
        for i in range(ncols.value):
            cn, ty, data = C.c_char_p(), C.c_int(), C.c_void_p()
            mi, ma = C.c_longlong(), C.c_longlong()
            mf, xf, ndv = C.c_double(), C.c_double(), C.c_size_t()
            if self.L.ref_catalog_column_info(self.cat, name.encode(), i, C.byref(cn), C.byref(ty), C.byref(data),
                                              C.byref(mi), C.byref(ma), C.byref(mf), C.byref(xf), C.byref(ndv)):
                raise self._err()
            dt = np.dtype(NP_DTYPES[ty.value])
            if rows.value:
                buf = (C.c_char * (dt.itemsize * rows.value)).from_address(data.value)
                arr = np.frombuffer(buf, dtype=dt).copy()
            else:
                arr = np.empty(0, dtype=dt)
            lo, hi = (mf.value, xf.value) if ty.value == DOUBLE else (mi.value, ma.value)
            out.append((cn.value.decode(), ty.value, arr, lo, hi, ndv.value))
        return out

    def table_dict(self, name):
        n = self.L.ref_catalog_dict_size(self.cat, name.encode())
        return [self.L.ref_catalog_dict_get(self.cat, name.encode(), i).decode() for i in range(n)]

    def query(self, sql: str) -> Result:
        r = self.L.ref_query(self.cat, sql.encode())
        if not r:
            raise self._err()
        try:
            rows = self.L.ref_result_rows(r)
            ncols = self.L.ref_result_cols(r)
            names, types, cols = [], [], []
            for i in range(ncols):
                names.append(self.L.ref_result_name(r, i).decode())
                t = self.L.ref_result_type(r, i)
                types.append(t)
                dt = np.dtype(NP_DTYPES[t])
                if rows:
                    buf = (C.c_char * (dt.itemsize * rows)).from_address(self.L.ref_result_data(r, i))
                    cols.append(np.frombuffer(buf, dtype=dt).copy())
                else:
                    cols.append(np.empty(0, dtype=dt))
            has_dict = bool(self.L.ref_result_has_dict(r))
            strings = []
            if has_dict:
                i = 0
                while True:
                    s = self.L.ref_result_dict_get(r, i)
                    if s is None:
                        break
                    strings.append(s.decode())
                    i += 1
            return Result(names, types, cols, rows, self.L.ref_result_seconds(r), has_dict, strings)
        finally:
            self.L.ref_result_free(r)

    def explain(self, sql: str) -> str:
        buf = C.create_string_buffer(8192)
        if self.L.ref_explain(sql.encode(), buf, 8192):
            raise self._err()
        return buf.value.decode()
