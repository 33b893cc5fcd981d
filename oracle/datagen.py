"""oracle/datagen.py — TEST INFRASTRUCTURE: numpy restatement of the synthetic-column generator.

The product generates benchmark tables directly in HBM (bo-sql_b200/csrc/bq_gen.cu).  So that any slice of a
table can be handed to the oracle, this file restates the same counter-based arithmetic in numpy:
value(row) = f(seed, stream, global_row) with a splitmix64-style hash.  tests/test_datagen.py compares the two
bit for bit on the GPU.  Table schemas follow SURVEY.md section 8d.
"""
from __future__ import annotations

import numpy as np

U64 = np.uint64
MASK = (1 << 64) - 1

GEN_SEQ, GEN_UNIFORM, GEN_UNIFORM_DIV, GEN_DATE, GEN_TABLE, GEN_HASHED, GEN_BUCKETS = range(7)
INT64, DOUBLE, STRING, DATE32 = 0, 1, 2, 3
NP_DTYPES = {INT64: np.int64, DOUBLE: np.float64, STRING: np.uint32, DATE32: np.int32}


def _mix64_arr(z):
    with np.errstate(over="ignore"):
        z = z + U64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> U64(30))) * U64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> U64(27))) * U64(0x94D049BB133111EB)
        return z ^ (z >> U64(31))


def _mix64_int(z: int) -> int:
    z = (z + 0x9E3779B97F4A7C15) & MASK
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & MASK
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & MASK
    return z ^ (z >> 31)


def row_hash(seed: int, stream: int, rows: np.ndarray) -> np.ndarray:
    """row_hash of bq_common.cuh: mix64(mix64(seed ^ stream*C) + row)."""
    base = _mix64_int((seed ^ ((stream * 0xD6E8FEB86659FD93) & MASK)) & MASK)
    with np.errstate(over="ignore"):
        return _mix64_arr(U64(base) + rows.astype(U64))


def generate(typ, n, dist, seed, stream, lo=0, hi=0, div=1.0, base_year=2024, n_years=1, cdf=None, modulus=0, row0=0, starts=None):
    """Rows [row0, row0+n) of a synthetic column, as the device kernel k_generate produces them."""
    rows = np.arange(row0, row0 + n, dtype=U64)
    h = row_hash(seed, stream, rows)
    rng = U64((hi - lo + 1) & MASK) if dist not in (GEN_SEQ, GEN_DATE, GEN_TABLE, GEN_BUCKETS) else U64(1)
    f = None
    stride = np.int64(modulus if modulus else 1)
    if dist == GEN_SEQ:
        v = np.int64(lo) + rows.astype(np.int64) * stride
    elif dist == GEN_UNIFORM:
        v = np.int64(lo) + (h % rng).astype(np.int64) * stride
    elif dist == GEN_UNIFORM_DIV:
        v = np.int64(lo) + (h % rng).astype(np.int64)
        f = v.astype(np.float64) / np.float64(div)
    elif dist == GEN_DATE:
        h2 = _mix64_arr(h)
        y = np.int64(base_year) + (h % U64(max(1, n_years))).astype(np.int64)
        m = 1 + (h2 % U64(12)).astype(np.int64)
        d = 1 + ((h2 >> U64(32)) % U64(28)).astype(np.int64)
        v = y * 10000 + m * 100 + d
    elif dist == GEN_TABLE:
        c = np.ascontiguousarray(cdf, dtype=U64)
        u = h >> U64(11)
        idx = np.searchsorted(c, u, side="right")      # first i with cdf[i] > u
        idx = np.minimum(idx, len(c) - 1)
        v = np.int64(lo) + idx.astype(np.int64)
    elif dist == GEN_BUCKETS:
        c = np.ascontiguousarray(cdf, dtype=U64)
        st = np.ascontiguousarray(starts, dtype=U64)
        u = h >> U64(11)
        idx = np.minimum(np.searchsorted(c, u, side="right"), len(c) - 1)
        width = st[idx + 1] - st[idx]
        v = np.int64(lo) + (st[idx] + _mix64_arr(h) % width).astype(np.int64)
    elif dist == GEN_HASHED:
        ident = h % rng
        with np.errstate(over="ignore"):
            salt = U64((seed * 0x2545F4914F6CDD1D) & MASK)
        v = np.int64(lo) + (_mix64_arr(ident ^ salt) % U64(modulus if modulus else 1)).astype(np.int64)
    else:
        raise ValueError("bad dist")
    if typ == DOUBLE:
        return f if f is not None else v.astype(np.float64)
    return v.astype(NP_DTYPES[typ])


# ---- table schemas: defined once, next to the product's generator binding --------------------------------
def _load_workload_definitions():
    """bo-sql_b200/synthetic.py loaded by path (the directory name has a hyphen, and the oracle must not pull the package's
    native libraries in just to read a few dictionaries)."""
    import importlib.util
    import os
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bo-sql_b200", "synthetic.py")
    spec = importlib.util.spec_from_file_location("_bosql_synthetic_definitions", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


_defs = _load_workload_definitions()
STATUS_DICT = _defs.STATUS_DICT
orders_schema = _defs.orders_schema
lineitem_schema = _defs.lineitem_schema
sweep_schema = _defs.sweep_schema
sweep_schema_c2 = _defs.sweep_schema_c2
C2_STR_DICT = _defs.C2_STR_DICT
zipf_cdf = _defs.zipf_cdf
zipf_buckets = _defs.zipf_buckets


def host_table(schema, n, seed, row0=0):
    """[(name, type, numpy array)] for rows [row0, row0+n) — stream id = column position."""
    return [(name, typ, generate(typ, n, seed=seed, stream=i, row0=row0, **spec)) for i, (name, typ, spec) in enumerate(schema)]
