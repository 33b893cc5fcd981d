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

GEN_SEQ, GEN_UNIFORM, GEN_UNIFORM_DIV, GEN_DATE, GEN_TABLE, GEN_HASHED = range(6)
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


def generate(typ, n, dist, seed, stream, lo=0, hi=0, div=1.0, base_year=2024, n_years=1, cdf=None, modulus=0, row0=0):
    """Rows [row0, row0+n) of a synthetic column, as the device kernel k_generate produces them."""
    rows = np.arange(row0, row0 + n, dtype=U64)
    h = row_hash(seed, stream, rows)
    rng = U64((hi - lo + 1) & MASK) if dist not in (GEN_SEQ, GEN_DATE, GEN_TABLE) else U64(1)
    f = None
    if dist == GEN_SEQ:
        v = np.int64(lo) + rows.astype(np.int64)
    elif dist == GEN_UNIFORM:
        v = np.int64(lo) + (h % rng).astype(np.int64)
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


# ---- table schemas -------------------------------------------------------------------
STATUS_DICT = ["COMPLETE", "PENDING", "CANCELLED", "RETURNED"]     # ids 0..3 (first-seen order)


def orders_schema(n_orders, prefix="", div=100.0):
    """o.order_id unique dense 1..N; o.status uniform over 4 ids; o.order_date 2024 days; o.total k/div."""
    p = prefix
    return [
        (p + "order_id", INT64, dict(dist=GEN_SEQ, lo=1)),
        (p + "status", STRING, dict(dist=GEN_UNIFORM, lo=0, hi=3)),
        (p + "order_date", DATE32, dict(dist=GEN_DATE, base_year=2024, n_years=1)),
        (p + "total", DOUBLE, dict(dist=GEN_UNIFORM_DIV, lo=100, hi=100000, div=div)),
    ]


def lineitem_schema(n_orders, n_sku=100000, prefix="l.", div=100.0, sku_type=INT64):
    p = prefix
    return [
        (p + "order_id", INT64, dict(dist=GEN_UNIFORM, lo=1, hi=n_orders)),
        (p + "sku", sku_type, dict(dist=GEN_UNIFORM, lo=0, hi=n_sku - 1)),
        (p + "qty", INT64, dict(dist=GEN_UNIFORM, lo=1, hi=50)),
        (p + "price", DOUBLE, dict(dist=GEN_UNIFORM_DIV, lo=100, hi=10000, div=div)),
    ]


def sweep_schema(div=100.0):
    """Filter-sweep table: one predicate column per type + v DOUBLE + w INT64 (sum of w stays below 2^53)."""
    return [
        ("c_i64", INT64, dict(dist=GEN_UNIFORM, lo=0, hi=999999)),
        ("c_f64", DOUBLE, dict(dist=GEN_UNIFORM_DIV, lo=0, hi=999999, div=div)),
        ("c_str", STRING, dict(dist=GEN_UNIFORM, lo=0, hi=99)),
        ("c_date", DATE32, dict(dist=GEN_DATE, base_year=2015, n_years=10)),
        ("v", DOUBLE, dict(dist=GEN_UNIFORM_DIV, lo=100, hi=100000, div=div)),
        ("w", INT64, dict(dist=GEN_UNIFORM, lo=1, hi=1000)),
    ]


def host_table(schema, n, seed, row0=0):
    """[(name, type, numpy array)] for rows [row0, row0+n) — stream id = column position."""
    return [(name, typ, generate(typ, n, seed=seed, stream=i, row0=row0, **spec)) for i, (name, typ, spec) in enumerate(schema)]


def zipf_cdf(n_keys: int, s: float = 1.1) -> np.ndarray:
    """53-bit integer thresholds of a Zipf(s) distribution over n_keys ranks (for GEN_TABLE)."""
    w = 1.0 / np.power(np.arange(1, n_keys + 1, dtype=np.float64), s)
    c = np.cumsum(w)
    c /= c[-1]
    t = np.floor(c * float(1 << 53)).astype(np.uint64)
    t[-1] = np.uint64(1 << 53)
    return t
