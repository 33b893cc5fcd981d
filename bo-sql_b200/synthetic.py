"""bosql_b200.synthetic — the synthetic workload definitions of SURVEY.md section 8d (which columns, which distributions).

Pure descriptions: a schema is a list of (column name, type, generator spec) that `Column.generate` (bo-sql_b200/csrc/
bq_gen.cu) turns into HBM-resident columns; value(row) = f(seed, stream, global row), so any rank can generate any shard.
`oracle/datagen.py` restates the generator's arithmetic in numpy for the tests and re-exports these definitions; nothing here
imports the oracle.
"""
from __future__ import annotations

import numpy as np

GEN_SEQ, GEN_UNIFORM, GEN_UNIFORM_DIV, GEN_DATE, GEN_TABLE, GEN_HASHED, GEN_BUCKETS = range(7)
INT64, DOUBLE, STRING, DATE32 = 0, 1, 2, 3

STATUS_DICT = ["COMPLETE", "PENDING", "CANCELLED", "RETURNED"]     # ids 0..3 (first-seen order)


def orders_schema(n_orders, prefix="", div=100.0, key_stride=1):
    """o.order_id unique, 1 + row * key_stride (dense for stride 1; a large stride makes the domain sparse, so that a join
    on it needs a real hash table); o.status uniform over 4 ids; o.order_date 2024 days; o.total k/div."""
    p = prefix
    return [
        (p + "order_id", INT64, dict(dist=GEN_SEQ, lo=1, modulus=key_stride)),
        (p + "status", STRING, dict(dist=GEN_UNIFORM, lo=0, hi=3)),
        (p + "order_date", DATE32, dict(dist=GEN_DATE, base_year=2024, n_years=1)),
        (p + "total", DOUBLE, dict(dist=GEN_UNIFORM_DIV, lo=100, hi=100000, div=div)),
    ]


def lineitem_schema(n_orders, n_sku=100000, prefix="l.", div=100.0, sku_type=INT64, key_stride=1):
    p = prefix
    return [
        (p + "order_id", INT64, dict(dist=GEN_UNIFORM, lo=1, hi=n_orders, modulus=key_stride)),
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


C2_STR_DICT = [f"s{i}" for i in range(16)]


def sweep_schema_c2(div=100.0):
    """Configuration 2's table: like sweep_schema, but c_str is GEOMETRIC over 16 ids (id i with probability 2^-(i+1)), so that
    `c_str = 's<i>'` selects 50 %, 25 %, ... 1.6 % of the rows and `!=` the complements (STRING supports only = and !=)."""
    w = np.power(0.5, np.arange(1, 17, dtype=np.float64))
    c = np.cumsum(w)
    c /= c[-1]
    t = np.floor(c * float(1 << 53)).astype(np.uint64)
    t[-1] = np.uint64(1 << 53)
    schema = sweep_schema(div)
    schema[2] = ("c_str", STRING, dict(dist=GEN_TABLE, lo=0, cdf=t))
    return schema


def zipf_cdf(n_keys: int, s: float = 1.1) -> np.ndarray:
    """53-bit integer thresholds of a Zipf(s) distribution over n_keys ranks (for GEN_TABLE)."""
    w = 1.0 / np.power(np.arange(1, n_keys + 1, dtype=np.float64), s)
    c = np.cumsum(w)
    c /= c[-1]
    t = np.floor(c * float(1 << 53)).astype(np.uint64)
    t[-1] = np.uint64(1 << 53)
    return t


def zipf_buckets(n_keys: int, s: float = 1.1, head: int = 1 << 16, per_octave: int = 64):
    """Zipf(s) over ALL n_keys ranks for GEN_BUCKETS: (cdf thresholds, bucket starts).

    One bucket per rank for the first `head` ranks (exact masses k^-s); beyond, `per_octave` buckets per doubling of the rank,
    each carrying the mass of its ranks (midpoint rule on t^-s, relative error < 1e-9 at these ranks) and spreading it
    uniformly inside - the density changes by 2^(1/64) across a bucket, about 1 %.  Every key of the domain can be drawn."""
    head = min(head, n_keys)
    k = np.arange(1, head + 1, dtype=np.float64)
    mass = [1.0 / np.power(k, s)]
    starts = [np.arange(head, dtype=np.uint64)]
    if n_keys > head:
        edges = [head]
        x = float(head)
        step = 2.0 ** (1.0 / per_octave)
        while edges[-1] < n_keys:
            x *= step
            e = min(n_keys, max(edges[-1] + 1, int(x)))
            edges.append(e)
        e = np.asarray(edges, dtype=np.float64)          # bucket i covers ranks e[i]+1 .. e[i+1]  (offsets e[i] .. e[i+1]-1)
        a, b = e[:-1] + 0.5, e[1:] + 0.5
        mass.append((np.power(a, 1.0 - s) - np.power(b, 1.0 - s)) / (s - 1.0))
        starts.append(np.asarray(edges[:-1], dtype=np.uint64))
    m = np.concatenate(mass)
    c = np.cumsum(m)
    c /= c[-1]
    t = np.floor(c * float(1 << 53)).astype(np.uint64)
    t[-1] = np.uint64(1 << 53)
    st = np.concatenate(starts + [np.asarray([n_keys], dtype=np.uint64)])
    return t, st
