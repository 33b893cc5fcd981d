"""CPU: seeded random statements, numpy oracle vs the compiled reference (oracle/_ref).

The hand-written golden statements pin the oracle where someone thought to look; this pins it where nobody did: 300 statements
drawn from a small grammar over the star-schema tables (arithmetic over mixed types, comparisons with either side a column or a
literal, AND/OR nests, joins, GROUP BY one or two keys, SUM/COUNT/AVG over expressions, ORDER BY).  For each one both
engines must agree on output names, types and rows - or fail with the same message."""
import numpy as np
import pytest

from oracle import oracle as orc
from tests.golden import cases
from tests.parity import assert_same_rows

INT_COLS = ["l.order_id", "l.sku", "l.qty"]
F64_COLS = ["l.price"]
O_INT = ["o.order_id"]
O_OTHER = ["o.status", "o.order_date", "o.total"]


def _value(rng, depth=0):
    """An arithmetic expression over lineitem columns and literals."""
    r = rng.random()
    if depth >= 2 or r < 0.45:
        if rng.random() < 0.7:
            return str(rng.choice(INT_COLS + F64_COLS))
        return str(int(rng.integers(0, 60)))
    op = rng.choice(["+", "-", "*", "/"], p=[0.3, 0.3, 0.3, 0.1])
    right = _value(rng, depth + 1)
    if op == "/" and right.strip("()").isdigit() and int(right.strip("()")) == 0:
        right = "7"
    return f"({_value(rng, depth + 1)} {op} {right})"


def _pred(rng, joined, depth=0):
    r = rng.random()
    if depth < 2 and r < 0.35:
        return f"({_pred(rng, joined, depth + 1)} {rng.choice(['AND', 'OR'])} {_pred(rng, joined, depth + 1)})"
    kind = rng.integers(0, 6 if joined else 4)
    cmp_op = str(rng.choice(["<", "<=", ">", ">=", "=", "!="]))
    if kind == 0:
        return f"l.qty {cmp_op} {int(rng.integers(0, 55))}"
    if kind == 1:
        return f"l.price {cmp_op} {int(rng.integers(0, 110))}"
    if kind == 2:
        return f"{int(rng.integers(0, 40))} {cmp_op} l.sku"
    if kind == 3:
        return f"{_value(rng, 1)} {cmp_op} {_value(rng, 1)}"
    if kind == 4:
        return f"o.status {rng.choice(['=', '!='])} '{rng.choice(['COMPLETE', 'PENDING', 'RETURNED', 'nope'])}'"
    return f"o.order_date {cmp_op} {20240000 + 100 * int(rng.integers(1, 13)) + int(rng.integers(1, 29))}"


def _statement(rng):
    joined = rng.random() < 0.5
    src = "lineitem l JOIN orders o ON l.order_id = o.order_id" if joined else "lineitem l"
    where = f" WHERE {_pred(rng, joined)}" if rng.random() < 0.7 else ""
    shape = rng.integers(0, 3)
    if shape == 0:                                     # plain rows; no LIMIT, because (order_id, sku) has ties and the reference's
        cols = ["l.order_id", "l.sku"] + ([str(rng.choice(O_OTHER))] if joined else []) + [f"{_value(rng)} AS x"]   # sort is unstable (H4)
        return f"SELECT {', '.join(cols)} FROM {src}{where} ORDER BY l.order_id, l.sku", None
    aggs = []
    for i in range(int(rng.integers(1, 4))):
        f = str(rng.choice(["SUM", "COUNT", "AVG"]))
        aggs.append(f"COUNT(*) AS a{i}" if f == "COUNT" else f"{f}({_value(rng)}) AS a{i}")
    if shape == 1:                                     # global aggregate
        return f"SELECT {', '.join(aggs)} FROM {src}{where}", None
    keys = [str(rng.choice(["l.sku", "l.qty"] + (["o.status", "o.order_date"] if joined else [])))]
    if rng.random() < 0.3:
        extra = str(rng.choice(["l.qty", "l.sku"]))
        if extra not in keys:
            keys.append(extra)
    sql = f"SELECT {', '.join(keys + aggs)} FROM {src}{where} GROUP BY {', '.join(keys)}"
    return sql, None


def _sweep_pred(rng, depth=0):
    if depth < 2 and rng.random() < 0.35:
        return f"({_sweep_pred(rng, depth + 1)} {rng.choice(['AND', 'OR'])} {_sweep_pred(rng, depth + 1)})"
    cmp_op = str(rng.choice(["<", "<=", ">", ">=", "=", "!="]))
    kind = rng.integers(0, 9)
    if kind == 0:
        return f"c_i64 {cmp_op} {int(rng.integers(0, 1000000))}"
    if kind == 1:
        return f"c_f64 {cmp_op} {int(rng.integers(0, 10000))}"          # DOUBLE column against an integer literal
    if kind == 2:
        return f"c_date {cmp_op} {20150000 + 10000 * int(rng.integers(0, 10)) + 100 * int(rng.integers(1, 13)) + int(rng.integers(1, 29))}"
    if kind == 3:
        return f"c_str {rng.choice(['=', '!='])} {int(rng.integers(0, 100))}"
    if kind == 4:
        return f"s {rng.choice(['=', '!='])} '{rng.choice(['zero', 'one', 'two', 'five', 'unseen'])}'"
    if kind == 5:
        return str(rng.choice(["z", "s", "w"]))                             # truthiness of a bare column (H7: StrId 0 is false)
    if kind == 6:
        return f"{int(rng.integers(-60, 60))} {cmp_op} z"                   # literal on the left: compare dispatches on INT64
    if kind == 7:
        return f"c_i64 {cmp_op} c_f64"                                      # INT64-left vs DOUBLE-right truncation (H6)
    return f"(z * {int(rng.integers(1, 5))} + w) {cmp_op} (v / {int(rng.integers(1, 9))})"


def _sweep_statement(rng):
    """The filter-sweep table: predicates over every type, GROUP BY dictionary / date / negative-integer keys, ORDER BY an alias."""
    where = f" WHERE {_sweep_pred(rng)}" if rng.random() < 0.8 else ""
    shape = rng.integers(0, 3)
    if shape == 0:
        return f"SELECT COUNT(*) AS n, SUM(v) AS sv, SUM(w) AS sw, AVG(v * w) AS a FROM t{where}"
    if shape == 1:
        key = str(rng.choice(["c_str", "s", "z", "c_date"]))
        order = " ORDER BY " + str(rng.choice(["n", "sv", key])) + str(rng.choice(["", " DESC"])) if rng.random() < 0.6 else ""
        return f"SELECT {key}, COUNT(*) AS n, SUM(v) AS sv, AVG(w) AS aw FROM t{where} GROUP BY {key}{order}"
    return f"SELECT c_i64, z, s, v * 2 AS v2 FROM t{where} ORDER BY c_i64{rng.choice(['', ' DESC'])}, z"


def test_random_sweep_statements_oracle_vs_reference(ref):
    from tests import golden_util as G
    (name, cols, dname), = cases.sweep_tables()
    o, r = orc.Oracle(), ref.RefEngine()
    o.add_table(name, cols, o.new_dict(cases.DICTS[dname]))
    r.add_table(name, cols, r.new_dict(cases.DICTS[dname]))
    rng = np.random.default_rng(7)
    ran = 0
    for i in range(300):
        sql = _sweep_statement(rng)
        try:
            want = r.query(sql)
        except RuntimeError as e:
            with pytest.raises(Exception) as mine:
                o.query(sql)
            assert str(mine.value) == str(e), f"#{i} {sql}: oracle says {mine.value!r}, reference says {e!r}"
            continue
        got = o.query(sql)
        assert got.names == want.names and got.types == want.types, f"#{i} {sql}"
        assert_same_rows(got.cols, want.cols, ordered_by=G.order_spec(sql, want.names), what=f"#{i} {sql}")
        ran += 1
    assert ran >= 250, ran


def test_random_statements_oracle_vs_reference(ref):
    tables = {name: (cols, dname) for name, cols, dname in cases.star_tables()}
    o, r = orc.Oracle(), ref.RefEngine()
    od, rd = o.new_dict(cases.DICTS["status"]), r.new_dict(cases.DICTS["status"])
    for eng, d in ((o, od), (r, rd)):
        for name in ("orders", "lineitem"):
            eng.add_table(name, tables[name][0], d)
    rng = np.random.default_rng(20240101)
    ran = errors = 0
    for i in range(300):
        sql, _ = _statement(rng)
        try:
            want = r.query(sql)
        except RuntimeError as e:
            with pytest.raises(Exception) as mine:
                o.query(sql)
            assert str(mine.value) == str(e), f"#{i} {sql}: oracle says {mine.value!r}, reference says {e!r}"
            errors += 1
            continue
        got = o.query(sql)
        assert got.names == want.names and got.types == want.types, f"#{i} {sql}"
        order = [(0, True), (1, True)] if " ORDER BY l.order_id, l.sku" in sql else None
        assert_same_rows(got.cols, want.cols, ordered_by=order, what=f"#{i} {sql}")
        ran += 1
    assert ran >= 200, (ran, errors)             # most statements execute; the rest agree on the error text
