"""CPU tests: the numpy oracle (oracle/oracle.py) is pinned to the reference.

1. against the frozen answers of the COMPILED REFERENCE (tests/golden/ref_vectors.json) — this includes the
   reference's own nine execution tests (tests/test_execution.cpp:127-270) and every hazard query;
2. against the compiled reference live, where oracle/_ref exists (this container and the GPU box);
3. the generator restatement (oracle/datagen.py) against the frozen first values of every fixture column.
"""
import numpy as np
import pytest

from oracle import oracle as orc
from tests import golden_util as G
from tests.golden import cases
from tests.parity import assert_same_rows

GOLD = G.load()


@pytest.fixture(scope="module")
def oracles():
    return G.build_engines(orc.Oracle)


@pytest.mark.parametrize("entry", GOLD["queries"], ids=lambda e: e["sql"][:70])
def test_oracle_matches_golden(oracles, entry):
    eng = oracles[entry["tables"]]
    if "error" in entry:
        with pytest.raises(orc.OracleError) as ei:
            eng.query(entry["sql"])
        assert str(ei.value) == entry["error"]
        return
    got = eng.query(entry["sql"])
    want = G.decode(entry)
    assert got.names == entry["names"] and got.types == entry["types"]
    # the oracle adds in the reference's row order, so even DOUBLE sums are bit-identical: compare exactly
    order = G.order_spec(entry["sql"], entry["names"])
    assert_same_rows(got.cols, want, ordered_by=order, what=entry["sql"])
    g = [c for c in got.cols]
    exact = sorted(zip(*[c.tolist() for c in g])) == sorted(zip(*[c.tolist() for c in want])) if g and len(g[0]) else True
    assert exact, "oracle result is not bit-identical to the reference's"


def test_reference_own_nine_cases(oracles):
    """tests/test_execution.cpp:127-270, asserted the way the reference asserts them."""
    e = oracles["fixture"]
    q = e.query
    assert q("SELECT orders.id FROM orders WHERE orders.qty > 15").cols[0].tolist() == [2, 3]
    r = q("SELECT orders.id, orders.qty * 2 AS double_qty FROM orders")
    assert (r.cols[0].tolist(), r.cols[1].tolist()) == ([1, 2, 3], [20, 40, 60])
    assert q("SELECT orders.id FROM orders LIMIT 2").cols[0].tolist() == [1, 2]
    r = q("SELECT orders.id, detail.region FROM orders INNER JOIN detail ON orders.id = detail.id")
    assert r.dict is not None and [r.dict.strings[i] for i in r.cols[1]] == ["north", "south"] and r.cols[0].tolist() == [1, 2]
    r = q("SELECT detail.region, SUM(orders.qty) AS total FROM orders INNER JOIN detail ON orders.id = detail.id GROUP BY detail.region")
    assert r.names == ["detail.region", "total"]
    assert sorted((r.dict.strings[a], b) for a, b in zip(r.cols[0].tolist(), r.cols[1].tolist())) == [("north", 10), ("south", 20)]
    r = q("SELECT COUNT(*) FROM orders")
    assert r.names == ["COUNT(*)"] and r.cols[0].tolist() == [3]
    r = q("SELECT orders.id, orders.qty FROM orders ORDER BY orders.qty DESC")
    assert (r.cols[0][0], r.cols[1][0], r.cols[0][-1]) == (3, 30, 1)
    r = q("SELECT orders.id, orders.qty FROM orders ORDER BY orders.qty DESC LIMIT 1")
    assert (r.cols[0].tolist(), r.cols[1].tolist()) == ([3], [30])
    r = q("SELECT detail.region, SUM(orders.qty) AS total FROM orders INNER JOIN detail ON orders.id = detail.id "
          "GROUP BY detail.region ORDER BY total DESC LIMIT 1")
    assert r.dict.strings[r.cols[0][0]] == "south" and r.cols[1].tolist() == [20]


def test_oracle_matches_compiled_reference_live(ref):
    """Same statements on larger tables than the fixtures, against oracle/_ref itself."""
    from oracle import datagen
    n = 20_011
    tab = datagen.host_table(datagen.sweep_schema(), n, seed=77)
    o, r = orc.Oracle(), ref.RefEngine()
    o.add_table("t", tab)
    r.add_table("t", tab)
    for sql in ["SELECT COUNT(*), SUM(v), SUM(w), AVG(v) FROM t WHERE c_i64 < 400000 AND c_date >= 20190101",
                "SELECT c_str, COUNT(*), SUM(v), AVG(w) FROM t GROUP BY c_str",
                "SELECT c_i64, v FROM t WHERE c_f64 > 9000 ORDER BY c_i64 DESC LIMIT 50",
                "SELECT c_date, SUM(v * w) AS x FROM t GROUP BY c_date ORDER BY x DESC LIMIT 25"]:
        got, want = o.query(sql), r.query(sql)
        assert got.names == want.names and got.types == want.types
        assert_same_rows(got.cols, want.cols, ordered_by=G.order_spec(sql, want.names), what=sql)


def test_datagen_frozen_values():
    """oracle/datagen.py is the numpy restatement of the device generator; its output is frozen in the fixtures."""
    for tset, builder in cases.TABLE_SETS.items():
        for name, cols, _ in builder():
            for cname, typ, arr in cols:
                frozen = GOLD["tables"][tset][name][cname]
                head = np.asarray(arr)[:5]
                got = [float(x).hex() for x in head.tolist()] if head.dtype.kind == "f" else [int(x) for x in head.tolist()]
                assert got == frozen, f"{tset}.{name}.{cname}"


def test_oracle_is_pinned_on_the_sharded_suite_statements(ref):
    """tests/dist_sql.py checks the multi-GPU operator layer against the numpy oracle; here the oracle itself is checked
    against the compiled reference on exactly those statements and tables (joins with payload, top-k over joins, LIMIT
    without ORDER BY, multi-key GROUP BY, the skewed join)."""
    from oracle import datagen
    from tests import dist_sql
    orders, lines = dist_sql.tables()
    o, r = orc.Oracle(), ref.RefEngine()
    od, rd = o.new_dict(datagen.STATUS_DICT), r.new_dict(datagen.STATUS_DICT)
    for eng, d in ((o, od), (r, rd)):
        eng.add_table("orders", orders, d)
        eng.add_table("lineitem", lines, d)
    for name, sql, order in dist_sql.QUERIES:
        got, want = o.query(sql), r.query(sql)
        assert got.names == want.names and got.types == want.types, name
        assert_same_rows(got.cols, want.cols, ordered_by=None if order in (None, "sharded") else order, what=name)
