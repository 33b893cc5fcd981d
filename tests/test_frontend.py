"""CPU tests of the product's host layer: SQL front end and planner of libbosql_b200_exec.so (no GPU touched).

The plan text must equal what the reference's parser + LogicalPlanner print for the same statement — including the seven
golden strings of the reference's tests/test_logical.cpp:5-58 — and parse errors must carry the reference's messages.
"""
import ctypes as C

import pytest

from tests import golden_util as G

GOLD = G.load()


@pytest.mark.parametrize("entry", GOLD["explain"], ids=lambda e: e["sql"][:60])
def test_plan_text_matches_reference(bq, entry):
    L = bq.exec_lib()
    buf = C.create_string_buffer(8192)
    rc = L.bqx_explain(entry["sql"].encode(), 0, buf, 8192)
    if "error" in entry:
        assert rc != 0 and L.bqx_last_error().decode() == entry["error"]
    else:
        assert rc == 0, L.bqx_last_error().decode()
        assert buf.value.decode() == entry["plan"]


def test_reference_logical_golden_strings(bq):
    """tests/test_logical.cpp:5-58 verbatim."""
    L = bq.exec_lib()
    cases = {
        "SELECT a, b FROM t": "LogicalProject(a, b)\n  LogicalScan(table=t, cols=a, b)",
        "SELECT a FROM t WHERE b > 10": "LogicalProject(a)\n  LogicalFilter((b > 10))\n    LogicalScan(table=t, cols=a, b)",
        "SELECT a FROM t1 INNER JOIN t2 ON t1.id = t2.id":
            "LogicalProject(a)\n  LogicalHashJoin(left_keys=t1.id, right_keys=t2.id)\n    LogicalScan(table=t1, cols=a, t1.id, t2.id)\n    LogicalScan(table=t2, cols=a, t1.id, t2.id)",
        "SELECT SUM(a) FROM t GROUP BY b": "LogicalProject(SUM(a))\n  LogicalAggregate(keys=b, aggs=SUM(a))\n    LogicalScan(table=t, cols=a, b)",
        "SELECT a FROM t ORDER BY b DESC": "LogicalOrder(by: b DESC)\n  LogicalProject(a)\n    LogicalScan(table=t, cols=a, b)",
        "SELECT a FROM t LIMIT 5": "LogicalLimit(5)\n  LogicalProject(a)\n    LogicalScan(table=t, cols=a)",
        "SELECT sku, SUM(qty) FROM lineitem WHERE qty > 10 GROUP BY sku ORDER BY SUM(qty) DESC LIMIT 5":
            "LogicalLimit(5)\n  LogicalOrder(by: SUM(qty) DESC)\n    LogicalProject(sku, SUM(qty))\n      LogicalAggregate(keys=sku, aggs=SUM(qty))\n        LogicalFilter((qty > 10))\n          LogicalScan(table=lineitem, cols=qty, sku)",
    }
    for sql, want in cases.items():
        buf = C.create_string_buffer(8192)
        assert L.bqx_explain(sql.encode(), 0, buf, 8192) == 0
        assert buf.value.decode() == want


def test_front_end_extensions_are_opt_in(bq):
    """BETWEEN and decimal literals exist only behind parse flags; by default the reference's quirks stay."""
    L = bq.exec_lib()
    buf = C.create_string_buffer(8192)
    sql = b"SELECT a FROM t WHERE d BETWEEN 1 AND 2 GROUP BY a"
    assert L.bqx_explain(sql, 0, buf, 8192) == 0
    assert "LogicalAggregate" not in buf.value.decode()          # the statement silently ends at BETWEEN
    assert L.bqx_explain(sql, bq.PARSE_BETWEEN, buf, 8192) == 0
    assert "LogicalFilter(((d >= 1) AND (d <= 2)))" in buf.value.decode() and "LogicalAggregate(keys=a" in buf.value.decode()
    assert L.bqx_explain(b"SELECT a FROM t WHERE x > 1.5", bq.PARSE_DECIMALS, buf, 8192) == 0
    assert "(x > 1.5)" in buf.value.decode()


def test_negative_literals_and_keyword_case_are_opt_in(bq):
    """The rest of SURVEY.md 8f N4: -5 / -1.5 as literals and keywords in any case, behind parse flags; without the flags
    the reference's behaviour stays (its primary() rejects a leading '-', its keywords match exactly: parser.cpp:82-103,278-331)."""
    L = bq.exec_lib()
    buf = C.create_string_buffer(8192)

    def explain(sql, flags):
        rc = L.bqx_explain(sql.encode(), flags, buf, 8192)
        return rc, (buf.value.decode() if rc == 0 else L.bqx_last_error().decode())

    rc, text = explain("SELECT a FROM t WHERE x > -5", 0)
    assert rc != 0 and text == "Unexpected token in expression"
    rc, text = explain("SELECT a FROM t WHERE x > -5 AND y = 3 - -2", bq.PARSE_NEGATIVE)
    assert rc == 0 and "((x > -5) AND (y = (3 - -2)))" in text
    rc, text = explain("SELECT a FROM t WHERE x > -1.5", bq.PARSE_NEGATIVE | bq.PARSE_DECIMALS)
    assert rc == 0 and "(x > -1.5)" in text
    rc, text = explain("SELECT a - 1 FROM t", bq.PARSE_NEGATIVE)            # a binary minus is still a binary minus
    assert rc == 0 and "(a - 1)" in text
    rc, text = explain("select a from t where x > 1", 0)
    assert rc != 0 and text == "Expected 0 got 13"                          # "select" is an identifier to the reference
    rc, text = explain("select a, Sum(b) as s from t where x > 1 group by a order by s desc limit 3", bq.PARSE_ANY_CASE)
    want = explain("SELECT a, SUM(b) AS s FROM t WHERE x > 1 GROUP BY a ORDER BY s DESC LIMIT 3", 0)
    assert rc == 0 and want[0] == 0
    assert text == want[1]
    rc, text = explain("select a from t where d between 1 and 2", bq.PARSE_ANY_CASE | bq.PARSE_BETWEEN)
    assert rc == 0 and "((d >= 1) AND (d <= 2))" in text


def test_plan_construction_errors_need_no_gpu(bq):
    """Constructor-time validation (unknown table / column / join key) happens before any kernel could run."""
    eng = bq.Engine()
    eng.add_table("orders", [("orders.id", 0, [1, 2, 3]), ("orders.qty", 0, [10, 20, 30])])
    eng.add_table("detail", [("detail.id", 0, [1, 2, 4])])
    for sql, msg in [("SELECT x FROM nowhere", "Table not found: nowhere"),
                     ("SELECT nope FROM orders", "Unknown column: nope"),
                     ("SELECT orders.id FROM orders JOIN detail ON detail.id = orders.id", "Join key not found: detail.id")]:
        with pytest.raises(bq.BqError) as ei:
            eng.plan(sql)
        assert str(ei.value) == msg
    plan = eng.plan("SELECT COUNT(*) FROM orders")
    assert plan.root_kind == "HashAggregate" and plan.names == ["COUNT(*)"] and plan.types == [0]
    plan = eng.plan("SELECT orders.id, SUM(orders.qty) AS total, AVG(orders.qty) FROM orders GROUP BY orders.id ORDER BY total LIMIT 1")
    assert plan.root_kind == "Limit" and plan.names == ["orders.id", "total", "AVG(orders.qty)"] and plan.types == [0, 0, 1]


def test_reference_parser_facts(bq):
    """tests/test_parser.cpp of the reference, restated on the plan text: precedence (:27-37), aggregate calls (:39-53),
    the three statements that must throw (:87-96), a trailing ';' and JOIN aliases (:98-123, :144-157).  Where the compiled
    reference is present the text is also compared with its own."""
    from oracle import ref_engine
    L = bq.exec_lib()
    buf = C.create_string_buffer(8192)

    def explain(sql):
        rc = L.bqx_explain(sql.encode(), 0, buf, 8192)
        return (rc, buf.value.decode() if rc == 0 else L.bqx_last_error().decode())

    rc, text = explain("SELECT 1 + 2 * 3 FROM t")
    assert rc == 0 and "(1 + (2 * 3))" in text                       # MUL binds tighter than ADD
    rc, text = explain("SELECT SUM(price), COUNT(*) FROM products")
    assert rc == 0 and "SUM(price)" in text and "COUNT(*)" in text
    for bad in ("SELECT name", "SELECT @name FROM table", "SELECT name FROM table WHERE"):
        assert explain(bad)[0] != 0, bad
    rc, text = explain("SELECT a, b FROM t;")
    assert rc == 0 and text == "LogicalProject(a, b)\n  LogicalScan(table=t, cols=a, b)"
    rc, text = explain("SELECT x FROM orders o JOIN lineitem l ON o.id = l.id WHERE qty > 10;")
    assert rc == 0 and "LogicalHashJoin" in text and "(qty > 10)" in text
    if ref_engine.available():
        ref = ref_engine.RefEngine()
        for sql in ("SELECT 1 + 2 * 3 FROM t", "SELECT SUM(price), COUNT(*) FROM products", "SELECT a, b FROM t;",
                    "SELECT x FROM orders o JOIN lineitem l ON o.id = l.id WHERE qty > 10;",
                    "SELECT a - b - c, a / b * c, a + b > c AND d = 1 OR e != 2 FROM t"):
            assert explain(sql) == (0, ref.explain(sql)), sql
        for bad in ("SELECT name", "SELECT @name FROM table", "SELECT name FROM table WHERE"):
            with pytest.raises(Exception) as e:
                ref.explain(bad)
            assert explain(bad) == (1, str(e.value)) or explain(bad)[1] == str(e.value), bad
