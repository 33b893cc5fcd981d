"""The fixed tables and statements behind tests/golden/ref_vectors.json (shared by the generator and the tests)."""
import numpy as np

from oracle import datagen

INT64, DOUBLE, STRING, DATE32 = 0, 1, 2, 3

DICTS = {
    "regions": ["north", "south", "west"],
    "words": ["zero", "one", "two", "three", "four", "five"],
    "status": datagen.STATUS_DICT,
}


def fixture_tables():
    """The reference's own test fixtures (tests/test_execution.cpp:13-63)."""
    return [
        ("orders", [("orders.id", INT64, np.array([1, 2, 3])), ("orders.qty", INT64, np.array([10, 20, 30]))], None),
        ("detail", [("detail.id", INT64, np.array([1, 2, 4])), ("detail.region", STRING, np.array([0, 1, 2], dtype=np.uint32))], "regions"),
    ]


def sweep_tables():
    n = 1500
    tab = datagen.host_table(datagen.sweep_schema(), n, seed=5)
    rng = np.random.default_rng(0)
    tab.append(("s", STRING, rng.integers(0, 6, size=n).astype(np.uint32)))
    tab.append(("z", INT64, rng.integers(-50, 50, size=n)))
    return [("t", tab, "words")]


def star_tables():
    n_orders, n_line, n_sku = 300, 2500, 40
    orders = datagen.host_table(datagen.orders_schema(n_orders, prefix="o."), n_orders, seed=1)
    line = datagen.host_table(datagen.lineitem_schema(n_orders, n_sku), n_line, seed=2)
    rng = np.random.default_rng(2)
    dup = [("d.k", INT64, rng.integers(1, 60, size=200)), ("d.w", DOUBLE, rng.integers(1, 64, size=200) / 4.0),
           ("d.tag", STRING, rng.integers(0, 4, size=200).astype(np.uint32))]
    probe = [("p.k", INT64, rng.integers(-5, 70, size=700)), ("p.v", DOUBLE, rng.integers(1, 64, size=700) / 4.0)]
    q1 = datagen.host_table(datagen.orders_schema(4000), 4000, seed=2024)
    return [("orders", orders, "status"), ("lineitem", line, "status"), ("dup", dup, "status"), ("probe", probe, "status"),
            ("q1orders", q1, "status")]


TABLE_SETS = {"fixture": fixture_tables, "sweep": sweep_tables, "star": star_tables}

_PREDS = [
    "c_i64 < 10000", "c_i64 >= 990000", "c_i64 != 697221", "c_f64 < 5000", "c_f64 >= 2500", "c_f64 != 344",
    "c_date >= 20180101 AND c_date <= 20181231", "c_date != 20170925", "c_str = 7", "c_str != 7", "s = 'two'", "s != 'zero'",
    "s = 'never-seen'", "c_i64 < 500000 AND c_f64 > 100 AND c_str != 3 AND c_date >= 20160101",
    "c_i64 < 100000 OR c_i64 > 900000", "(c_i64 < 100000 OR c_f64 > 9000) AND s = 'one'", "z", "s", "z AND s",
    "c_i64 < c_f64", "c_f64 < c_i64", "z * 2 + 1 > w / 10", "w / 7 = 3", "c_f64 / 0 > 5", "5 < z", "c_i64 < 0",
    "c_date >= 4294967296 + 20240101",
]

QUERIES = [
    # the reference's nine execution tests (tests/test_execution.cpp:127-270)
    ("fixture", "SELECT orders.id FROM orders WHERE orders.qty > 15"),
    ("fixture", "SELECT orders.id, orders.qty * 2 AS double_qty FROM orders"),
    ("fixture", "SELECT orders.id FROM orders LIMIT 2"),
    ("fixture", "SELECT orders.id, detail.region FROM orders INNER JOIN detail ON orders.id = detail.id"),
    ("fixture", "SELECT detail.region, SUM(orders.qty) AS total FROM orders INNER JOIN detail ON orders.id = detail.id GROUP BY detail.region"),
    ("fixture", "SELECT COUNT(*) FROM orders"),
    ("fixture", "SELECT orders.id, orders.qty FROM orders ORDER BY orders.qty DESC"),
    ("fixture", "SELECT orders.id, orders.qty FROM orders ORDER BY orders.qty DESC LIMIT 1"),
    ("fixture", "SELECT detail.region, SUM(orders.qty) AS total FROM orders INNER JOIN detail ON orders.id = detail.id GROUP BY detail.region ORDER BY total DESC LIMIT 1"),
    # errors and quirks
    ("fixture", "SELECT nope FROM orders"),
    ("fixture", "SELECT orders.id FROM orders WHERE missing > 1"),
    ("fixture", "SELECT orders.id FROM nowhere"),
    ("fixture", "SELECT orders.id FROM orders JOIN detail ON detail.id = orders.id"),
    ("fixture", "SELECT orders.id FROM orders WHERE orders.qty / 0 > 1"),
    ("fixture", "SELECT detail.id FROM detail WHERE detail.region < 'north'"),
    ("fixture", "SELECT orders.id FROM orders ORDER BY SUM(orders.qty)"),
    ("fixture", "SELECT orders.id FROM orders WHERE"),
    ("fixture", "SELECT FROM orders"),
    ("fixture", "SELECT orders.id, detail.id FROM orders JOIN detail ON orders.id < detail.id"),
    ("fixture", "SELECT COUNT(*), SUM(orders.qty) FROM orders JOIN detail ON orders.id < detail.id"),
    ("fixture", "SELECT COUNT(*) FROM orders WHERE orders.qty > 100"),
    ("fixture", "SELECT orders.id FROM orders WHERE orders.qty BETWEEN 10 AND 20 GROUP BY orders.id"),     # BETWEEN is not a keyword
] + [("sweep", f"SELECT COUNT(*), SUM(v), SUM(w), AVG(v), AVG(w) FROM t WHERE {p}") for p in _PREDS] + [
    ("sweep", f"SELECT c_i64, v, s, c_date FROM t WHERE {p}") for p in _PREDS[::4]] + [
    ("sweep", "SELECT SUM(v * w), SUM(w * w), SUM(v + v), SUM(w - 5), SUM(100 - w), SUM(v / w), SUM(w / 3) FROM t WHERE c_str != 50"),
    ("sweep", "SELECT SUM(c_str), SUM(c_date), AVG(c_date) FROM t WHERE c_i64 < 100000"),
    ("sweep", "SELECT s, COUNT(*), SUM(v) AS total, AVG(w) FROM t GROUP BY s"),
    ("sweep", "SELECT c_date, COUNT(*) FROM t GROUP BY c_date"),
    ("sweep", "SELECT z, SUM(w), AVG(v) FROM t GROUP BY z ORDER BY z"),
    ("sweep", "SELECT c_f64, COUNT(*) FROM t WHERE c_i64 < 50000 GROUP BY c_f64"),
    ("sweep", "SELECT s, z, COUNT(*), SUM(v) FROM t GROUP BY s, z"),
    ("sweep", "SELECT z + 1, COUNT(*) FROM t GROUP BY z + 1"),
    ("sweep", "SELECT s, SUM(v) AS total FROM t GROUP BY s ORDER BY total DESC LIMIT 3"),
    ("sweep", "SELECT c_str, SUM(w) AS sw, COUNT(*) AS n FROM t GROUP BY c_str ORDER BY n DESC, c_str LIMIT 10"),
    ("sweep", "SELECT COUNT(*), SUM(v) FROM t WHERE c_i64 < 0"),
    ("sweep", "SELECT c_i64, w * 2 + z AS e, v / 4, z < 0, c_date FROM t WHERE z >= 0"),
    ("sweep", "SELECT z, c_i64 FROM t ORDER BY z DESC, c_i64 LIMIT 100"),
    ("sweep", "SELECT c_i64, v FROM t LIMIT 7"),
    ("star", "SELECT order_date, SUM(total) AS revenue FROM q1orders WHERE status = 'COMPLETE' AND order_date >= 20240101 AND order_date <= 20240131 GROUP BY order_date ORDER BY order_date"),
    ("star", "SELECT l.sku, SUM(l.qty * l.price) AS rev FROM lineitem l JOIN orders o ON l.order_id = o.order_id WHERE o.status = 'COMPLETE' GROUP BY l.sku ORDER BY rev DESC LIMIT 20"),
    ("star", "SELECT COUNT(*), SUM(l.qty) FROM lineitem l JOIN orders o ON l.order_id = o.order_id"),
    ("star", "SELECT o.status, COUNT(*), SUM(l.price), AVG(o.total) FROM lineitem l JOIN orders o ON l.order_id = o.order_id GROUP BY o.status"),
    ("star", "SELECT COUNT(*) FROM lineitem l JOIN orders o ON l.order_id = o.order_id WHERE l.price > o.total"),
    ("star", "SELECT COUNT(*), SUM(p.v * d.w) FROM probe p JOIN dup d ON p.k = d.k"),
    ("star", "SELECT d.tag, COUNT(*), SUM(p.v) FROM probe p JOIN dup d ON p.k = d.k WHERE d.w > 4 GROUP BY d.tag"),
    ("star", "SELECT p.k, p.v, d.w, d.tag FROM probe p JOIN dup d ON p.k = d.k"),
    ("star", "SELECT COUNT(*) FROM probe p JOIN dup d ON p.k = d.w"),
    # round 2: any number of sort keys (operator.cpp:1115-1122), several GROUP BY expressions (:972-982), ORDER BY on the key
    # of a dense aggregate in both directions, a full ORDER BY over un-aggregated rows, LIMIT above Project / Selection
    ("sweep", "SELECT s, z, c_str, w, c_date, c_i64 FROM t ORDER BY s, z DESC, c_str, w DESC, c_date, c_i64 DESC"),
    ("sweep", "SELECT s, z, c_str, w, c_date FROM t WHERE w < 300 ORDER BY c_date DESC, w, c_str DESC, z, s DESC LIMIT 40"),
    ("sweep", "SELECT z + 1, w - 1, COUNT(*), SUM(v) FROM t WHERE w < 20 GROUP BY z + 1, w - 1"),
    ("sweep", "SELECT s, z * 2, COUNT(*) FROM t GROUP BY s, z * 2"),
    ("sweep", "SELECT c_date, SUM(v) AS total, COUNT(*) AS n FROM t GROUP BY c_date ORDER BY c_date"),
    ("sweep", "SELECT c_date, SUM(v) AS total FROM t GROUP BY c_date ORDER BY c_date DESC LIMIT 9"),
    ("sweep", "SELECT c_f64, c_i64 FROM t ORDER BY c_f64 DESC, c_i64"),
    ("sweep", "SELECT c_i64 / 7, v * 2 FROM t WHERE z > 10 LIMIT 12"),
    ("sweep", "SELECT c_i64 FROM t LIMIT 0"),
]

EXPLAIN = [
    # tests/test_logical.cpp:5-58 of the reference
    "SELECT a, b FROM t",
    "SELECT a FROM t WHERE b > 10",
    "SELECT a FROM t1 INNER JOIN t2 ON t1.id = t2.id",
    "SELECT SUM(a) FROM t GROUP BY b",
    "SELECT a FROM t ORDER BY b DESC",
    "SELECT a FROM t LIMIT 5",
    "SELECT sku, SUM(qty) FROM lineitem WHERE qty > 10 GROUP BY sku ORDER BY SUM(qty) DESC LIMIT 5",
    # more shapes and quirks
    "SELECT l.sku, SUM(l.qty * l.price) AS rev FROM lineitem l JOIN orders o ON l.order_id = o.order_id WHERE o.status = 'COMPLETE' GROUP BY l.sku ORDER BY rev DESC LIMIT 20",
    "SELECT order_date, SUM(total) AS revenue FROM orders WHERE status = 'COMPLETE' AND order_date BETWEEN 20240101 AND 20240131 GROUP BY order_date ORDER BY order_date",
    "SELECT COUNT(*) FROM orders",
    "SELECT * FROM t",
    "SELECT a + b * 2 - c / 4 FROM t WHERE (a < 1 OR b >= 2) AND c != 3",
    "SELECT a FROM t x JOIN u y ON x.k = y.k JOIN v z ON y.k = z.k",
    "SELECT a FROM t WHERE a = 'it''s'",
    "select a from t",
    "SELECT a FROM t WHERE a ! 5",
    "SELECT a FROM t WHERE a # 5",
    "SELECT a FROM t LIMIT x",
    "SELECT a AS FROM t",
    "SELECT a, FROM t",
    "SELECT AVG(a), COUNT(b), SUM(c) AS s, d FROM t GROUP BY d, e HAVING d > 1 ORDER BY s ASC, d DESC LIMIT 3;",
]
