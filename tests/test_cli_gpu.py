"""Configuration 1 (BASELINE.json configs[0]): `bq <csv> --sql <query>` — our command line beside the reference's.

The reference binary is `oracle/_ref/bq_ref` (the unmodified sources compiled by oracle/Makefile); ours is
`bo-sql_b200/bq_b200` (bo-sql_b200/host/bq_cli.cpp).  Both read the same CSV and must print the same table
(src/cli/main.cpp:59-129, src/exec/execution.cpp:8-61, src/exec/formatter.cpp): text cells identical, DOUBLE cells
(six printed decimals) equal within 1e-6 absolute + 1e-12 relative since SUM order differs.
"""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OURS = os.path.join(ROOT, "bo-sql_b200", "bq_b200")
REF = os.path.join(ROOT, "oracle", "_ref", "bq_ref")

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def orders_csv(tmp_path_factory):
    from oracle import datagen
    n = 20000
    cols = {name: arr for name, _, arr in datagen.host_table(datagen.orders_schema(n), n, seed=7)}
    path = tmp_path_factory.mktemp("cli") / "orders.csv"
    status = datagen.STATUS_DICT
    with open(path, "w") as f:
        f.write("order_id,customer_id,status,total,order_date\n")
        for i in range(n):
            oid = int(cols["order_id"][i])
            f.write(f"{oid},{oid % 500},{status[int(cols['status'][i])]},{float(cols['total'][i])!r},{int(cols['order_date'][i])}\n")
    return str(path)


def _run(binary, csv, sql, fmt=None, stdin=None):
    if not os.path.exists(binary):
        if binary == REF:
            pytest.skip("oracle/_ref/bq_ref was not built (needs /root/reference at build time)")
        pytest.fail(f"{binary} is missing: run __graft_entry__.build()")
    cmd = [binary] + ([csv] if csv else []) + ["--sql", sql] + (["--output-format", fmt] if fmt else [])
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=120, stdin=stdin)
    return r.returncode, r.stdout, r.stderr


def _same(a, b):
    la, lb = a.strip().splitlines(), b.strip().splitlines()
    assert len(la) == len(lb), (a[:400], b[:400])
    for x, y in zip(la, lb):
        if x == y:
            continue
        sep = "|" if x.startswith("|") else ","
        cx, cy = [c.strip() for c in x.split(sep)], [c.strip() for c in y.split(sep)]
        assert len(cx) == len(cy), (x, y)
        for u, v in zip(cx, cy):
            if u != v:
                fu, fv = float(u), float(v)           # raises (fails) if a non-numeric cell differs
                assert abs(fu - fv) <= 2e-6 + 1e-12 * abs(fv), (x, y)


QUERIES = [
    "SELECT customer_id, SUM(total) AS revenue, COUNT(*) AS n FROM table WHERE order_date >= 20240301 AND order_date <= 20240930 GROUP BY customer_id ORDER BY customer_id LIMIT 50",
    "SELECT order_id, status, total FROM table WHERE total > 900 ORDER BY order_id LIMIT 25",
    "SELECT status, COUNT(*), AVG(total) FROM table GROUP BY status ORDER BY status",
    "SELECT order_id, total * 2 AS dbl FROM table WHERE customer_id = 7 ORDER BY order_id",
    "SELECT * FROM table WHERE order_id < 5 ORDER BY order_id",
    "SELECT order_id FROM table WHERE order_id < 0",
    "SELECT order_date, SUM(total) AS revenue FROM table WHERE status = 'COMPLETE' AND order_date >= 20240101 AND order_date <= 20240131 GROUP BY order_date ORDER BY order_date",
]


@pytest.mark.parametrize("sql", QUERIES)
@pytest.mark.parametrize("fmt", [None, "csv"])
def test_cli_matches_reference_binary(orders_csv, sql, fmt):
    rc_r, out_r, _ = _run(REF, orders_csv, sql, fmt)
    rc_o, out_o, err_o = _run(OURS, orders_csv, sql, fmt)
    assert rc_o == rc_r == 0, err_o
    _same(out_o, out_r)


def test_cli_reads_stdin(orders_csv):
    sql = "SELECT COUNT(*) AS n FROM table WHERE status = 'PENDING'"
    with open(orders_csv) as f:
        rc_r, out_r, _ = _run(REF, None, sql, "csv", stdin=f)
    with open(orders_csv) as f:
        rc_o, out_o, err_o = _run(OURS, None, sql, "csv", stdin=f)
    assert rc_o == rc_r == 0, err_o
    _same(out_o, out_r)


def test_cli_argument_errors(orders_csv):
    for args in (["--bogus"], [orders_csv, "extra.csv", "--sql", "SELECT 1"], [orders_csv, "--sql"],
                 [orders_csv, "--sql", "SELECT * FROM table", "--output-format", "xml"], ["/nonexistent.csv", "--sql", "SELECT * FROM table"]):
        r = subprocess.run([OURS] + args, capture_output=True, text=True, timeout=60)
        q = subprocess.run([REF] + args, capture_output=True, text=True, timeout=60)
        assert r.returncode == q.returncode == 1, (args, r.returncode, q.returncode)
