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


def _rows(markdown):
    """cells of a markdown table (both command lines print one), header and rule dropped"""
    lines = [ln[ln.index("|"):] for ln in markdown.splitlines() if "|" in ln]      # the REPL prints its prompt (and colour codes) before the header
    return [[c.strip() for c in ln.strip().strip("|").split("|")] for ln in lines[2:]]


def test_cli_two_tables_join_like_the_reference_repl(tmp_path):
    """the README's Q2 through the command line.  The reference can only do this from its REPL
    (LOAD TABLE ... twice, src/cli/main.cpp:152-168); ours takes --table name=file.csv.  lineitem has no STRING column, so
    the reference's left-else-right dictionary choice (operator.cpp:694-704) is well defined and both must agree."""
    from oracle import datagen
    n_o, n_l = 3000, 20000
    o = {name: arr for name, _, arr in datagen.host_table(datagen.orders_schema(n_o, prefix="o."), n_o, seed=11)}
    li = {name: arr for name, _, arr in datagen.host_table(datagen.lineitem_schema(n_o, 50), n_l, seed=12)}
    op, lp = tmp_path / "orders.csv", tmp_path / "lineitem.csv"
    with open(op, "w") as f:
        f.write("o.order_id,o.status,o.total\n")
        for i in range(n_o):
            f.write(f"{int(o['o.order_id'][i])},{datagen.STATUS_DICT[int(o['o.status'][i])]},{float(o['o.total'][i])!r}\n")
    with open(lp, "w") as f:
        f.write("l.order_id,l.sku,l.qty,l.price\n")
        for i in range(n_l):
            f.write(f"{int(li['l.order_id'][i])},{int(li['l.sku'][i])},{int(li['l.qty'][i])},{float(li['l.price'][i])!r}\n")
    sql = ("SELECT l.sku, SUM(l.qty * l.price) AS rev FROM lineitem l JOIN orders o ON l.order_id = o.order_id "
           "WHERE o.status = 'COMPLETE' GROUP BY l.sku ORDER BY rev DESC LIMIT 20")
    r = subprocess.run([OURS, "--table", f"lineitem={lp}", "--table", f"orders={op}", "--sql", sql], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    ours = _rows(r.stdout)
    assert len(ours) == 20
    if not os.path.exists(REF):
        pytest.skip("oracle/_ref/bq_ref was not built")
    repl = f"LOAD TABLE lineitem FROM '{lp}'\nLOAD TABLE orders FROM '{op}'\n{sql}\nEXIT\n"
    q = subprocess.run([REF], input=repl, capture_output=True, text=True, timeout=120)
    ref = _rows(q.stdout)
    assert len(ref) == 20, q.stdout[-500:]
    for a, b in zip(ours, ref):
        assert a[0] == b[0] and abs(float(a[1]) - float(b[1])) <= 2e-6 + 1e-12 * abs(float(b[1])), (a, b)


def test_cli_shared_dictionary_and_extended_sql(tmp_path):
    """Both sides of the join carry STRING columns: with one dictionary for all tables of an invocation the right side's
    strings decode correctly (the reference would read them through the left table's dictionary).  --extended-sql turns on
    BETWEEN, decimal / negative literals and keywords in any case."""
    a, b = tmp_path / "a.csv", tmp_path / "b.csv"
    a.write_text("a.id,a.colour,a.v\n1,red,-1.5\n2,green,2.5\n3,blue,4.0\n")
    b.write_text("b.id,b.shape\n1,square\n2,circle\n3,square\n")
    sql = "select a.colour, b.shape, a.v from a join b on a.id = b.id where a.v between -2.0 and 2.5 order by a.v"
    r = subprocess.run([OURS, "--table", f"a={a}", "--table", f"b={b}", "--extended-sql", "--output-format", "csv", "--sql", sql],
                       capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    assert r.stdout.strip().splitlines() == ["a.colour,b.shape,a.v", "red,square,-1.500000", "green,circle,2.500000"], r.stdout
