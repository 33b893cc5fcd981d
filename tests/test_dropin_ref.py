"""The drop-in, demonstrated: the reference's OWN callers of the operator interface and its OWN test
sources, compiled unmodified where they lie under /root/reference against the product's headers and linked with
libbosql_b200_exec.so (tests/cpp/dropin_ref/Makefile; built by __graft_entry__.build() where /root/reference exists, the
binaries travel to the GPU box).  What is compiled from the reference: src/exec/physical_planner.cpp and
src/exec/execution.cpp (the only callers of Operator::open/next/close and of the operator constructors), its logical
planner, parser, formatter and CLI main; tests/test_*.cpp.  What is NOT: src/exec/operator.cpp, src/exec/expression.cpp,
src/storage/*, src/catalog/* - the part the product replaces.

  not gpu: the 23 cases that need no execution (types, columnar, catalog, CSV loader, parser, logical plans)
  gpu    : the nine execution cases of tests/test_execution.cpp:127-270 on the GPU operators, and the reference's CLI main
           (`bq_dropin`) against the reference's own binary on the same CSV
"""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "tests", "cpp", "dropin_ref", "build")
TESTS = os.path.join(BUILD, "dropin_ref_tests")
CLI = os.path.join(BUILD, "bq_dropin")
REF_CLI = os.path.join(ROOT, "oracle", "_ref", "bq_ref")


def _need(path):
    if not os.path.exists(path):
        pytest.skip(f"{os.path.relpath(path, ROOT)} was not built (needs /root/reference at build time: make -C tests/cpp/dropin_ref)")


def _cases(arg):
    _need(TESTS)
    r = subprocess.run([TESTS, arg], capture_output=True, text=True, timeout=300)
    lines = r.stdout.strip().splitlines()
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    return lines


def test_reference_unit_tests_compile_and_pass_on_the_product_headers():
    lines = _cases("~[exec]")
    assert lines[-1].startswith("23 test cases, 0 failed"), lines[-1]
    assert sum(1 for ln in lines if ln.startswith("ok ")) == 23


@pytest.mark.gpu
def test_reference_execution_tests_run_on_the_gpu_operators():
    lines = _cases("[exec]")
    assert lines[-1].startswith("9 test cases, 0 failed"), "\n".join(lines)
    want = ["Selection filters rows", "Projection evaluates expressions", "Limit short-circuits output", "Hash join produces matching rows",
            "Aggregate computes totals", "Global aggregate counts rows", "Order by sorts descending", "Order by with limit returns top row",
            "Top region by quantity"]
    ok = [ln for ln in lines if ln.startswith("ok ")]
    for name in want:
        assert any(name in ln for ln in ok), f"{name}: not among the passing cases: {ok}"


@pytest.mark.gpu
def test_reference_cli_main_runs_on_the_gpu_operators(tmp_path):
    """src/cli/main.cpp, compiled unmodified: load_csv -> parse_sql -> build_logical_plan -> build_physical_plan -> run_query."""
    _need(CLI)
    _need(REF_CLI)
    from oracle import datagen
    n = 5000
    cols = {name: arr for name, _, arr in datagen.host_table(datagen.orders_schema(n), n, seed=7)}
    path = tmp_path / "orders.csv"
    with open(path, "w") as f:
        f.write("order_id,status,total,order_date\n")
        for i in range(n):
            f.write(f"{int(cols['order_id'][i])},{datagen.STATUS_DICT[int(cols['status'][i])]},{float(cols['total'][i])!r},{int(cols['order_date'][i])}\n")
    for sql in ("SELECT order_date, SUM(total) AS revenue FROM table WHERE status = 'COMPLETE' AND order_date >= 20240101 AND order_date <= 20240131 "
                "GROUP BY order_date ORDER BY order_date",
                "SELECT order_id, total FROM table WHERE total > 990 ORDER BY total DESC LIMIT 7",
                "SELECT COUNT(*) FROM table"):
        outs = []
        for binary in (CLI, REF_CLI):
            r = subprocess.run([binary, str(path), "--sql", sql, "--output-format", "csv"], capture_output=True, text=True, timeout=120)
            assert r.returncode == 0, (binary, r.stdout[-500:], r.stderr[-500:])
            outs.append(r.stdout.strip().splitlines())
        got, want = outs
        assert len(got) == len(want), (got[:5], want[:5])
        for g, w in zip(got, want):
            if g == w:
                continue
            gc, wc = g.split(","), w.split(",")
            assert len(gc) == len(wc), (g, w)
            for a, b in zip(gc, wc):
                if a != b:
                    assert abs(float(a) - float(b)) <= 2e-6 + 1e-12 * abs(float(b)), (g, w)
