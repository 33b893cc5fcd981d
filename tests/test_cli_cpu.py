"""CPU: the `bq_b200` command line where no kernel is involved - argument handling, CSV load errors, and a scan-only statement
(`SELECT * FROM table` pages the host columns straight out, as the reference's ColumnarScan does) - against the reference
binary `oracle/_ref/bq_ref` when it was built here, else against the expected text."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OURS = os.path.join(ROOT, "bo-sql_b200", "bq_b200")
REF = os.path.join(ROOT, "oracle", "_ref", "bq_ref")
ENV = dict(os.environ, CUDA_VISIBLE_DEVICES="")


def _run(binary, args, stdin_text=None):
    return subprocess.run([binary] + args, capture_output=True, text=True, timeout=60, env=ENV, input=stdin_text)


@pytest.fixture()
def csv(tmp_path):
    p = tmp_path / "people.csv"
    p.write_text("id,name,score,day\n1,Alice,3.5,20240105\n2,Bob,4.25,20240106\n3,Alice,5,20240107\n")
    return str(p)


@pytest.mark.parametrize("fmt", [None, "csv", "markdown", "CSV"])
def test_scan_only_statement_prints_like_the_reference(csv, fmt):
    args = [csv, "--sql", "SELECT * FROM table"] + (["--output-format", fmt] if fmt else [])
    ours = _run(OURS, args)
    assert ours.returncode == 0, ours.stderr
    if (fmt or "markdown").lower() == "csv":
        assert ours.stdout == "id,name,score,day\n1,Alice,3.500000,20240105\n2,Bob,4.250000,20240106\n3,Alice,5.000000,20240107\n"
    else:
        assert ours.stdout.splitlines()[0].replace(" ", "") == "|id|name|score|day|"
        assert "| 2  | Bob   | 4.250000 | 20240106 |" in ours.stdout
    if os.path.exists(REF):
        ref = _run(REF, args)
        assert ref.returncode == 0 and ours.stdout == ref.stdout


def test_stdin_csv(csv):
    text = open(csv).read()
    args = ["--sql", "SELECT * FROM table", "--output-format", "csv"]
    ours = _run(OURS, args, stdin_text=text)
    assert ours.returncode == 0 and ours.stdout.startswith("id,name,score,day\n1,Alice")
    if os.path.exists(REF):
        assert ours.stdout == _run(REF, args, stdin_text=text).stdout


@pytest.mark.parametrize("args", [["--bogus"], ["a.csv", "b.csv", "--sql", "SELECT 1"], ["a.csv", "--sql"],
                                  ["a.csv", "--sql", "SELECT * FROM table", "--output-format", "xml"],
                                  ["/nonexistent/file.csv", "--sql", "SELECT * FROM table"]])
def test_argument_and_load_errors_exit_1(args):
    ours = _run(OURS, args)
    assert ours.returncode == 1 and ours.stderr.strip()
    if os.path.exists(REF):
        assert _run(REF, args).returncode == 1


def test_statement_errors_are_reported_not_fatal(csv):
    """The reference prints the message and still exits 0 (src/cli/main.cpp:54-56); an aggregate needs the GPU, so here the
    product must say so loudly instead of computing anything on the host."""
    bad = _run(OURS, [csv, "--sql", "SELECT nope FROM table"])
    assert bad.returncode == 0 and "Unknown column: nope" in bad.stderr and bad.stdout == ""
    agg = _run(OURS, [csv, "--sql", "SELECT name, SUM(score) FROM table GROUP BY name"])
    assert agg.returncode == 0 and agg.stdout == "" and "CUDA" in agg.stderr
