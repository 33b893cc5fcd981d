"""GPU: the seeded random statements of tests/test_oracle_fuzz.py through the product's operator layer, against the compiled
reference (or the numpy oracle when oracle/_ref is absent): names, types and rows must agree, errors must carry the same text."""
import numpy as np
import pytest

from oracle import oracle as orc
from tests.golden import cases
from tests.parity import assert_same_rows
from tests.test_oracle_fuzz import _statement

pytestmark = pytest.mark.gpu


def test_random_statements_gpu_vs_reference(bq):
    from oracle import ref_engine
    tables = {name: (cols, dname) for name, cols, dname in cases.star_tables()}
    checker = ref_engine.RefEngine() if ref_engine.available() else orc.Oracle()
    g = bq.Engine()
    for eng in (checker, g):
        d = eng.new_dict(cases.DICTS["status"])
        for name in ("orders", "lineitem"):
            eng.add_table(name, tables[name][0], d)
    rng = np.random.default_rng(20240101)
    ran = errors = 0
    for i in range(300):
        sql, _ = _statement(rng)
        try:
            want = checker.query(sql)
        except Exception as e:  # noqa: BLE001 - RuntimeError (reference) or OracleError
            with pytest.raises(bq.BqError) as mine:
                g.query(sql)
            assert str(mine.value) == str(e), f"#{i} {sql}: GPU says {mine.value!r}, reference says {e!r}"
            errors += 1
            continue
        try:
            got = g.query(sql)
        except bq.BqError as e:
            raise AssertionError(f"#{i} {sql}: the reference answers, the GPU path raises {e}") from e
        assert got.names == want.names and got.types == want.types, f"#{i} {sql}"
        order = [(0, True), (1, True)] if " ORDER BY l.order_id, l.sku" in sql else None
        assert_same_rows(got.cols, want.cols, ordered_by=order, what=f"#{i} {sql}")
        ran += 1
    assert ran >= 200, (ran, errors)


def test_random_sweep_statements_gpu_vs_reference(bq):
    """The filter-sweep family: predicates over every type (dictionary ids, dates, DOUBLE against integer literals, truthiness,
    INT64-left against DOUBLE-right), GROUP BY dictionary / date / negative-integer keys, ORDER BY an aggregate alias."""
    from oracle import ref_engine
    from tests import golden_util as G
    from tests.test_oracle_fuzz import _sweep_statement
    (name, cols, dname), = cases.sweep_tables()
    checker = ref_engine.RefEngine() if ref_engine.available() else orc.Oracle()
    g = bq.Engine()
    for eng in (checker, g):
        eng.add_table(name, cols, eng.new_dict(cases.DICTS[dname]))
    rng = np.random.default_rng(7)
    ran = 0
    for i in range(300):
        sql = _sweep_statement(rng)
        try:
            want = checker.query(sql)
        except Exception as e:  # noqa: BLE001
            with pytest.raises(bq.BqError) as mine:
                g.query(sql)
            assert str(mine.value) == str(e), f"#{i} {sql}: GPU says {mine.value!r}, reference says {e!r}"
            continue
        try:
            got = g.query(sql)
        except bq.BqError as e:
            raise AssertionError(f"#{i} {sql}: the reference answers, the GPU path raises {e}") from e
        assert got.names == want.names and got.types == want.types, f"#{i} {sql}"
        assert_same_rows(got.cols, want.cols, ordered_by=G.order_spec(sql, want.names), what=f"#{i} {sql}")
        ran += 1
    assert ran >= 250, ran
