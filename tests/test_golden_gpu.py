"""GPU: the operator layer against the frozen answers of the compiled reference (tests/golden/ref_vectors.json).

Unlike tests/test_sql_gpu.py this needs no reference library at run time: the fixtures travel with the repository."""
import pytest

from tests import golden_util as G
from tests.parity import assert_same_rows

pytestmark = pytest.mark.gpu
GOLD = G.load()


@pytest.fixture(scope="module")
def engines(bq):
    return G.build_engines(bq.Engine)


@pytest.mark.parametrize("entry", GOLD["queries"], ids=lambda e: e["sql"][:70])
def test_gpu_matches_golden(bq, engines, entry):
    eng = engines[entry["tables"]]
    if "error" in entry:
        with pytest.raises(bq.BqError) as ei:
            eng.query(entry["sql"])
        assert str(ei.value) == entry["error"]
        return
    got = eng.query(entry["sql"])
    assert got.names == entry["names"] and got.types == entry["types"] and got.has_dict == entry["has_dict"]
    order = G.order_spec(entry["sql"], entry["names"])
    if order is None and " LIMIT " in entry["sql"] or ("GROUP BY" not in entry["sql"] and "COUNT(" not in entry["sql"] and "SUM(" not in entry["sql"]):
        # plain row streams keep scan / probe order: exact sequence
        import numpy as np
        for g, w in zip(got.cols, G.decode(entry)):
            assert g.dtype == w.dtype and np.array_equal(g, w), entry["sql"]
    else:
        assert_same_rows(got.cols, G.decode(entry), ordered_by=order, what=entry["sql"])
