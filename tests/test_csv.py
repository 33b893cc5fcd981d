"""CPU: CSV ingest (bo-sql_b200/host/csv_ingest.cpp) against the reference's load_csv, cell for cell.

Golden facts from the reference's own tests/test_csv.cpp:7-54 first (types INT64/STRING/DOUBLE, min/max, NDV, dictionary ids
in first-seen order), then a differential run against the compiled reference on files that exercise every inference rule
(src/storage/csv_loader.cpp:48-161): the 8-character date rule and its range, integers parsed through double, prefix
parsing of std::stod ("12abc" is 12), empty lines, trailing commas, CR line ends, an empty table."""
import numpy as np
import pytest

INT64, DOUBLE, STRING, DATE32 = 0, 1, 2, 3


def write(tmp_path, name, text):
    p = tmp_path / name
    p.write_bytes(text.encode())
    return str(p)


def test_reference_csv_golden(bq, tmp_path):
    """tests/test_csv.cpp:7-54 of the reference, restated."""
    path = write(tmp_path, "t.csv", "id,name,value\n1,Alice,100.5\n2,Bob,200.0\n3,Charlie,300.75\n")
    eng = bq.Engine()
    eng.load_csv(path, "t")
    cols = eng.table_columns("t")
    assert [(c[0], c[1]) for c in cols] == [("id", INT64), ("name", STRING), ("value", DOUBLE)]
    assert (cols[0][3], cols[0][4]) == (1, 3)
    assert cols[1][5] == 3
    assert (cols[2][3], cols[2][4]) == (100.5, 300.75)
    assert eng.table_dict("t") == ["Alice", "Bob", "Charlie"]
    assert cols[1][2].tolist() == [0, 1, 2]


FILES = {
    "orders": "order_id,status,order_date,total\n" + "".join(
        f"{i + 1},{['COMPLETE', 'PENDING', 'CANCELLED', 'RETURNED'][(i * 7) % 4]},{20240000 + 100 * (1 + i % 12) + 1 + i % 28},{(i * 37 % 9000 + 100) / 100}\n"
        for i in range(500)),
    "dates_out_of_range": "d\n20240101\n18991231\n",
    "dates_wrong_width": "d\n20240101\n2024011\n",
    "ints_via_double": "a,b,c\n1,1.0,1e3\n-2,2.00,0x10\n9007199254740993,3,4\n",
    "prefix_parse": "a,b\n12abc,1.5x\n7,2\n",
    "mixed_becomes_string": "a,b\n1,x\n2.5,y\nz,x\n",
    "nan_inf": "a,b\nnan,1\ninf,2\n1,3\n",
    "empty_lines_and_trailing_comma": "a,b\n\n1,2\n\n3,4\n",
    "crlf": "a,b\r\n1,x\r\n2,y\r\n",
    "header_only": "a,b,c\n",
    "no_final_newline": "a,b\n1,2\n3,4",
    "empty_cells": "a,b,c\n,1,x\n2,,y\n",
    "negative_zero": "a\n-0.0\n0.0\n1.5\n",
    "big": "k,v,s,d\n" + "".join(f"{(i * 2654435761) % 1000003},{(i % 977) / 8},{'s' + str(i % 53)},{20150101 + (i % 9) * 10000}\n" for i in range(20000)),
}


def _numeric_stress():
    """One column per numeric spelling, 4000 rows each: every spelling the fast decimal path accepts, sits next to, or must
    hand to strtod (long digit strings, exponents, hex, signs, bare points, 16-17 significant digits, 2^53 neighbours)."""
    rng = np.random.default_rng(2024)
    n = 4000
    x = rng.uniform(-1e6, 1e6, n)
    big = rng.integers(2**52, 2**54, n)
    cols = {
        "f2": [f"{v:.2f}" for v in x],
        "f15": [f"{v:.15g}" for v in x],
        "f17": [f"{v:.17g}" for v in x],                                   # 17 significant digits: not the fast path
        "tiny": [f"{v * 1e-12:.22f}" for v in x],                          # many fractional digits
        "frac23": [f"0.{int(abs(v)) % 10}{'0' * 21}{int(abs(v)) % 7}" for v in x],   # 23 fractional digits
        "exp": [f"{v:.6e}" for v in x],
        "ints": [str(int(v)) for v in x],
        "big": [str(int(v)) for v in big],                                 # around 2^53: INT64 parsed through double
        "plus": [f"+{abs(v):.3f}" for v in x],
        "lead0": [f"000{abs(v):.4f}" for v in x],
        "point": [(f".{int(abs(v)) % 1000:03d}" if i % 2 else f"{int(abs(v))}.") for i, v in enumerate(x)],
        "negzero": ["-0" if i % 3 == 0 else ("-0.000" if i % 3 == 1 else "0") for i in range(n)],
        "sig16": [f"{int(abs(v) * 1e10) + 10**15}" for v in x],            # 16 digits
        "suffix": [f"{v:.2f}abc" if i % 5 == 0 else f"{v:.2f}" for i, v in enumerate(x)],
        "spaces": [f" {v:.2f}" if i % 4 == 0 else f"{v:.2f}" for i, v in enumerate(x)],
        "hex": ["0x1A" if i % 9 == 0 else str(i) for i in range(n)],
        "dates": [str(20240000 + 100 * (1 + i % 12) + 1 + i % 28) for i in range(n)],
        "dates_pad": [f"{i % 99999999:08d}" for i in range(n)],           # 8 characters but mostly outside the date range
    }
    names = list(cols)
    return ",".join(names) + "\n" + "".join(",".join(cols[k][i] for k in names) + "\n" for i in range(n))


FILES["numeric_stress"] = _numeric_stress()


@pytest.mark.parametrize("name", sorted(FILES))
def test_loader_matches_reference(bq, ref, tmp_path, name):
    path = write(tmp_path, name + ".csv", FILES[name])
    r = ref.RefEngine()
    g = bq.Engine()
    try:
        r.load_csv(path, "table")
    except RuntimeError as e:
        with pytest.raises(bq.BqError) as ei:
            g.load_csv(path, "table")
        assert str(ei.value) == str(e)
        return
    g.load_csv(path, "table")
    want, got = r.table_columns("table"), g.table_columns("table")
    assert len(want) == len(got)
    for w, x in zip(want, got):
        assert (w[0], w[1]) == (x[0], x[1]), f"{name}: column {w[0]} type {w[1]} vs {x[1]}"
        assert w[2].dtype == x[2].dtype and np.array_equal(w[2], x[2], equal_nan=w[2].dtype.kind == "f"), f"{name}: data of {w[0]}"
        if w[1] == DOUBLE:
            assert (w[3] == x[3] or (np.isnan(w[3]) and np.isnan(x[3]))) and (w[4] == x[4] or (np.isnan(w[4]) and np.isnan(x[4])))
        elif w[1] != STRING:
            assert (w[3], w[4]) == (x[3], x[4]), f"{name}: min/max of {w[0]}"
        if not (w[1] == DOUBLE and np.isnan(w[2]).any()):
            assert w[5] == x[5], f"{name}: ndv of {w[0]}"
    assert r.table_dict("table") == g.table_dict("table")


def test_row_size_mismatch_and_missing_file(bq, ref, tmp_path):
    path = write(tmp_path, "bad.csv", "a,b\n1,2\n3\n")
    for eng, exc in ((ref.RefEngine(), RuntimeError), (bq.Engine(), bq.BqError)):
        with pytest.raises(exc, match="Row size mismatch"):
            eng.load_csv(path, "table")
        with pytest.raises(exc, match="Cannot open file"):
            eng.load_csv(str(tmp_path / "nope.csv"), "table")
