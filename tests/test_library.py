"""CPU tests: the C-ABI libraries load, export every symbol their headers declare, and refuse to run without a GPU."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared(header, prefix):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(" + prefix + r"[a-z0-9_]+)\s*\(", text)))


def exported(lib):
    out = subprocess.run(["nm", "-D", "--defined-only", lib], capture_output=True, text=True, check=True).stdout
    return {line.split()[-1] for line in out.splitlines() if " T " in line}


def test_kernel_library_exports_its_header(bq):
    names = declared("bosql_b200.h", "bq_")
    assert len(names) >= 40
    have = exported(bq.KERNEL_LIB)
    missing = [n for n in names if n not in have]
    assert not missing, f"declared in include/bosql_b200.h but not exported: {missing}"
    L = bq.kernel_lib()
    for n in names:
        assert hasattr(L, n)
    # and the binding knows every one of them
    unbound = [n for n in names if n not in L._bq_signatures]
    assert not unbound, f"exported but not bound in bo-sql_b200/__init__.py: {unbound}"


def test_operator_library_exports_its_header(bq):
    names = declared("bosql_b200_exec.h", "bqx_")
    assert len(names) >= 30
    have = exported(bq.EXEC_LIB)
    missing = [n for n in names if n not in have]
    assert not missing, f"declared in include/bosql_b200_exec.h but not exported: {missing}"
    L = bq.exec_lib()
    unbound = [n for n in names if n not in L._bqx_signatures]
    assert not unbound, f"exported but not bound in bo-sql_b200/engine.py: {unbound}"


def test_kernels_are_sm_100a_only(bq):
    out = subprocess.run(["cuobjdump", "-lelf", bq.KERNEL_LIB], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_no_cpu_fallback_without_a_device(bq):
    """On a machine without a GPU the product must fail loudly, not compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(bq.BqError, match="no CUDA device"):
        bq.Context(0)
    eng = bq.Engine()
    eng.add_table("t", [("a", 0, [1, 2, 3])])
    with pytest.raises(bq.BqError, match="no CUDA device"):
        eng.query("SELECT COUNT(*) FROM t")


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under bo-sql_b200/ may import, load or link it."""
    pat = re.compile(r"^\s*(from|import)\s+oracle\b|libbosql_ref|#include\s+\"[^\"]*oracle", re.M)
    for dirpath, _, files in os.walk(os.path.join(ROOT, "bo-sql_b200")):
        if os.sep + "build" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".cpp", ".hpp", ".cu", ".cuh", ".h", "Makefile")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert not pat.search(text), f"{f} references the oracle"


def test_bench_uses_the_oracle_only_in_its_checker_legs():
    """bench.py may touch oracle/ in the reference arm (run_reference) and in the cpu_baseline block of run_ours, nowhere
    else: the measured path imports the workload definitions from the package (bosql_b200.synthetic)."""
    import ast
    src = open(os.path.join(ROOT, "bench.py")).read()
    tree = ast.parse(src)
    lines = src.splitlines()
    funcs = {n.name: n for n in ast.walk(tree) if isinstance(n, ast.FunctionDef)}
    for node in ast.walk(tree):
        if isinstance(node, ast.ImportFrom) and node.module and node.module.split(".")[0] == "oracle" or \
           isinstance(node, ast.Import) and any(a.name.split(".")[0] == "oracle" for a in node.names):
            inside = [name for name, f in funcs.items() if f.lineno <= node.lineno <= f.end_lineno]
            if "run_reference" in inside:
                continue
            assert "run_ours" in inside, f"bench.py:{node.lineno} imports the oracle at module level"
            # inside run_ours: only under the cpu_baseline block
            back = "\n".join(lines[max(0, node.lineno - 12):node.lineno])
            assert "cpu_baseline" in back, f"bench.py:{node.lineno}: oracle import outside the cpu_baseline leg"
    for script in ("stress_configs.py", "perf_probe.py"):
        text = open(os.path.join(ROOT, "scripts", script)).read()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", text, re.M), f"scripts/{script} imports the oracle"
