"""CPU: the C++ face of the drop-in.  tests/cpp/host_contract.cpp restates the reference's own C++ unit tests that need no
execution (types, columnar containers, catalog round trip through load_csv, plan construction facts) against
bo-sql_b200/host/*.hpp and links the two product libraries - exactly what a C++ caller of the reference would compile."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "bo-sql_b200")


def test_host_contract_compiles_and_holds(tmp_path):
    exe = tmp_path / "host_contract"
    cmd = ["g++", "-std=c++20", "-O1", "-Wall", "-Wextra", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(PKG, "host"),
           os.path.join(ROOT, "tests", "cpp", "host_contract.cpp"), "-o", str(exe), "-L", PKG, "-lbosql_b200_exec", "-lbosql_b200",
           f"-Wl,-rpath,{PKG}"]
    built = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert built.returncode == 0, built.stderr[-3000:]
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")           # belt and braces: nothing here may need a device
    ran = subprocess.run([str(exe), str(tmp_path)], capture_output=True, text=True, timeout=60, env=env)
    assert ran.returncode == 0, ran.stderr[-3000:]
    assert "host contract ok" in ran.stdout
