"""CPU: k_group_tables (bo-sql_b200/csrc/bq_groupby.cuh) - the very source nvcc compiles for sm_100a - built for the host
with the fibre-based CUDA emulation of tests/cpp/emu/cuda_emu.hpp and checked against std::map (accumulate and emit of
HashAggregate, src/exec/operator.cpp:984-1062).  The emulation runs every thread of a block between two synchronisation
points in a fixed order; EMU_SHUFFLE reverses that order on every other sweep, so a result that depended on which thread
wins a slot would differ between the two runs.  This checks the kernel's logic where there is no GPU; the product path
stays the CUDA build (tests/test_kernels_gpu.py::test_partition_aggregate runs the same cases on the device)."""
import ctypes as C
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_group_tables_kernel_on_the_cuda_emulation(tmp_path):
    exe = tmp_path / "group_tables_emu"
    src = os.path.join(ROOT, "tests", "cpp", "emu", "group_tables_emu.cpp")
    built = subprocess.run(["g++", "-std=c++20", "-O1", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "tests", "cpp", "emu"),
                            src, "-o", str(exe)], capture_output=True, text=True, timeout=300)
    assert built.returncode == 0, built.stderr[-3000:]
    for extra in ({}, {"EMU_SHUFFLE": "1"}):
        ran = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300, env=dict(os.environ, **extra))
        assert ran.returncode == 0, ran.stdout[-2000:] + ran.stderr[-2000:]
        assert "group tables emulation ok" in ran.stdout


def test_group_tables_sizing_rule():
    """bq_group_tables_plan needs no device: partitions x splits such that every 8192-slot (two sums: 4096-slot) table
    expects a load of at most 0.55, at most four splits per partition."""
    from __graft_entry__ import load_package
    L = load_package().kernel_lib()

    def plan(ndv, n_args):
        lp, sp = C.c_int(), C.c_int()
        return (lp.value, sp.value) if L.bq_group_tables_plan(ndv, n_args, C.byref(lp), C.byref(sp)) else None

    assert plan(12_500_000, 1) == (10, 3)           # configuration 4 on one GPU: 3072 tables of ~4070 groups
    assert plan(4_000_000, 1) == (10, 1)
    assert plan(4_000_000, 2) == (10, 2)
    assert plan(300_000, 1) == (7, 1)
    assert plan(100, 0) == (4, 1)
    assert plan(20_000_000, 1) is None              # more than four splits: the L2-resident table stays
    assert plan(0, 1) is None and plan(1000, 3) is None
    for ndv in (50_000, 3_333_333, 9_999_999, 18_000_000):
        for n_args in (0, 1, 2):
            got = plan(ndv, n_args)
            if got:
                slots = 4096 if n_args == 2 else 8192
                assert ndv / ((1 << got[0]) * got[1]) <= 0.55 * slots + 1
