#!/usr/bin/env bash
# Round evidence on one B200 (run through gpurun): GPU tests, the default bench line, the ncu launch list of a short
# bench run and one `ncu --set full` capture of the fused scan kernel (Q1's and Q2's instance).  Outputs: gpurun_out/.
set -u
R=${1:-r1}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout=600 2>&1 | tail -5 > gpurun_out/pytest_${R}.log
python bench.py > gpurun_out/bench_${R}_n1.json 2> gpurun_out/bench_${R}_n1.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_${R}_reference.json 2>> gpurun_out/bench_${R}_n1.err
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu"
$CMD > gpurun_out/plain_${R}.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${R}.csv $CMD > gpurun_out/ncu_launch_${R}.log 2>&1
$CMD > gpurun_out/plain2_${R}.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k "regex:^k_scan$" -s 4 -c 3 -o gpurun_out/prof_${R}_scan $CMD > gpurun_out/ncu_full_${R}.log 2>&1
tail -3 gpurun_out/pytest_${R}.log
