#!/bin/bash
# Round-2 evidence run on one B200: full GPU test suite, then the ncu launch list and one `--set full` capture of the step
# kernel on the benchmark workload (B200_PROFILING.md recipe: each ncu pass only after the same command exited 0 without ncu).
set -x
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu_b.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu_b.log); tail -4 gpurun_out/r02_pytest_gpu_b.log
F="--no-cpu-baseline --no-e2e --no-other-variants --no-other-workloads"
timeout 300 python bench.py --steps 2 --warmup 3 $F > gpurun_out/r02_bench_short.json 2> gpurun_out/r02_bench_short.err || exit 1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 $F > gpurun_out/r02_ncu_launches.log 2>&1; echo launches rc=$?
timeout 900 ncu --set full --clock-control none --import-source on -k regex:column_step -s 3 -c 1 -f -o gpurun_out/r02_full python bench.py --steps 2 --warmup 3 $F > gpurun_out/r02_ncu_full.log 2>&1; echo full rc=$?
ls -la gpurun_out/
