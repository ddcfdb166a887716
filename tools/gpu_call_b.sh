#!/bin/bash
# GPU call B: full GPU suite, bench with breakdown, ncu launch list + one full capture of the SYRK
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "=== tests"; timeout 600 python -m pytest tests -m gpu -q > gpurun_out/b_tests.log 2>&1; echo rc=$?; tail -8 gpurun_out/b_tests.log
echo "=== probe"; timeout 300 python tools/perf_probe.py > gpurun_out/b_probe.log 2>&1; echo rc=$?; tail -6 gpurun_out/b_probe.log | cut -c1-1500
echo "=== bench"; timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/b_bench.log 2>&1; echo rc=$?; tail -1 gpurun_out/b_bench.log
echo "=== ncu launches"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1900 -c 900 --csv --log-file gpurun_out/launches_r1c.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/b_ncu_launches.log 2>&1; echo rc=$?
echo "=== ncu full"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:syrk_split3 -s 3 -c 1 -o gpurun_out/syrk_r1c python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/b_ncu_full.log 2>&1; echo rc=$?
ls -la gpurun_out | tail -12
