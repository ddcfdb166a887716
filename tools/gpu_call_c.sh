#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "=== tests"; timeout 600 python -m pytest tests -m gpu -q > gpurun_out/c_tests.log 2>&1; echo rc=$?; tail -6 gpurun_out/c_tests.log
echo "=== probe"; timeout 300 python tools/perf_probe.py > gpurun_out/c_probe.log 2>&1; echo rc=$?; grep "topk_eig" gpurun_out/c_probe.log | cut -c1-600
echo "=== bench"; timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/c_bench.log 2>&1; echo rc=$?; tail -1 gpurun_out/c_bench.log
echo "=== ncu launches (eig kernels)"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"symm_panel|gram|jacobi|chol|panel_|ritz|row_prep|scatter|mirror|oe_apply|inter_block" -s 60 -c 400 --csv --log-file gpurun_out/launches_r1d.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/c_ncu_launches.log 2>&1; echo rc=$?
