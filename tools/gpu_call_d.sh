#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "=== tests"; timeout 900 python -m pytest tests -m gpu -q > gpurun_out/d_tests.log 2>&1; echo rc=$?; tail -4 gpurun_out/d_tests.log
echo "=== probe"; timeout 300 python tools/perf_probe.py > gpurun_out/d_probe.log 2>&1; echo rc=$?; tail -1 gpurun_out/d_probe.log | cut -c1-1200
echo "=== bench"; timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/d_bench.log 2>&1; echo rc=$?; tail -1 gpurun_out/d_bench.log
echo "=== ncu launches"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"symm_panel|gram|jacobi|chol|panel_|ritz|row_prep|scatter|mirror|oe_apply|inter_block|row_ptr|fill_rows|zero_fill|intra_diag|finalize" -s 60 -c 400 --csv --log-file gpurun_out/launches_r1e.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/d_ncu_launches.log 2>&1; echo rc=$?
