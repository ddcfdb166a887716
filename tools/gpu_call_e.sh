#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "=== tests"; timeout 900 python -m pytest tests -m gpu -q > gpurun_out/e_tests.log 2>&1; echo rc=$?; tail -3 gpurun_out/e_tests.log
echo "=== probe"; timeout 300 python tools/perf_probe.py > gpurun_out/e_probe.log 2>&1; echo rc=$?; tail -1 gpurun_out/e_probe.log | cut -c1-400
echo "=== bench"; timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/e_bench.log 2>&1; echo rc=$?; tail -1 gpurun_out/e_bench.log | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print(json.dumps({k: d[k] for k in ('value', 'ms_per_step', 'breakdown_ms', 'hbm_kernels', 'clocks')}))
print('e2e', d['e2e']['ms_per_step'], 'gemm frac', d['roofline']['frac'], d['roofline']['frac_of_burst_peak'])"
echo "=== ncu launches"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"symm_panel|jacobi|row_prep|scatter|mirror|row_ptr|fill_rows|zero_fill" -s 20 -c 120 --csv --log-file gpurun_out/launches_r1f.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/e_ncu_launches.log 2>&1; echo rc=$?
python tools/summarize_launches.py gpurun_out/launches_r1f.csv
