#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "=== bench"; timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/final_bench.log 2>&1; echo rc=$?; tail -1 gpurun_out/final_bench.log | cut -c1-300
echo "=== ncu launches"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 1900 -c 900 --csv --log-file gpurun_out/launches_r1_final.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/final_ncu_launches.log 2>&1; echo rc=$?
python tools/summarize_launches.py gpurun_out/launches_r1_final.csv | head -30
