#!/usr/bin/env bash
# round 2, call Z3 (1 GPU): ncu launch list of the timed steps, final code
set -u
mkdir -p gpurun_out
HIA_BENCH_PROFILE=1 timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-checks --no-c3 > gpurun_out/r2z3_plain.log 2>&1 && \
HIA_BENCH_PROFILE=1 timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off -c 400 \
    --csv --log-file gpurun_out/r2z3_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-checks --no-c3 > gpurun_out/r2z3_ncu1.log 2>&1
echo "ncu launches rc=$?"
