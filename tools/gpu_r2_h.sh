#!/usr/bin/env bash
# round 2, call H (1 GPU): whole GPU suite (chain test, CSR, metrics ...), bench, C3, launch list of the bench
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider --deselect tests/test_gpu_multi.py --durations=8 > gpurun_out/r2h_tests.log 2>&1
echo "tests rc=$?"; tail -14 gpurun_out/r2h_tests.log
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/r2h_bench.log 2> gpurun_out/r2h_bench.err
echo "bench rc=$?"; tail -c 800 gpurun_out/r2h_bench.err
HIA_EIG_FIRST_CHECK=6 HIA_EIG_CHECK_EVERY=1 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-checks > gpurun_out/r2h_bench_fc6.log 2>&1
echo "bench fc6 rc=$?"
timeout 400 python tools/bench_configs.py --only C3 --c3-workers 8 --c3-branches 8 > gpurun_out/r2h_c3.log 2>&1
echo "c3 rc=$?"; grep "whole_sample" gpurun_out/r2h_c3.log | cut -c1-300
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-checks > gpurun_out/r2h_plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1500 \
    --csv --log-file gpurun_out/r2h_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-checks > gpurun_out/r2h_ncu1.log 2>&1
echo "ncu launches rc=$?"
