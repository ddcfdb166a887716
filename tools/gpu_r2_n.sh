#!/usr/bin/env bash
# round 2, call N (1 GPU): partial projected eigensolver (tests, bench, check cadence, C3)
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider --deselect tests/test_gpu_multi.py --deselect tests/test_gpu_scale.py > gpurun_out/r2n_tests.log 2>&1
echo "tests rc=$?"; tail -6 gpurun_out/r2n_tests.log
timeout 600 python -m pytest tests/test_gpu_scale.py -q -x -p no:cacheprovider -k "top3" > gpurun_out/r2n_scale.log 2>&1
echo "scale rc=$?"; tail -3 gpurun_out/r2n_scale.log
timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/r2n_bench.log 2> gpurun_out/r2n_bench.err
echo "bench rc=$?"; tail -c 600 gpurun_out/r2n_bench.err
HIA_EIG_JACOBI=1 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-checks --no-c3 > gpurun_out/r2n_bench_jacobi.log 2>&1
echo "bench (jacobi) rc=$?"
for fc in 5 6 7; do
HIA_EIG_FIRST_CHECK=$fc HIA_EIG_CHECK_EVERY=1 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-checks --no-c3 > gpurun_out/r2n_bench_fc$fc.log 2>&1
echo "bench first-check $fc rc=$?"
done
timeout 400 python tools/bench_configs.py --only C3 --c3-workers 8 --c3-branches 8 > gpurun_out/r2n_c3.log 2>&1
echo "c3 rc=$?"; grep "whole_sample" gpurun_out/r2n_c3.log | cut -c1-200
