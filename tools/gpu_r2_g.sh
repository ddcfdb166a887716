#!/usr/bin/env bash
# round 2, call G (1 GPU): CSR tests, Jacobi fix, check cadence, C3 graph, ncu of jacobi / prep_reg
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -p no:cacheprovider --deselect tests/test_gpu_scale.py --deselect tests/test_gpu_multi.py > gpurun_out/r2g_tests.log 2>&1
echo "tests rc=$?"; tail -5 gpurun_out/r2g_tests.log
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/r2g_bench.log 2> gpurun_out/r2g_bench.err
echo "bench rc=$?"; tail -c 800 gpurun_out/r2g_bench.err
for fc in 5 6 7; do
HIA_EIG_FIRST_CHECK=$fc HIA_EIG_CHECK_EVERY=1 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-checks > gpurun_out/r2g_bench_fc$fc.log 2>&1
echo "bench first-check $fc rc=$?"
done
timeout 400 python tools/bench_configs.py --only C3 --c3-workers 8 --c3-branches 8 > gpurun_out/r2g_c3.log 2>&1
echo "c3 rc=$?"; grep "whole_sample" gpurun_out/r2g_c3.log | cut -c1-300
timeout 300 python tools/profile_step.py > gpurun_out/r2g_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:"jacobi|row_prep_reg|symm_upper_kernel" -c 3 \
    -o gpurun_out/r2g_full python tools/profile_step.py > gpurun_out/r2g_ncu2.log 2>&1
echo "ncu full rc=$?"
