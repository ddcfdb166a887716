#!/usr/bin/env bash
# round 2, call E (1 GPU): eigen changes (cooperative step, upper-triangle pass), row prep clusters
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -p no:cacheprovider --deselect tests/test_gpu_scale.py --deselect tests/test_gpu_multi.py > gpurun_out/r2e_tests.log 2>&1
echo "tests rc=$?"; tail -5 gpurun_out/r2e_tests.log
timeout 600 python -m pytest tests/test_gpu_scale.py -q -x -p no:cacheprovider -k "top3 or scatter or pearson" > gpurun_out/r2e_scale.log 2>&1
echo "scale rc=$?"; tail -5 gpurun_out/r2e_scale.log
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/r2e_bench.log 2> gpurun_out/r2e_bench.err
echo "bench rc=$?"; tail -c 800 gpurun_out/r2e_bench.err
HIA_EIG_UPPER=0 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2e_bench_noupper.log 2>&1
echo "bench (full pass) rc=$?"
HIA_EIG_COOP=0 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2e_bench_nocoop.log 2>&1
echo "bench (no coop) rc=$?"
HIA_PREP_REG=0 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2e_bench_prep2pass.log 2>&1
echo "bench (two-pass prep) rc=$?"
