#!/usr/bin/env bash
# round 2, call J (1 GPU): cp.async eigen pass + L2-hinted row prep; launch list of the timed bench steps
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider --deselect tests/test_gpu_multi.py --deselect tests/test_gpu_scale.py > gpurun_out/r2j_tests.log 2>&1
echo "tests rc=$?"; tail -4 gpurun_out/r2j_tests.log
timeout 600 python -m pytest tests/test_gpu_scale.py -q -x -p no:cacheprovider -k "top3" > gpurun_out/r2j_scale.log 2>&1
echo "scale rc=$?"; tail -3 gpurun_out/r2j_scale.log
timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/r2j_bench.log 2> gpurun_out/r2j_bench.err
echo "bench rc=$?"; tail -c 800 gpurun_out/r2j_bench.err
HIA_EIG_ASYNC=0 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-checks --no-c3 > gpurun_out/r2j_bench_noasync.log 2>&1
echo "bench (register-staged pass) rc=$?"
HIA_BENCH_PROFILE=1 timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-checks --no-c3 > gpurun_out/r2j_plain.log 2>&1 && \
HIA_BENCH_PROFILE=1 timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off \
    --csv --log-file gpurun_out/r2j_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-checks --no-c3 > gpurun_out/r2j_ncu1.log 2>&1
echo "ncu launches rc=$?"
