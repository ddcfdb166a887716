#!/usr/bin/env bash
# round 2, call O (1 GPU): the driver's own end-of-round sequence (whole GPU suite, smoke, bench) + launch list + C3
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu -p no:cacheprovider --durations=6 > gpurun_out/r2o_tests.log 2>&1
echo "tests rc=$?"; tail -12 gpurun_out/r2o_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2o_smoke.log 2>&1
echo "smoke rc=$?"; tail -2 gpurun_out/r2o_smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2o_bench.log 2> gpurun_out/r2o_bench.err
echo "bench rc=$?"; tail -c 600 gpurun_out/r2o_bench.err
timeout 400 python tools/bench_configs.py --only C3 --c3-workers 4,8,12 --c3-branches 8 > gpurun_out/r2o_c3.log 2>&1
echo "c3 rc=$?"; grep "whole_sample" gpurun_out/r2o_c3.log | cut -c1-200
HIA_BENCH_PROFILE=1 timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-checks --no-c3 > gpurun_out/r2o_plain.log 2>&1 && \
HIA_BENCH_PROFILE=1 timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off \
    --csv --log-file gpurun_out/r2o_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-checks --no-c3 > gpurun_out/r2o_ncu1.log 2>&1
echo "ncu launches rc=$?"
