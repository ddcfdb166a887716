#!/usr/bin/env bash
# round 2, call Z1 (1 GPU): full gpu suite, smoke, the driver's bench command, then the ncu launch list of the timed steps
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --durations=8 > gpurun_out/r2z_tests.log 2>&1
echo "tests rc=$?"; tail -14 gpurun_out/r2z_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2z_smoke.log 2>&1
echo "smoke rc=$?"; tail -2 gpurun_out/r2z_smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2z_bench.log 2> gpurun_out/r2z_bench.err
echo "bench rc=$?"; tail -c 300 gpurun_out/r2z_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2z_ref.log 2> gpurun_out/r2z_ref.err
echo "ref rc=$?"; tail -c 600 gpurun_out/r2z_ref.log
HIA_BENCH_PROFILE=1 timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-checks --no-c3 > gpurun_out/r2z_plain.log 2>&1 && \
HIA_BENCH_PROFILE=1 timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off -c 400 \
    --csv --log-file gpurun_out/r2z_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-checks --no-c3 > gpurun_out/r2z_ncu1.log 2>&1
echo "ncu launches rc=$?"
