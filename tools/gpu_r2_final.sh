#!/usr/bin/env bash
# round 2, last 1-GPU call: the full gpu suite, smoke and the driver's bench command on the final code
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --durations=6 > gpurun_out/r2fin_tests.log 2>&1
echo "tests rc=$?"; tail -12 gpurun_out/r2fin_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2fin_smoke.log 2>&1
echo "smoke rc=$?"; tail -2 gpurun_out/r2fin_smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2fin_bench.log 2> gpurun_out/r2fin_bench.err
echo "bench rc=$?"; tail -c 300 gpurun_out/r2fin_bench.err
