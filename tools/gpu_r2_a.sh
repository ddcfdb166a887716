#!/usr/bin/env bash
# round 2, call A (1 GPU): whole GPU suite incl. the BASELINE-scale tests, then the headline bench
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -rxX -x -p no:cacheprovider --durations=15 > gpurun_out/r2a_tests.log 2>&1
echo "tests rc=$?"
tail -5 gpurun_out/r2a_tests.log
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/r2a_bench.log 2> gpurun_out/r2a_bench.err
echo "bench rc=$?"
tail -c 600 gpurun_out/r2a_bench.err
