#!/usr/bin/env bash
# round 2, call P (1 GPU): inter-statistics fusion (tests, bench on / off)
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider --deselect tests/test_gpu_multi.py --deselect tests/test_gpu_scale.py > gpurun_out/r2p_tests.log 2>&1
echo "tests rc=$?"; tail -6 gpurun_out/r2p_tests.log
timeout 600 python -m pytest tests/test_gpu_scale.py -q -x -p no:cacheprovider -k "c2" > gpurun_out/r2p_scale.log 2>&1
echo "scale rc=$?"; tail -3 gpurun_out/r2p_scale.log
timeout 400 python bench.py --steps 10 --warmup 3 --no-c3 > gpurun_out/r2p_bench.log 2> gpurun_out/r2p_bench.err
echo "bench rc=$?"; tail -c 600 gpurun_out/r2p_bench.err
HIA_FUSE_STATS=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-checks --no-c3 > gpurun_out/r2p_bench_nofuse.log 2>&1
echo "bench (no fusion) rc=$?"
