#!/usr/bin/env bash
# round 2, call C (1 GPU): BASELINE-scale parity tests + SYRK A/B against the round-1 build
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_scale.py -q -x -p no:cacheprovider --durations=10 > gpurun_out/r2b_scale.log 2>&1
echo "scale rc=$?"; tail -15 gpurun_out/r2b_scale.log
timeout 200 python tools/ab_syrk.py hiarch_b200/build/ab/libhiarch_b200_r1.so hiarch_b200/libhiarch_b200.so > gpurun_out/r2b_ab.log 2>&1
echo "ab rc=$?"; cat gpurun_out/r2b_ab.log
