#!/bin/bash
# 2 GPUs: row-sharded parity test (incl. forced restarts)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "=== multi tests"; timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q > gpurun_out/f_tests.log 2>&1; echo rc=$?; tail -15 gpurun_out/f_tests.log | cut -c1-600
