#!/usr/bin/env bash
# round 2, call S (1 GPU): the pipeline tests that cross the wavelet
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pipeline.py -q -p no:cacheprovider -k "norm_dis_step or one_click_chain" > gpurun_out/r2s_pipe.log 2>&1
echo "pipeline rc=$?"; tail -40 gpurun_out/r2s_pipe.log
