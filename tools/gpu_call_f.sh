#!/bin/bash
# 2 GPUs: row-sharded parity test + the sharded chain on config C2 (30 894 bins)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "=== multi tests"; timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q > gpurun_out/f_tests.log 2>&1; echo rc=$?; tail -3 gpurun_out/f_tests.log
echo "=== sharded C2 on 2 GPUs (from counts)"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29701 tools/bench_sharded.py --res 100000 --from-counts --steps 2 > gpurun_out/f_sharded.log 2>&1; echo rc=$?; tail -2 gpurun_out/f_sharded.log | cut -c1-1500
