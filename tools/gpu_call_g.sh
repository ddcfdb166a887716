#!/bin/bash
# 8 GPUs: config C5 (123 541 bins) row-sharded Pearson + top-3 eigenvectors
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29702 tools/bench_sharded.py --res 25000 --steps 2 > gpurun_out/g_sharded.log 2>&1; echo rc=$?; tail -2 gpurun_out/g_sharded.log | cut -c1-1500
