#!/bin/bash
# 2 GPUs: the whole GPU suite (incl. the row-sharded tests) + bench.py launched like the driver does for N=2
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "=== tests"; timeout 600 python -m pytest tests -m gpu -q > gpurun_out/z_tests.log 2>&1; echo rc=$?; tail -3 gpurun_out/z_tests.log | cut -c1-300
echo "=== bench N=2"; timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/z_bench2.log 2>&1; echo rc=$?; tail -1 gpurun_out/z_bench2.log | cut -c1-400
