#!/usr/bin/env bash
# round 2, final 4-GPU call: the driver's N = 4 command
set -u
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29753 \
    bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r2f4_bench4.log 2> gpurun_out/r2f4_bench4.err
echo "bench4 rc=$?"; tail -c 600 gpurun_out/r2f4_bench4.err
