#!/usr/bin/env bash
# round 2, last multi-GPU call (4 GPUs): the row-sharded C5 chain only, final code
set -u
mkdir -p gpurun_out
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29757 \
    bench.py --gpus 4 --steps 3 --warmup 3 --only-row-sharded > gpurun_out/r2f4b_rs.log 2> gpurun_out/r2f4b_rs.err
echo "rs4 rc=$?"; tail -c 200 gpurun_out/r2f4b_rs.err
