#!/usr/bin/env bash
# round 2, call T (2 GPUs): row-sharded chain only, with the eigen phase split
set -u
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29743 \
    bench.py --gpus 2 --steps 3 --warmup 3 --only-row-sharded > gpurun_out/r2t_rs2.log 2> gpurun_out/r2t_rs2.err
echo "rs2 rc=$?"; tail -c 300 gpurun_out/r2t_rs2.err; tail -c 1500 gpurun_out/r2t_rs2.log
