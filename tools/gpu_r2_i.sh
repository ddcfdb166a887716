#!/usr/bin/env bash
# round 2, call I (8 GPUs): row-sharded tests at world 3 / 4 / 8, then the driver's own N = 8 command
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -q -x -p no:cacheprovider -k "4000 or 2900 or 3333-4" > gpurun_out/r2i_multi.log 2>&1
echo "multi rc=$?"; tail -6 gpurun_out/r2i_multi.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29721 \
    bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/r2i_bench8.log 2> gpurun_out/r2i_bench8.err
echo "bench8 rc=$?"; tail -c 1500 gpurun_out/r2i_bench8.err
