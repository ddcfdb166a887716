#!/usr/bin/env bash
# round 2, final 8-GPU call: row-sharded tests at world 4 / 8, then the driver's own N = 8 command
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -q -x -p no:cacheprovider -k "4000 or 3333-4 or 2900" > gpurun_out/r2f8_multi.log 2>&1
echo "multi rc=$?"; tail -6 gpurun_out/r2f8_multi.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29751 \
    bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2f8_bench8.log 2> gpurun_out/r2f8_bench8.err
echo "bench8 rc=$?"; tail -c 600 gpurun_out/r2f8_bench8.err
