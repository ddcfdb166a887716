#!/usr/bin/env bash
# round 2, final 2-GPU call: row-sharded tests at world 2, the driver's N = 2 command
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -q -x -p no:cacheprovider > gpurun_out/r2f2_multi.log 2>&1
echo "multi rc=$?"; tail -4 gpurun_out/r2f2_multi.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29755 \
    bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2f2_bench2.log 2> gpurun_out/r2f2_bench2.err
echo "bench2 rc=$?"; tail -c 300 gpurun_out/r2f2_bench2.err
