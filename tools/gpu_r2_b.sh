#!/usr/bin/env bash
# round 2, call B (2 GPUs): row-sharded tests (ring over peer memory), bench --gpus 2 with the row_sharded
# sub-record; plus the scale tests and the SYRK A/B on GPU 0
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -q -x -p no:cacheprovider > gpurun_out/r2b_multi.log 2>&1
echo "multi rc=$?"; tail -15 gpurun_out/r2b_multi.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 \
    bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2b_bench2.log 2> gpurun_out/r2b_bench2.err
echo "bench2 rc=$?"; tail -c 1500 gpurun_out/r2b_bench2.err
