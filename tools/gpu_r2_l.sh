#!/usr/bin/env bash
# round 2, call L (2 GPUs): restart change (k_keep) on 1 and 2 GPUs, bench --gpus 2 with k_chunk 256 on the 50 kb chain
set -u
mkdir -p gpurun_out
CUDA_VISIBLE_DEVICES=0 timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider --deselect tests/test_gpu_multi.py --deselect tests/test_gpu_scale.py > gpurun_out/r2l_tests.log 2>&1
echo "tests rc=$?"; tail -4 gpurun_out/r2l_tests.log
timeout 600 python -m pytest tests/test_gpu_multi.py -q -x -p no:cacheprovider > gpurun_out/r2l_multi.log 2>&1
echo "multi rc=$?"; tail -4 gpurun_out/r2l_multi.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 \
    bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2l_bench2.log 2> gpurun_out/r2l_bench2.err
echo "bench2 rc=$?"; tail -c 600 gpurun_out/r2l_bench2.err
CUDA_VISIBLE_DEVICES=0 timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2l_smoke.log 2>&1
echo "smoke rc=$?"; tail -2 gpurun_out/r2l_smoke.log
