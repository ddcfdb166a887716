#!/usr/bin/env bash
# round 2, call Q (2 GPUs): final code on 2 GPUs (row-sharded tests, bench --gpus 2 with the driver's arguments)
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -q -x -p no:cacheprovider > gpurun_out/r2q_multi.log 2>&1
echo "multi rc=$?"; tail -4 gpurun_out/r2q_multi.log
timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29741 \
    bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2q_bench2.log 2> gpurun_out/r2q_bench2.err
echo "bench2 rc=$?"; tail -c 600 gpurun_out/r2q_bench2.err
for c in 2 3; do
CUDA_VISIBLE_DEVICES=0 HIA_PREP_CTAS=$c timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-checks --no-c3 > gpurun_out/r2q_bench_prepctas$c.log 2>&1
echo "bench prep ctas $c rc=$?"
done
