#!/usr/bin/env bash
# round 2, call D (2 GPUs): GPU suite without the slow scale tests, bench N=1 (upper-only + register row prep),
# bench --gpus 2 (per-phase O/E timing of the row-sharded chain), SYRK A/B
set -u
mkdir -p gpurun_out
CUDA_VISIBLE_DEVICES=0 timeout 600 python -m pytest tests -m gpu -q -x -p no:cacheprovider --deselect tests/test_gpu_scale.py --deselect tests/test_gpu_multi.py > gpurun_out/r2d_tests.log 2>&1
echo "tests rc=$?"; tail -5 gpurun_out/r2d_tests.log
CUDA_VISIBLE_DEVICES=0 timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/r2d_bench.log 2> gpurun_out/r2d_bench.err
echo "bench rc=$?"; tail -c 800 gpurun_out/r2d_bench.err
CUDA_VISIBLE_DEVICES=0 HIA_PREP_REG=0 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2d_bench_prep2pass.log 2>&1
echo "bench (two-pass prep) rc=$?"
CUDA_VISIBLE_DEVICES=0 HIA_UPPER_ONLY=0 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2d_bench_mirror.log 2>&1
echo "bench (mirrored) rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 \
    bench.py --gpus 2 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2d_bench2.log 2> gpurun_out/r2d_bench2.err
echo "bench2 rc=$?"; tail -c 800 gpurun_out/r2d_bench2.err
CUDA_VISIBLE_DEVICES=0 timeout 200 python tools/ab_syrk.py hiarch_b200/build/ab/libhiarch_b200_r1.so hiarch_b200/libhiarch_b200.so > gpurun_out/r2d_ab.log 2>&1
echo "ab rc=$?"; cat gpurun_out/r2d_ab.log
