#!/usr/bin/env bash
# round 2, call M (4 GPUs): k_chunk sweep of the row-sharded chain at 25 kb (123 541 bins) on 4 GPUs
set -u
mkdir -p gpurun_out
for kc in 512 1024 256; do
HIA_K_CHUNK=$kc timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 2973$((kc % 10)) \
    bench.py --gpus 4 --only-row-sharded > gpurun_out/r2m_rs4_kc$kc.log 2> gpurun_out/r2m_rs4_kc$kc.err
echo "k_chunk $kc rc=$?"; tail -c 300 gpurun_out/r2m_rs4_kc$kc.err
done
