#!/bin/bash
# 4 GPUs: config C5 from synthetic COUNT rows (the clustered-spectrum case): O/E+log -> Pearson -> top-3 eigvec
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29704 tools/bench_sharded.py --res 25000 --from-counts --steps 1 > gpurun_out/h4_sharded.log 2>&1; echo rc=$?; tail -3 gpurun_out/h4_sharded.log | cut -c1-1500
