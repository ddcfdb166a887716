#!/usr/bin/env bash
# round 2, call F (1 GPU): C3 as one CUDA graph per sample; ncu launch list + full capture of the HBM / eigen kernels
set -u
mkdir -p gpurun_out
timeout 400 python tools/bench_configs.py --only C3 --c3-workers 8 > gpurun_out/r2f_c3.log 2>&1
echo "c3 rc=$?"; grep "whole_sample" gpurun_out/r2f_c3.log | cut -c1-300
timeout 300 python tools/profile_step.py > gpurun_out/r2f_plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off \
    --csv --log-file gpurun_out/r2f_launches.csv python tools/profile_step.py > gpurun_out/r2f_ncu1.log 2>&1
echo "ncu launches rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:"row_prep|oe_apply_upper|fill_rows|symm_|jacobi|lanczos_step|intra_diag|inter_block" -c 14 \
    -o gpurun_out/r2f_full python tools/profile_step.py > gpurun_out/r2f_ncu2.log 2>&1
echo "ncu full rc=$?"; ls -la gpurun_out/r2f_full.ncu-rep
