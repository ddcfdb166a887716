#!/usr/bin/env bash
# round 2, call Z2 (1 GPU): ncu --set full of the step's kernels (one launch each)
set -u
mkdir -p gpurun_out
HIA_BENCH_PROFILE=1 timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-checks --no-c3 > gpurun_out/r2z2_plain.log 2>&1 && \
HIA_BENCH_PROFILE=1 timeout 1200 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:"syrk_split3|symm_panel_tma|fill_rows_fast|row_prep|oe_apply_upper|inter_block|row_ptr|lanczos_step|tri_eig" -c 10 \
    -o gpurun_out/r2z2_step_full python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-checks --no-c3 > gpurun_out/r2z2_ncu.log 2>&1
echo "ncu full rc=$?"; tail -3 gpurun_out/r2z2_ncu.log
