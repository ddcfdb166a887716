#!/usr/bin/env bash
# round 2, call K (1 GPU): streaming hand-off, fill kernel layout; ncu --set full of the SYRK (roofline.traffic)
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider --deselect tests/test_gpu_multi.py --deselect tests/test_gpu_scale.py > gpurun_out/r2k_tests.log 2>&1
echo "tests rc=$?"; tail -4 gpurun_out/r2k_tests.log
timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/r2k_bench.log 2> gpurun_out/r2k_bench.err
echo "bench rc=$?"; tail -c 800 gpurun_out/r2k_bench.err
HIA_BENCH_PROFILE=1 timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-checks --no-c3 > gpurun_out/r2k_plain.log 2>&1 && \
HIA_BENCH_PROFILE=1 timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"syrk_split3|symm_panel_async|fill_rows|row_prep" -c 4 \
    -o gpurun_out/r2k_syrk_full python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-checks --no-c3 > gpurun_out/r2k_ncu.log 2>&1
echo "ncu full rc=$?"
