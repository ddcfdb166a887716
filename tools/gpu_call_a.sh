#!/bin/bash
# GPU call A: parity of both CTA modes + both precisions, sweep timings
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "=== CTAS=1 tests"; HIA_SYRK_CTAS=1 timeout 300 python -m pytest tests -m gpu -q -x -k "similarity or preprocess or remap or checkerboard or pipeline" > gpurun_out/a_tests_cta1.log 2>&1; echo rc=$?; tail -3 gpurun_out/a_tests_cta1.log
echo "=== CTAS=2 tests"; HIA_SYRK_CTAS=2 timeout 300 python -m pytest tests -m gpu -q > gpurun_out/a_tests_cta2.log 2>&1; echo rc=$?; tail -5 gpurun_out/a_tests_cta2.log
echo "=== sweep CTAS=1"; HIA_SYRK_CTAS=1 timeout 240 python tools/syrk_sweep.py 30894 30894 0,256,1024 2,0 > gpurun_out/a_sweep_cta1.log 2>&1; echo rc=$?; cat gpurun_out/a_sweep_cta1.log | tail -12
echo "=== sweep CTAS=2"; HIA_SYRK_CTAS=2 timeout 240 python tools/syrk_sweep.py 30894 30894 0,256,1024 2,0 > gpurun_out/a_sweep_cta2.log 2>&1; echo rc=$?; cat gpurun_out/a_sweep_cta2.log | tail -12
echo "=== bench"; timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/a_bench.log 2>&1; echo rc=$?; tail -2 gpurun_out/a_bench.log
