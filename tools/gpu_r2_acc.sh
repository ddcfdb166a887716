#!/usr/bin/env bash
# round 2 (1 GPU): fixed-point fold of the SYRK chunk partials: accuracy probes, Pearson parity tests, perf per k_chunk
set -u
mkdir -p gpurun_out
timeout 600 python tools/accuracy_probe.py 2048 25000 > gpurun_out/acc_probe2.log 2>&1; echo "probe rc=$?"
grep -v tf32 gpurun_out/acc_probe2.log | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        r = json.loads(l); print(r['row0'], r['precision'], r['k_chunk'], 'max %.2e mean %.2e p99.9 %.2e signed %.2e' % (r['max_abs_err'], r['mean_abs_err'], r['p99.9'], r['mean_signed_err']))
"
grep tf32 gpurun_out/acc_probe2.log | cut -c1-200
timeout 600 python tools/accuracy_split.py > gpurun_out/acc_split2.log 2>&1; echo "split rc=$?"; cut -c1-600 gpurun_out/acc_split2.log
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider -x -k "similarity or pearson or cosine or sim_ or metric or checkerboard or hot_path or scale or chain" --deselect tests/test_gpu_scale.py::test_c4_oe_with_the_bin_size_quirk_at_real_scale > gpurun_out/r2acc_tests.log 2>&1
echo "tests rc=$?"; tail -4 gpurun_out/r2acc_tests.log
for kc in 512 256 128; do
HIA_K_CHUNK=$kc timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-c3 > gpurun_out/r2acc_bench_kc$kc.log 2>&1
echo "bench kc=$kc rc=$?"
python - <<PY
import json
for l in open("gpurun_out/r2acc_bench_kc$kc.log"):
    if l.startswith("{"):
        r = json.loads(l); print($kc, r["ms_per_step"], r["breakdown_ms"]["similarity_gemm"], r["checks"]["pearson_stripe_max_abs_err"], r["eigvals"])
PY
done
