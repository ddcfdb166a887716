#!/usr/bin/env bash
# round 2, call X (1 GPU): plain-store fill kernel: parity suite, then A/B against the all-atomic fill
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider -x --deselect tests/test_gpu_scale.py::test_c2_top3_eigenvectors_against_arpack --deselect tests/test_gpu_scale.py::test_c4_oe_with_the_bin_size_quirk_at_real_scale > gpurun_out/r2x_tests.log 2>&1
echo "tests rc=$?"; tail -5 gpurun_out/r2x_tests.log
for v in 1 0 1 0; do
HIA_FILL_ATOMIC=$v timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-c3 > gpurun_out/r2x_bench_atomic$v.log 2>&1
echo "bench atomic=$v rc=$?"
python - <<PY
import json
for l in open("gpurun_out/r2x_bench_atomic$v.log"):
    if l.startswith("{"):
        r = json.loads(l); print($v, r["ms_per_step"], json.dumps(r["breakdown_ms"]), r["e2e"]["ms_per_step"], r["checks"]["scatter_upper_mismatches"], {k: round(v["frac"], 3) for k, v in r["hbm_kernels"].items()})
PY
done
