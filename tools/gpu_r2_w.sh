#!/usr/bin/env bash
# round 2, call W (1 GPU): fp32-arithmetic row prep: parity suite, then the step
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider -x --deselect tests/test_gpu_scale.py::test_c2_top3_eigenvectors_against_arpack --deselect tests/test_gpu_scale.py::test_c4_oe_with_the_bin_size_quirk_at_real_scale > gpurun_out/r2w_tests.log 2>&1
echo "tests rc=$?"; tail -5 gpurun_out/r2w_tests.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-c3 > gpurun_out/r2w_bench.log 2>&1
echo "bench rc=$?"
python - <<PY
import json
for l in open("gpurun_out/r2w_bench.log"):
    if l.startswith("{"):
        r = json.loads(l); print(r["ms_per_step"], json.dumps(r["breakdown_ms"]), json.dumps(r.get("checks")), {k: round(v["frac"], 3) for k, v in r["hbm_kernels"].items()})
PY
