#!/usr/bin/env bash
# round 2 (1 GPU): validation of the 1024-thread row prep: similarity tests at every size, the step with its checks
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider -x -k "similarity or pearson or cosine or sim_ or metric or checkerboard or hot_path or c5_row or c2_ or chain" --deselect tests/test_gpu_scale.py::test_c2_top3_eigenvectors_against_arpack > gpurun_out/r2val_tests.log 2>&1
echo "tests rc=$?"; tail -4 gpurun_out/r2val_tests.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-c3 > gpurun_out/r2val_bench.log 2>&1
echo "bench rc=$?"
python - <<'PY'
import json
for l in open("gpurun_out/r2val_bench.log"):
    if l.startswith("{"):
        r = json.loads(l); print(r["ms_per_step"], json.dumps(r["breakdown_ms"]), json.dumps(r["checks"]))
PY
