#!/usr/bin/env bash
# round 2, call Y (1 GPU): row_ptr kernel with the shuffle; scatter parity + the step
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider -x -k "scatter or pipeline or hot_path or csr or preprocess" > gpurun_out/r2y_tests.log 2>&1
echo "tests rc=$?"; tail -4 gpurun_out/r2y_tests.log
for v in 0 1; do
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-c3 > gpurun_out/r2y_bench$v.log 2>&1
echo "bench rc=$?"
python - <<PY
import json
for l in open("gpurun_out/r2y_bench$v.log"):
    if l.startswith("{"):
        r = json.loads(l); print(r["ms_per_step"], json.dumps(r["breakdown_ms"]), r["e2e"]["ms_per_step"], r["checks"]["scatter_upper_mismatches"], {k: round(v["frac"], 3) for k, v in r["hbm_kernels"].items()})
PY
done
