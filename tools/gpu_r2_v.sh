#!/usr/bin/env bash
# round 2, call V (1 GPU): inter-tile fast path of the O/E apply kernel (parity first), scatter segment variants
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider -x --deselect tests/test_gpu_scale.py::test_c2_top3_eigenvectors_against_arpack > gpurun_out/r2v_tests.log 2>&1
echo "tests rc=$?"; tail -5 gpurun_out/r2v_tests.log
for v in 0 1 2 3; do
HIA_SCATTER_SEG=$v timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-c3 > gpurun_out/r2v_bench_seg$v.log 2>&1
echo "bench seg=$v rc=$?"
python - <<PY
import json
for l in open("gpurun_out/r2v_bench_seg$v.log"):
    if l.startswith("{"):
        r = json.loads(l); print($v, r["ms_per_step"], json.dumps(r["breakdown_ms"]), json.dumps(r.get("checks"))[:300])
PY
done
