#!/usr/bin/env bash
# round 2 (1 GPU): row prep with 1024-thread CTAs (2 per SM) against 512-thread CTAs (4 per SM)
set -u
mkdir -p gpurun_out
for v in 512 1024 512 1024; do
HIA_PREP_THREADS=$v timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-c3 --no-checks > gpurun_out/r2prep_$v.log 2>&1
echo "bench threads=$v rc=$?"
python - <<PY
import json
for l in open("gpurun_out/r2prep_$v.log"):
    if l.startswith("{"):
        r = json.loads(l); print($v, r["ms_per_step"], json.dumps(r["breakdown_ms"]))
PY
done
