#!/usr/bin/env bash
# round 2 (1 GPU): 2^-30 fixed-point fold: accuracy probe, the similarity tests + the new C5-row-length test, the step
set -u
mkdir -p gpurun_out
timeout 600 python tools/accuracy_probe.py 2048 25000 > gpurun_out/acc_probe5.log 2>&1; echo "probe rc=$?"
python - <<'PY'
import json
for l in open("gpurun_out/acc_probe5.log"):
    if l.startswith("{"):
        r = json.loads(l); print(r["row0"], r["precision"], r["k_chunk"], "max %.2e mean %.2e p99.9 %.2e signed %.2e" % (r["max_abs_err"], r["mean_abs_err"], r["p99.9"], r["mean_signed_err"]))
PY
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider -x -k "similarity or pearson or cosine or sim_ or metric or checkerboard or hot_path or c5_row or c2_pearson or chain or smoke" > gpurun_out/r2acc2_tests.log 2>&1
echo "tests rc=$?"; tail -4 gpurun_out/r2acc2_tests.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-c3 > gpurun_out/r2acc2_bench.log 2>&1
echo "bench rc=$?"
python - <<'PY'
import json
for l in open("gpurun_out/r2acc2_bench.log"):
    if l.startswith("{"):
        r = json.loads(l); print(r["ms_per_step"], json.dumps(r["breakdown_ms"]), r["checks"]["pearson_stripe_max_abs_err"], r["eigvals"])
PY
