#!/usr/bin/env bash
# round 2 (1 GPU): fp16 planes scaled by 2^15 (no multiply before the fixed-point fold): accuracy probe, tests, step
set -u
mkdir -p gpurun_out
timeout 600 python tools/accuracy_probe.py 2048 25000 > gpurun_out/acc_probe6.log 2>&1; echo "probe rc=$?"
python - <<'PY'
import json
for l in open("gpurun_out/acc_probe6.log"):
    if l.startswith("{"):
        r = json.loads(l); print(r["row0"], r["precision"], r["k_chunk"], "max %.2e mean %.2e signed %.2e" % (r["max_abs_err"], r["mean_abs_err"], r["mean_signed_err"]))
PY
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider -x -k "similarity or pearson or cosine or sim_ or metric or checkerboard or hot_path or c5_row or c2_ or chain or smoke" --deselect tests/test_gpu_scale.py::test_c2_top3_eigenvectors_against_arpack > gpurun_out/r2val2_tests.log 2>&1
echo "tests rc=$?"; tail -4 gpurun_out/r2val2_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
for i in 1 2; do
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-c3 > gpurun_out/r2val2_bench$i.log 2>&1
echo "bench rc=$?"
python - <<PY
import json
for l in open("gpurun_out/r2val2_bench$i.log"):
    if l.startswith("{"):
        r = json.loads(l); print(r["ms_per_step"], json.dumps(r["breakdown_ms"]), r["checks"]["pearson_stripe_max_abs_err"], r["eigvals"])
PY
done
