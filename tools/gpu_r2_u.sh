#!/usr/bin/env bash
# round 2, call U (1 GPU): TMA-staged eigen pass: parity tests, then the step with each variant
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider -k "eig or lanczos or hot_path or chain or scale" --deselect tests/test_gpu_scale.py::test_c4_oe_with_the_bin_size_quirk_at_real_scale > gpurun_out/r2u_tests.log 2>&1
echo "tests rc=$?"; tail -5 gpurun_out/r2u_tests.log
for v in 0 1 2 3; do
HIA_EIG_TMA=$v timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-checks --no-c3 > gpurun_out/r2u_bench_tma$v.log 2>&1
echo "bench tma=$v rc=$?"
python - <<PY
import json
for l in open("gpurun_out/r2u_bench_tma$v.log"):
    if l.startswith("{"):
        r = json.loads(l); print($v, r["ms_per_step"], json.dumps(r["breakdown_ms"]), r.get("eigvals"))
PY
done
