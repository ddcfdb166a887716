#!/usr/bin/env bash
# round 2, call R (1 GPU): wavelet (f4) parity, full gpu suite with the new inter-block loop, bench
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_wavelet.py -q -x -p no:cacheprovider > gpurun_out/r2r_wavelet.log 2>&1
echo "wavelet rc=$?"; tail -25 gpurun_out/r2r_wavelet.log
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --deselect tests/test_gpu_wavelet.py --durations=6 > gpurun_out/r2r_tests.log 2>&1
echo "tests rc=$?"; tail -30 gpurun_out/r2r_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2r_bench.log 2> gpurun_out/r2r_bench.err
echo "bench rc=$?"; tail -c 400 gpurun_out/r2r_bench.err
timeout 300 python - > gpurun_out/r2r_wavelet_time.log 2>&1 <<'PY'
import torch, time, numpy as np
from hiarch_b200 import ops, _lib
for n in (600, 1250, 2490):
    x = torch.randn((n, n), device="cuda", dtype=torch.float32)
    ws = torch.empty(_lib.load().hia_wavelet_workspace_bytes(n, n, 0), dtype=torch.uint8, device="cuda")
    out = ops.alloc_map(n, "cuda")
    for _ in range(3): ops.wavelet_denoise_db3(x, out=out, workspace=ws)
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(10): ops.wavelet_denoise_db3(x, out=out, workspace=ws)
    e1.record(); torch.cuda.synchronize()
    print(n, "ms per block", e0.elapsed_time(e1) / 10)
PY
echo "wavelet time rc=$?"; cat gpurun_out/r2r_wavelet_time.log
