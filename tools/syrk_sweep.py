"""Time hia_similarity_gemm for both operand splits and several k_chunk values, with the max abs
error of a row stripe against fp64 (development probe).

    python tools/syrk_sweep.py [m] [k] [k_chunks,comma] [precisions,comma (0 = tf32x3, 2 = f16x3)]
"""
import sys, os, json
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import torch
from hiarch_b200 import _lib, ops

def main():
    m = int(sys.argv[1]) if len(sys.argv) > 1 else 30894
    k = int(sys.argv[2]) if len(sys.argv) > 2 else m
    chunks = [int(c) for c in sys.argv[3].split(",")] if len(sys.argv) > 3 else [0, 256, 512, 1024, 1 << 20]
    precs = [int(c) for c in sys.argv[4].split(",")] if len(sys.argv) > 4 else [2, 0]
    lib = _lib.load()
    g = torch.Generator(device="cuda"); g.manual_seed(0)
    x = ops.alloc_map(m, cols=k)
    x.normal_(generator=g)
    # correlated rows: a few shared components, so that entries are not all ~0
    base = torch.randn(4, k, device="cuda", generator=g)
    load = torch.randn(m, 4, device="cuda", generator=g)
    x.add_(load @ base).add_(0.3)
    out = ops.alloc_map(m)
    st = torch.cuda.current_stream().cuda_stream
    stripe = min(m, 384)
    xs = x[:stripe].double(); xs = xs - xs.mean(1, keepdim=True); xs = xs / xs.norm(dim=1, keepdim=True)
    ref = torch.empty((stripe, m), dtype=torch.float64, device="cuda")
    for c0 in range(0, m, 4096):
        xb = x[c0:c0 + 4096].double(); xb = xb - xb.mean(1, keepdim=True); xb = xb / xb.norm(dim=1, keepdim=True)
        ref[:, c0:c0 + xb.shape[0]] = xs @ xb.T
    ref[torch.arange(stripe), torch.arange(stripe)] = 1.0
    res = {}
    for prec in precs:
        ws = torch.empty(lib.hia_similarity_workspace_bytes(m, k, prec), dtype=torch.uint8, device="cuda")
        tp = []
        for rep in range(3):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            _lib.check(lib.hia_similarity_prep(x.data_ptr(), m, k, x.stride(0), 1, prec, ws.data_ptr(), ws.numel(), st), "prep")
            b.record(); torch.cuda.synchronize()
            tp.append(a.elapsed_time(b))
        print(f"precision={prec} prep: {min(tp):.3f} ms", flush=True)
        for kc in chunks:
            ts = []
            for rep in range(3):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                _lib.check(lib.hia_similarity_gemm(m, k, out.data_ptr(), out.stride(0), 1, prec, 0, -1, kc, ws.data_ptr(), ws.numel(), st), "gemm")
                b.record(); torch.cuda.synchronize()
                ts.append(a.elapsed_time(b))
            err = float((out[:stripe, :m].double() - ref).abs().max())
            res[f"{prec}:{kc}"] = dict(ms=min(ts), prep_ms=min(tp), err=err)
            print(f"precision={prec} k_chunk={kc}: {min(ts):.2f} ms  logical {m*m*k/min(ts)/1e9:.1f} TF  "
                  f"executed {3*m*m*k/min(ts)/1e9:.1f} TF  max abs err (stripe of {stripe} rows) {err:.3e}", flush=True)
        del ws
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open(f"gpurun_out/syrk_sweep_{m}x{k}.json", "w"), indent=1)
main()
