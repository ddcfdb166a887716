"""Time hia_similarity_gemm for several k_chunk values (development probe)."""
import sys, os, json
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import torch
from hiarch_b200 import _lib, ops

def main():
    m = int(sys.argv[1]) if len(sys.argv) > 1 else 30894
    k = int(sys.argv[2]) if len(sys.argv) > 2 else m
    chunks = [int(c) for c in sys.argv[3].split(",")] if len(sys.argv) > 3 else [256, 512, 1024, 4096, 1 << 20]
    lib = _lib.load()
    g = torch.Generator(device="cuda"); g.manual_seed(0)
    x = ops.alloc_map(m, cols=k)
    x.normal_(generator=g)
    out = ops.alloc_map(m)
    ws = torch.empty(lib.hia_similarity_workspace_bytes(m, k, 0), dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    _lib.check(lib.hia_similarity_prep(x.data_ptr(), m, k, x.stride(0), 1, ws.data_ptr(), ws.numel(), st), "prep")
    res = {}
    for kc in chunks:
        ts = []
        for rep in range(3):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            _lib.check(lib.hia_similarity_gemm(m, k, out.data_ptr(), out.stride(0), 1, 0, -1, kc, ws.data_ptr(), ws.numel(), st), "gemm")
            b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        res[kc] = min(ts)
        print(f"k_chunk={kc}: {min(ts):.2f} ms  logical {m*m*k/min(ts)/1e9:.1f} TF  executed {3*m*m*k/min(ts)/1e9:.1f} TF", flush=True)
main()
