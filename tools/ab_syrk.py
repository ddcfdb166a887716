"""A/B timing of hia_similarity_gemm between two builds of the library on the SAME box, interleaved
(the contraction runs under the power cap: box-to-box variation is a few per cent).

    python tools/ab_syrk.py hiarch_b200/build/ab/libhiarch_b200_r1.so hiarch_b200/libhiarch_b200.so [m]
"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import torch
from hiarch_b200 import _lib


def bind(path):
    lib = ctypes.CDLL(os.path.abspath(path))
    for name in ("hia_similarity_workspace_bytes", "hia_similarity_prep", "hia_similarity_gemm"):
        res, args = _lib.PROTOTYPES[name]
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    return lib


def main():
    paths = sys.argv[1:3]
    m = int(sys.argv[3]) if len(sys.argv) > 3 else 30894
    libs = [bind(p) for p in paths]
    g = torch.Generator(device="cuda")
    g.manual_seed(0)
    x = torch.empty((m, (m + 3) // 4 * 4), device="cuda")[:, :m]
    x.normal_(generator=g)
    out = torch.empty_like(x)
    st = torch.cuda.current_stream().cuda_stream
    ws = torch.empty(max(l.hia_similarity_workspace_bytes(m, m, 2) for l in libs), dtype=torch.uint8, device="cuda")
    libs[0].hia_similarity_prep(x.data_ptr(), m, m, x.stride(0), 1, 2, ws.data_ptr(), ws.numel(), st)
    times = [[], []]
    for rep in range(6):
        for i, l in enumerate(libs):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            rc = l.hia_similarity_gemm(m, m, out.data_ptr(), out.stride(0), 1, 2, 0, -1, 0, ws.data_ptr(), ws.numel(), st)
            b.record()
            torch.cuda.synchronize()
            assert rc == 0, rc
            if rep > 0:
                times[i].append(a.elapsed_time(b))
    for p, t in zip(paths, times):
        print(f"{p}: {min(t):.2f} min / {sum(t) / len(t):.2f} mean ms  {[round(v, 2) for v in t]}", flush=True)


main()
