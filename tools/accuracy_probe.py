"""Single-GPU accuracy probe of the SYRK at genome-wide 25 kb row lengths (K = 123 541): a block of
`m` rows of the synthetic count map -> O/E + log (statistics of these rows only) -> Pearson of the block
against itself on the tcgen05 path, for several K-chunks / precisions, against torch fp64.

    python tools/accuracy_probe.py [m] [res]
"""
import ctypes
import json
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import numpy as np
import torch

from hiarch_b200 import _lib, ops, synth
from hiarch_b200.sharded import obs_d_exp_rows


def main():
    m = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
    res = int(sys.argv[2]) if len(sys.argv) > 2 else 25000
    lib = _lib.load()
    dev = torch.device("cuda", 0)
    lens = [n for _, n in synth.hg38_bins(res)]
    n = int(sum(lens))
    seps = synth.chr_seps_from_lens(lens)
    layout = ops.GenomeLayout(seps, dev)
    out = {"bins": n, "rows": m, "res": res, "runs": []}
    for row0 in (1000, lens[0] + lens[1] // 2):          # inside chr1; inside chr2
        counts, _ = synth.device_count_rows(lens, row0, m, 5, dev)
        oe, _ = obs_d_exp_rows(counts, layout, res, row0)
        del counts
        a = oe.double()
        a = a - a.mean(dim=1, keepdim=True)
        a = a / a.norm(dim=1, keepdim=True)
        ref = (a @ a.t()).clamp(-1.0, 1.0)
        ref.fill_diagonal_(1.0)
        del a
        for prec_name, prec in (("f16x3", ops.PREC_3XF16), ("tf32x3", ops.PREC_3XTF32)):
            kpad = lib.hia_sim_kpad(n, prec)
            elem = lib.hia_sim_plane_elem_bytes(prec)
            pdt = torch.float16 if elem == 2 else torch.float32
            planes = torch.zeros((2, m, kpad), dtype=pdt, device=dev)
            _lib.check(lib.hia_sim_prep_rows(oe.data_ptr(), m, n, oe.stride(0), ops.SIM_PEARSON, prec,
                                             planes[0].data_ptr(), planes[1].data_ptr(), ops._stream()), "prep")
            ws = torch.empty(lib.hia_sim_gemm_rows_workspace_bytes(m, n), dtype=torch.uint8, device=dev)
            ld = ops.padded_ld(m)
            for k_chunk in ((128, 256, 512, 1024, 2048) if prec == ops.PREC_3XF16 else (0,)):
                c = torch.zeros((m, ld), dtype=torch.float32, device=dev)
                _lib.check(lib.hia_sim_gemm_block(planes[0].data_ptr(), planes[1].data_ptr(), m,
                                                  planes[0].data_ptr(), planes[1].data_ptr(), m, prec, n,
                                                  0, m, 0, m, 1, c.data_ptr(), ld, c.data_ptr(), ld,
                                                  ops.SIM_PEARSON, k_chunk, ws.data_ptr(), ws.numel(), ops._stream()),
                           "gemm")
                torch.cuda.synchronize()
                err = (c[:, :m].double() - ref).abs()
                big = ref.abs() > 0.5
                rec = {"row0": row0, "precision": prec_name, "k_chunk": k_chunk, "max_abs_err": float(err.max()),
                       "mean_abs_err": float(err.mean()), "p99.9": float(err.flatten().kthvalue(int(err.numel() * 0.999)).values),
                       "mean_signed_err": float((c[:, :m].double() - ref).mean()),
                       "max_err_where_|rho|>0.5": float(err[big].max()) if big.any() else None,
                       "frac_|rho|>0.5": float(big.double().mean())}
                out["runs"].append(rec)
                print(json.dumps(rec), flush=True)
        del oe, ref
    json.dump(out, open("gpurun_out/accuracy_probe.json", "w"), indent=1)


main()
