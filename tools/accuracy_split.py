"""Where the SYRK's error at K = 123 541 comes from, for a few strongly correlated row pairs:
  repr : fp64 sum of the exact products of the fp16 hi/lo planes (3 terms)  vs  the fp64 Pearson value
  fold : the exact per-chunk partials folded sequentially in fp32 (RN)      vs  their fp64 sum
  mma  : the GPU value                                                       vs  the fp32 fold of exact partials
    python tools/accuracy_split.py
"""
import json
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import numpy as np
import torch

from hiarch_b200 import _lib, ops, synth
from hiarch_b200.sharded import obs_d_exp_rows


def main():
    m, res = 512, 25000
    lib = _lib.load()
    dev = torch.device("cuda", 0)
    lens = [n for _, n in synth.hg38_bins(res)]
    n = int(sum(lens))
    layout = ops.GenomeLayout(synth.chr_seps_from_lens(lens), dev)
    row0 = 1000
    counts, _ = synth.device_count_rows(lens, row0, m, 5, dev)
    oe, _ = obs_d_exp_rows(counts, layout, res, row0)
    a = oe.double()
    a = a - a.mean(dim=1, keepdim=True)
    a = a / a.norm(dim=1, keepdim=True)
    ref = (a @ a.t())
    prec = ops.PREC_3XF16
    kpad = lib.hia_sim_kpad(n, prec)
    planes = torch.zeros((2, m, kpad), dtype=torch.float16, device=dev)
    _lib.check(lib.hia_sim_prep_rows(oe.data_ptr(), m, n, oe.stride(0), ops.SIM_PEARSON, prec,
                                     planes[0].data_ptr(), planes[1].data_ptr(), ops._stream()), "prep")
    ws = torch.empty(lib.hia_sim_gemm_rows_workspace_bytes(m, n), dtype=torch.uint8, device=dev)
    ld = ops.padded_ld(m)
    hi, lo = planes[0].double(), planes[1].double()
    pairs = [(i, i + d) for i in range(0, 400, 25) for d in (1, 7, 60)]
    for k_chunk in (128, 512, 2048):
        c = torch.zeros((m, ld), dtype=torch.float32, device=dev)
        _lib.check(lib.hia_sim_gemm_block(planes[0].data_ptr(), planes[1].data_ptr(), m, planes[0].data_ptr(),
                                          planes[1].data_ptr(), m, prec, n, 0, m, 0, m, 1, c.data_ptr(), ld,
                                          c.data_ptr(), ld, ops.SIM_PEARSON, k_chunk, ws.data_ptr(), ws.numel(),
                                          ops._stream()), "gemm")
        torch.cuda.synchronize()
        e_repr, e_fold, e_mma, e_tot = [], [], [], []
        nch = (kpad + k_chunk - 1) // k_chunk
        for i, j in pairs:
            prod = hi[i] * hi[j] + hi[i] * lo[j] + lo[i] * hi[j]            # exact in fp64 (22 + 22 bits)
            prod = torch.nn.functional.pad(prod, (0, nch * k_chunk - kpad))
            part = prod.view(nch, k_chunk).sum(dim=1).cpu().numpy()          # fp64 partials
            exact = part.sum() * 2.0 ** -30
            acc = np.float32(0)
            for p in part:
                acc = np.float32(acc + np.float32(p))
            fold = float(acc) * 2.0 ** -30
            gpu = float(c[i, j])
            r = float(ref[i, j])
            e_repr.append(exact - r); e_fold.append(fold - exact); e_mma.append(gpu - fold); e_tot.append(gpu - r)
        f = lambda v: {"mean_abs": float(np.mean(np.abs(v))), "max_abs": float(np.max(np.abs(v))), "mean": float(np.mean(v))}
        print(json.dumps({"k_chunk": k_chunk, "pairs": len(pairs), "rho_mean_abs": float(np.mean([abs(float(ref[i, j])) for i, j in pairs])),
                          "repr": f(e_repr), "fold": f(e_fold), "mma": f(e_mma), "total": f(e_tot)}), flush=True)


main()
