"""Quick device timings of the hot-path kernels (CUDA events). Not a benchmark contract --
a development probe run under gpurun."""
import argparse
import json
import sys
import os
import time

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import numpy as np
import torch
from hiarch_b200 import ops, synth


def timeit(fn, reps=3, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts), sorted(ts)[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--res", type=int, default=100000)
    ap.add_argument("--chroms", type=int, default=24)
    ap.add_argument("--sim-rows", type=int, default=0, help="0 = whole map")
    ap.add_argument("--k-chunk", type=int, default=0)
    args = ap.parse_args()
    bins = synth.hg38_bins(args.res)[: args.chroms]
    lens = [n for _, n in bins]
    t0 = time.time()
    dense, seps = synth.device_dense_map(lens, seed=0)
    torch.cuda.synchronize()
    n = dense.shape[0]
    print(f"N={n} generated in {time.time()-t0:.1f}s", flush=True)
    layout = ops.GenomeLayout(seps)
    dmap = ops.alloc_map(n)
    dmap.copy_(dense)
    del dense
    cells = n * n
    res = {}
    # scatter from COO
    up = torch.triu(dmap)
    idx = up.nonzero(as_tuple=False)
    rows, cols = idx[:, 0].to(torch.int32).contiguous(), idx[:, 1].to(torch.int32).contiguous()
    vals = up[idx[:, 0], idx[:, 1]].contiguous()
    del up, idx
    nnz = rows.numel()
    out = ops.alloc_map(n)
    best, med = timeit(lambda: ops.scatter_coo_sym(rows, cols, vals, n, out=out))
    gb = (12 * nnz + 4 * cells) / 1e9
    res["scatter"] = dict(ms=best, gbs=gb / best * 1e3, nnz=nnz)
    assert torch.equal(out, dmap)
    del rows, cols, vals
    oe = ops.alloc_map(n)
    state = ops.OEState(layout)
    best, med = timeit(lambda: ops.obs_d_exp(dmap, layout, args.res, log=True, out=oe, state=state))
    res["oe_log_total"] = dict(ms=best, gbs=10 * cells / 1e9 / best * 1e3, cells_per_s=cells / best * 1e3)
    lib = __import__("hiarch_b200._lib", fromlist=["x"]).load()
    st = torch.cuda.current_stream().cuda_stream
    best, _ = timeit(lambda: lib.hia_oe_expected_pass(dmap.data_ptr(), n, dmap.stride(0), layout.chr_start.data_ptr(),
        layout.chr_start_host.ctypes.data, layout.n_chr, state.diag_sum.data_ptr(), state.diag_minpos.data_ptr(),
        state.inter_sum.data_ptr(), state.inter_nnz.data_ptr(), state.inter_minpos.data_ptr(), layout.bin_chr.data_ptr(), st))
    res["oe_expected_pass"] = dict(ms=best, gbs=2 * cells / 1e9 / best * 1e3)
    best, _ = timeit(lambda: lib.hia_oe_apply(dmap.data_ptr(), oe.data_ptr(), n, dmap.stride(0), layout.chr_start.data_ptr(),
        layout.n_chr, layout.bin_chr.data_ptr(), state.inv_expected.data_ptr(), state.inter_inv.data_ptr(),
        state.min_pos_oe.data_ptr(), 1, st))
    res["oe_apply_log"] = dict(ms=best, gbs=8 * cells / 1e9 / best * 1e3)
    print(json.dumps(res), flush=True)
    # similarity: both operand splits
    m = n if args.sim_rows == 0 else min(n, args.sim_rows)
    x = oe[:m]
    simout = ops.alloc_map(m)
    sub = 256
    xs = x[:sub].double()
    xs = xs - xs.mean(dim=1, keepdim=True)
    xs = xs / xs.norm(dim=1, keepdim=True)
    ref = xs @ xs.T
    for prec, name in ((ops.PREC_3XTF32, "tf32x3"), (ops.PREC_3XF16, "f16x3")):
        ws = torch.empty(lib.hia_similarity_workspace_bytes(m, n, prec), dtype=torch.uint8, device="cuda")
        bp, _ = timeit(lambda: lib.hia_similarity_prep(x.data_ptr(), m, n, x.stride(0), 1, prec, ws.data_ptr(), ws.numel(), st), reps=3)
        best, med = timeit(lambda: lib.hia_similarity_gemm(m, n, simout.data_ptr(), simout.stride(0), 1, prec, 0, -1,
                                                           args.k_chunk, ws.data_ptr(), ws.numel(), st), reps=2)
        err = (simout[:sub, :sub].double() - ref).abs()
        err.fill_diagonal_(0)
        # logical flops of the computed (upper) tiles ~ m*m*k (half of 2*m*m*k)
        res[f"similarity_{name}"] = dict(prep_ms=bp, gemm_ms=best, m=m, k=n, logical_tflops_half=m * m * n / best / 1e9,
                                         executed_mma_tflops=3 * m * m * n / best / 1e9, max_abs_err_corner=float(err.max()))
        print(json.dumps(res[f"similarity_{name}"]), flush=True)
        del ws
    t0 = time.time()
    w, v, info = ops.topk_eig(simout, k=3, return_info=True)
    torch.cuda.synchronize()
    print("topk_eig wall", time.time() - t0, "s; eigvals", w.cpu().numpy(), "info", info, flush=True)
    best, _ = timeit(lambda: ops.topk_eig(simout, k=3, max_restarts=1), reps=2)
    res["topk_eig_one_cycle_12steps"] = dict(ms=best)
    print(json.dumps(res), flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open("gpurun_out/perf_probe.json", "w"), indent=1)


if __name__ == "__main__":
    main()
