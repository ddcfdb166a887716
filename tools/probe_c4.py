"""One short GPU call for config C4 (human chr1 at 10 kb, 24 896 bins): stage timings with the
current defaults, the two settings of the histogram aggregation (HIA_SEL_AGG=1/0) for the exact
percentile selects, and the prefix-sum intra-x path against the direct kernel (parity at a size the
direct kernel finishes quickly, timing at full size).

    python tools/probe_c4.py [--res 10000] [--direct-full]
Writes gpurun_out/probe_c4.json.
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import numpy as np
import torch

from hiarch_b200 import ops, synth


ONCE = False          # --once: every stage exactly once (the ncu launch-list pass)


def timed(fn, reps=2, warm=1):
    if ONCE:
        reps, warm = 1, 0
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    best = None
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        t = a.elapsed_time(b)
        best = t if best is None else min(best, t)
    return best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--res", type=int, default=10000)
    ap.add_argument("--direct-full", action="store_true", help="also time the direct intra-x kernel at full size (~4 s per call)")
    ap.add_argument("--once", action="store_true", help="run every stage once, default settings only (for ncu)")
    args = ap.parse_args()
    global ONCE
    ONCE = args.once
    t_start = time.time()
    out = {}

    os.makedirs("gpurun_out", exist_ok=True)

    def note(k, v):
        out[k] = v
        print(k, json.dumps(v), flush=True)
        json.dump(out, open("gpurun_out/probe_c4.json", "w"), indent=1)      # survives a cut-off call

    n = [nb for name, nb in synth.hg38_bins(args.res) if name == "chr1"][0]
    dense, seps = synth.device_dense_map([n], seed=7)
    layout = ops.GenomeLayout(seps)
    raw = ops.alloc_map(n)
    raw.copy_(dense)
    del dense
    cells = n * n
    note("bins", n)
    oe, _ = ops.obs_d_exp(raw, layout, args.res, log=True)
    de = ops.norm_de_mtx(oe)
    del raw, oe
    torch.cuda.empty_cache()
    filt = ops.alloc_map(n)

    # ---- exact percentile selects: aggregated vs plain shared-memory atomics
    for agg in (("1",) if ONCE else ("1", "0")):
        os.environ["HIA_SEL_AGG"] = agg
        ms = timed(lambda: ops.main_filter(de, layout, out=filt, decimals=-1))
        note(f"main_filter_agg{agg}", {"ms": ms, "gbs_8B_per_cell": 8 * cells / ms / 1e6})
        ms = timed(lambda: ops.remove_outliers(filt, out=filt))
        note(f"remove_outliers_agg{agg}", {"ms": ms, "gbs_12B_per_cell_3pass_select_plus_clip": 20 * cells / ms / 1e6})
        if agg == "1":
            ref_filt = filt.clone()
        else:
            note("main_filter_agg_settings_identical", bool(torch.equal(ref_filt, filt)))
            del ref_filt
    os.environ["HIA_SEL_AGG"] = "1"
    if not ONCE:
        ops.main_filter(de, layout, out=filt, decimals=-1)

    # ---- intra-x
    ms = timed(lambda: ops.intra_x_response(filt, method="prefix"))
    note("intra_x_prefix", {"ms": ms, "workspace_mb": ops._lib.load().hia_intra_x_workspace_bytes(n) / 2**20,
                            "cells_per_s": cells / ms * 1e3})
    m = 1000 if ONCE else 5000
    sub = filt[:m, :m]
    a = ops.intra_x_response(sub, method="prefix")
    b = ops.intra_x_response(sub, method="direct")
    rel = float(((a - b).abs() / b.abs().clamp_min(1e-300)).max().item())
    note("intra_x_prefix_vs_direct_5000", {"max_rel_diff": rel,
                                           "ms_prefix": timed(lambda: ops.intra_x_response(sub, method="prefix")),
                                           "ms_direct": timed(lambda: ops.intra_x_response(sub, method="direct"), reps=1)})
    if args.direct_full:
        note("intra_x_direct", {"ms": timed(lambda: ops.intra_x_response(filt, method="direct"), reps=1, warm=0)})

    # ---- the other C4 stages (same as tools/bench_configs.py)
    ms = timed(lambda: ops.intra_low_profile(filt, layout))
    note("intra_low_profile", {"ms": ms, "gbs_4B_per_cell": 4 * cells / ms / 1e6})
    ms = timed(lambda: ops.gf_scores(filt, layout, centers=[n // 2]))
    note("gf_s2_pool_scores", {"ms": ms, "gbs_4B_per_cell": 4 * cells / ms / 1e6})
    ms = timed(lambda: ops.get_local(de, n))
    smin, smax = int(n * .05), int(n * .15)
    band_cells = 2 * sum(n - d for d in range(smin, smax))
    note("checkerboard_get_local", {"ms": ms, "band_cells": band_cells, "band_cells_per_s": band_cells / ms * 1e3})
    tot = out["main_filter_agg1"]["ms"] + out["intra_x_prefix"]["ms"] + out["intra_low_profile"]["ms"] + \
        out["gf_s2_pool_scores"]["ms"] + out["checkerboard_get_local"]["ms"]
    note("c4_total", {"ms": tot, "cells_per_s": cells / tot * 1e3})
    note("wall_s", time.time() - t_start)


if __name__ == "__main__":
    main()
