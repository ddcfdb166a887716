"""Device timings of the two BASELINE.json configurations bench.py does not cover (one GPU, CUDA
events, synthetic maps built in HBM):

  C3  one sample of the multi-species batch: 24 per-chromosome maps at 50 kb (935 ... 4 980 bins),
      each through scatter -> O/E+log -> Pearson -> top-3 eigenvectors
  C4  human chr1 at 10 kb (24 896 bins): GF_S1 filter, row profile + smoothing, intra-x response,
      GF_S2 pooling + template scores, checkerboard (band-limited cosine distance + entropy)

Every stage is timed on its own and a failure of one stage does not stop the others.
    python tools/bench_configs.py [--c3-res 50000] [--c4-res 10000]
"""
import argparse
import json
import os
import sys
import time
import traceback

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import numpy as np
import torch

from hiarch_b200 import ops, synth
from hiarch_b200.pipeline import HotPath, HotPathPool, SampleGraph, dense_to_coo


def timed(fn, reps=2, warm=1):
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


def stage(res, name, fn, **extra):
    try:
        ms = timed(fn)
        res[name] = dict(ms=ms, **{k: (v(ms) if callable(v) else v) for k, v in extra.items()})
    except Exception as exc:                                   # keep going: one GPU call, many stages
        res[name] = dict(error=f"{type(exc).__name__}: {exc}")
        traceback.print_exc()
    print(name, json.dumps(res[name]), flush=True)


def run_c3(args):
    bins = synth.hg38_bins(args.c3_res)
    paths, coos, cells = [], [], 0
    for i, (_, n) in enumerate(bins):
        dense, seps = synth.device_dense_map([n], seed=100 + i)
        coos.append(dense_to_coo(dense))
        paths.append(HotPath(seps, args.c3_res, k_eig=3))
        cells += n * n
        del dense
    res = {"chromosomes": len(bins), "bins_min": min(n for _, n in bins), "bins_max": max(n for _, n in bins),
           "cells_per_sample": cells}

    def sample():
        for hp, (r, c, v) in zip(paths, coos):
            hp.run_device(r, c, v)

    stage(res, "whole_sample", sample, cells_per_s=lambda ms: cells / ms * 1e3,
          samples_per_s=lambda ms: 1e3 / ms)
    # the same sample with the chromosomes dealt to W host threads / streams (HotPathPool)
    for w in [int(x) for x in args.c3_workers.split(",") if x]:
        pool = HotPathPool(paths, workers=w)
        stage(res, f"whole_sample_pool{w}", lambda: pool.run_device(coos),
              cells_per_s=lambda ms: cells / ms * 1e3, samples_per_s=lambda ms: 1e3 / ms)
        pool.close()
    # the whole sample as ONE CUDA graph with forked branches (device time incl. the result read-back)
    for br in [int(x) for x in args.c3_branches.split(",") if x]:
        try:
            sg = SampleGraph(paths, [r.numel() for r, _, _ in coos], branches=br)
            sg.load(coos)
            sg.capture()
            stage(res, f"whole_sample_graph{br}", lambda: sg.run(coos), cells_per_s=lambda ms: cells / ms * 1e3,
                  samples_per_s=lambda ms: 1e3 / ms)
            res[f"whole_sample_graph{br}"]["eig_fallbacks"] = sg.fallbacks
            stage(res, f"whole_sample_graph{br}_replay_only", lambda: sg.graph.replay())
            del sg
        except Exception as exc:
            res[f"whole_sample_graph{br}"] = dict(error=f"{type(exc).__name__}: {exc}")
            traceback.print_exc()
    return res


def run_c4(args):
    n = [nb for name, nb in synth.hg38_bins(args.c4_res) if name == "chr1"][0]
    dense, seps = synth.device_dense_map([n], seed=7)
    layout = ops.GenomeLayout(seps)
    raw = ops.alloc_map(n)
    raw.copy_(dense)
    del dense
    cells = n * n
    res = {"bins": n, "cells": cells}
    oe, _ = ops.obs_d_exp(raw, layout, args.c4_res, log=True)
    de = ops.norm_de_mtx(oe)                                   # the .de_ode map the scorers read (has_neg)
    del raw
    filt = ops.alloc_map(n)
    stage(res, "gf_s1_main_filter", lambda: ops.main_filter(de, layout, out=filt, decimals=-1),
          gbs_8B_per_cell=lambda ms: 8 * cells / ms / 1e6)
    stage(res, "gf_s1_intra_low_profile", lambda: ops.intra_low_profile(filt, layout),
          gbs_4B_per_cell=lambda ms: 4 * cells / ms / 1e6)
    stage(res, "gf_s1_intra_x_response", lambda: ops.intra_x_response(filt))
    stage(res, "gf_s2_pool_scores", lambda: ops.gf_scores(filt, layout, centers=[n // 2]),
          gbs_4B_per_cell=lambda ms: 4 * cells / ms / 1e6)
    smin, smax = int(n * .05), int(n * .15)
    band_cells = 2 * sum(n - d for d in range(smin, smax))
    stage(res, "checkerboard_get_local", lambda: ops.get_local(de, n), band_cells=band_cells,
          band_cells_per_s=lambda ms: band_cells / ms * 1e3)
    tot = sum(v["ms"] for v in res.values() if isinstance(v, dict) and "ms" in v)
    res["total_ms"] = tot
    res["cells_per_s"] = cells / tot * 1e3 if tot else None
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--c3-res", type=int, default=50000)
    ap.add_argument("--c4-res", type=int, default=10000)
    ap.add_argument("--c3-workers", default="2,4,8", help="HotPathPool sizes to time after the sequential run")
    ap.add_argument("--c3-branches", default="4,8,12", help="SampleGraph branch counts to time")
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    out = {}
    t0 = time.time()
    for name, fn in (("C3", run_c3), ("C4", run_c4)):
        if args.only and args.only != name:
            continue
        try:
            out[name] = fn(args)
        except Exception as exc:
            out[name] = dict(error=f"{type(exc).__name__}: {exc}")
            traceback.print_exc()
        torch.cuda.empty_cache()
        print(name, "done at", round(time.time() - t0, 1), "s", flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/bench_configs.json", "w"), indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
