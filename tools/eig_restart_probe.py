"""Development probe: top-k eigen solver on a hard (clustered-spectrum) Pearson map -- the row-sharded
C5 run from synthetic counts needed several restart cycles and lost convergence after an early-stopped
cycle; reproduce on ONE GPU at --res 50000 (61 775 bins) with the single-GPU driver."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import numpy as np
import torch

from hiarch_b200 import ops, synth
from bench_sharded import make_count_rows


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--res", type=int, default=50000)
    ap.add_argument("--depth-scale", type=float, default=0.25, help="scale lambda like a 4x finer binning")
    args = ap.parse_args()
    lens = [nb for _, nb in synth.hg38_bins(args.res)]
    n = sum(lens)
    t0 = time.time()
    counts, seps = make_count_rows(lens, 0, n, 5, "cuda", lam_inter=0.6 * args.depth_scale)
    torch.cuda.synchronize()
    print(f"n={n} counts in {time.time() - t0:.1f}s", flush=True)
    layout = ops.GenomeLayout(seps)
    oe, _ = ops.obs_d_exp(counts, layout, args.res, log=True)
    del counts
    sim = ops.similarity(oe, flavour=ops.SIM_PEARSON)
    del oe
    torch.cuda.synchronize()
    print(f"similarity done {time.time() - t0:.1f}s", flush=True)
    res = {}
    for name, kw in (("default", {}),
                     ("no_early", dict(angle_tol=0.0, tol=1e-6)),
                     ("steps16", dict(steps=16))):
        t1 = time.time()
        w, v, info = ops.topk_eig(sim, k=3, return_info=True, **kw)
        torch.cuda.synchronize()
        res[name] = dict(eigvals=[float(x) for x in w.cpu()], info=info, wall_s=time.time() - t1)
        print(name, json.dumps(res[name]), flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open("gpurun_out/eig_restart_probe.json", "w"), indent=1)


if __name__ == "__main__":
    main()
