"""Where the time of the NormDis step (publish/HiArch/NormDis.py: read .mtx -> preprocess_mtx -> O/E ->
.ode.mtx -> clip / wavelet / median / z-score -> .de_ode.mtx) goes on a mid-size sample
(chr1 + chr2 + chr3 at 100 kb, 6 895 bins): wall time of a warm call + cProfile's top entries.

    python tools/profile_normdis.py [outdir]
"""
import cProfile
import io
import os
import pstats
import sys
import tempfile
import time

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import numpy as np
import torch

from hiarch_b200 import mtx_io, oe_map, synth


def main():
    bins = synth.hg38_bins(100000)[:3]
    names, lens = [c for c, _ in bins], [n for _, n in bins]
    with tempfile.TemporaryDirectory() as tmp:
        t0 = time.perf_counter()
        r, c, v, _ = synth.synth_coo(lens, 3, 0.6)
        prefix = f"{tmp}/s"
        mtx_io.write_mtx(f"{prefix}_normalized.mtx", np.asarray(r), np.asarray(c), np.asarray(v, dtype=np.float64))
        with open(f"{prefix}.window.bed", "w") as fh:
            for (name, s, e, i) in synth.window_table(names, lens, 100000):
                fh.write(f"{name}\t{s}\t{e}\t{i}\n")
        print(f"sample: {sum(lens)} bins, {len(r)} entries, written in {time.perf_counter() - t0:.1f} s", flush=True)
        for it in range(3):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            if it == 2:
                pr = cProfile.Profile()
                pr.enable()
            oe_map.norm_dis(f"{prefix}_normalized.mtx", f"{prefix}.window.bed", f"{tmp}/res{it}", None, "True", "True")
            torch.cuda.synchronize()
            if it == 2:
                pr.disable()
            print(f"norm_dis call {it}: {time.perf_counter() - t0:.3f} s", flush=True)
        s = io.StringIO()
        pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(28)
        print(s.getvalue()[:6000])


main()
