"""One genome-wide step (config C2) with the profiler range around the kernels of interest, for
`ncu --profile-from-start off`:  default step, then the opt-in variants (register row prep,
upper-triangle eigen pass) so that one capture holds both sides of each comparison."""
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import torch

from hiarch_b200 import synth
from hiarch_b200.pipeline import HotPath, dense_to_coo


def main():
    res = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
    lens = [n for _, n in synth.hg38_bins(res)]
    dense, seps = synth.device_dense_map(lens, seed=0)
    rows, cols, vals = dense_to_coo(dense)
    del dense
    torch.cuda.empty_cache()
    hp = HotPath(seps, res, k_eig=3, max_nnz=rows.numel())
    hp.run_device(rows, cols, vals)                      # warm-up (attributes, allocations)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    hp.run_device(rows, cols, vals)
    torch.cuda.synchronize()
    os.environ["HIA_PREP_REG"] = "1"
    hp.similarity_prep()
    torch.cuda.synchronize()
    os.environ["HIA_PREP_REG"] = "0"
    os.environ["HIA_EIG_UPPER"] = "1"
    hp._eig_ws = {}
    hp.eigen()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print("ok")


main()
