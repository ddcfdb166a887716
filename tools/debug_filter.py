import sys, os
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import numpy as np, torch
from hiarch_b200 import ops, synth
from oracle import hia_oracle as O
g = dict(np.load("tests/golden/g3chr.npz"))
n = int(g["in.lens"].sum()); seps = synth.chr_seps_from_lens([int(x) for x in g["in.lens"]])
layout = ops.GenomeLayout(seps)
dev = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a)).to("cuda", dt)
dense = ops.scatter_coo_sym(dev(g["de_file.rows"], torch.int32), dev(g["de_file.cols"], torch.int32), dev(g["de_file.vals"], torch.float32), n, fill_mode=ops.FILL_MIN, layout=layout)
d64 = O.scatter_coo_sym(g["de_file.rows"], g["de_file.cols"], g["de_file.vals"], n, has_neg=True)
print("dense diff", np.abs(dense.cpu().numpy() - d64).max())
scratch = ops.alloc_map(n); out = ops.alloc_map(n)
for (s1, e1) in seps:
    size = int((e1 - s1 + 1) * .1)
    for (s2, e2) in seps:
        src = dense[s1:e1+1, s2:e2+1]; dst = out[s1:e1+1, s2:e2+1]
        ops.box_filter_reflect(src, size, out=dst, scratch=scratch[s1:e1+1, s2:e2+1])
        ref = O.box_filter_reflect(d64[s1:e1+1, s2:e2+1], size)
        b = dst.cpu().numpy().astype(np.float64)
        e_box = np.abs(b - ref).max()
        p = ops.percentiles(dst, [66, 34]).cpu().numpy()
        pref = [np.percentile(ref, 66), np.percentile(ref, 34)]
        core = ref[(ref < pref[0]) & (ref > pref[1])]
        zr = O.norm_to_zero_mean(ref)
        tmp = ops.alloc_map(dst.shape[0], cols=dst.shape[1]); tmp.copy_(dst)
        z = ops.norm_to_zero_mean(tmp).cpu().numpy()
        print(f"blk ({s1},{s2}) size {size}: box err {e_box:.2e}  pct {p} ref {pref}  core mean {core.mean():.6f} std {core.std():.3e} n {core.size}  z err {np.abs(z - zr).max():.2e}")
