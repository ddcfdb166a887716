"""Seeded synthetic Hi-C contact maps (power-law distance decay + compartments).

The generator SURVEY.md section 8(d) defines: bins from hg38 chromosome lengths at a given
resolution; contact rate  lambda(i,j) = A (d+1)^-alpha (1 + beta c_i c_j)  inside a chromosome
(d = bin distance) and  lambda_inter (1 + beta c_i c_j)  between chromosomes, with c a smoothed
random +-1 compartment track; counts ~ Poisson(lambda); only upper-triangle non-zeros are kept.
Files are written in the reference's wire format: `<sample>_normalized.mtx` = tab separated
"row col value" with "%d\t%d\t%.2f" (publish/new_hic_class.py:2267-2299) and a 4-column
`.window.bed` (chr, start, end, index; publish/new_hic_class.py:63-68).

Host (numpy) generation is for test-sized maps; `device_dense_map` builds the large
genome-wide maps directly in HBM (a 10^9-line text file is not a sensible bench input).
"""
import numpy as np

HG38_LENS = {
    "chr1": 248956422, "chr2": 242193529, "chr3": 198295559, "chr4": 190214555,
    "chr5": 181538259, "chr6": 170805979, "chr7": 159345973, "chr8": 145138636,
    "chr9": 138394717, "chr10": 133797422, "chr11": 135086622, "chr12": 133275309,
    "chr13": 114364328, "chr14": 107043718, "chr15": 101991189, "chr16": 90338345,
    "chr17": 83257441, "chr18": 80373285, "chr19": 58617616, "chr20": 64444167,
    "chr21": 46709983, "chr22": 50818468, "chrX": 156040895, "chrY": 57227415,
}

A_INTRA, ALPHA, BETA = 2000.0, 1.08, 0.5


def hg38_bins(resolution, chroms=None):
    """[(chrom, n_bins)] with n_bins = ceil(len / resolution)."""
    chroms = list(HG38_LENS) if chroms is None else chroms
    return [(c, -(-HG38_LENS[c] // resolution)) for c in chroms]


def chr_seps_from_lens(lens):
    """[(start, end_inclusive)] like GenomeIndex.chr_seps (publish/new_hic_class.py:95-97)."""
    seps, s = [], 0
    for n in lens:
        seps.append((s, s + n - 1))
        s += n
    return seps


def compartment_track(n, rng, smooth=25):
    """Smoothed random +-1 track in [-1, 1]."""
    raw = rng.choice([-1.0, 1.0], size=n + 2 * smooth)
    ker = np.ones(2 * smooth + 1) / (2 * smooth + 1)
    sm = np.convolve(raw, ker, mode="valid")[:n]
    return np.sign(sm + 1e-12) * np.minimum(1.0, np.abs(sm) * 3)


def rate_matrix(lens, seed=0, lam_inter=0.6, smooth=None, x_shape=0.0, depth=1.0):
    """Dense fp64 rate matrix lambda for chromosomes of `lens` bins (host, small sizes).
    x_shape > 0 adds a Rabl-like X pattern (anti-diagonal band through the chromosome centre)
    so that the `intra_x` anchor method has something to find."""
    rng = np.random.default_rng(seed)
    n = int(sum(lens))
    seps = chr_seps_from_lens(lens)
    comp = np.concatenate([compartment_track(L, rng, smooth or max(3, L // 40)) for L in lens])
    cc = 1.0 + BETA * np.outer(comp, comp)
    lam = lam_inter * cc
    for (s, e) in seps:
        L = e - s + 1
        d = np.abs(np.subtract.outer(np.arange(L), np.arange(L)))
        blk = A_INTRA * (d + 1.0) ** (-ALPHA) * cc[s:e + 1, s:e + 1]
        if x_shape > 0:
            i = np.arange(L)
            anti = np.abs(np.add.outer(i, i) - (L - 1))
            blk = blk + x_shape * np.exp(-(anti / max(2.0, 0.03 * L)) ** 2)
        lam[s:e + 1, s:e + 1] = blk
    return lam * depth, seps


def sample_counts(lam, seed=0, jitter=True):
    """Poisson counts, symmetric, optional ICE-like float jitter (values then have 2 decimals)."""
    rng = np.random.default_rng(seed + 7919)
    up = np.triu(rng.poisson(np.triu(lam)).astype(np.float64))
    if jitter:
        up = np.round(up * (1.0 + 0.05 * np.triu(rng.standard_normal(lam.shape))), 2)
        up[up < 0] = 0.0
    return up                                  # upper triangle incl. diagonal


def synth_coo(lens, seed=0, lam_inter=0.6, jitter=True, x_shape=0.0, depth=1.0):
    """(rows, cols, vals, seps): upper-triangle non-zeros of a seeded synthetic map."""
    lam, seps = rate_matrix(lens, seed, lam_inter, x_shape=x_shape, depth=depth)
    up = sample_counts(lam, seed, jitter)
    r, c = np.nonzero(up)
    return r.astype(np.int32), c.astype(np.int32), up[r, c], seps


def window_table(names, lens, bin_size):
    """Rows (chr, start, end, index) of a .window.bed."""
    rows, idx = [], 0
    for name, L in zip(names, lens):
        for b in range(L):
            rows.append((name, b * bin_size, (b + 1) * bin_size, idx))
            idx += 1
    return rows


def write_sample(prefix, names, lens, bin_size, seed=0, lam_inter=0.6, x_shape=0.0, depth=1.0):
    """Write `<prefix>_normalized.mtx` + `<prefix>.window.bed` in the reference's format."""
    r, c, v, seps = synth_coo(lens, seed, lam_inter, x_shape=x_shape, depth=depth)
    np.savetxt(f"{prefix}_normalized.mtx", np.vstack([r, c, v]).T, fmt="%d\t%d\t%.2f")
    with open(f"{prefix}.window.bed", "w") as fh:
        for (name, s, e, i) in window_table(names, lens, bin_size):
            fh.write(f"{name}\t{s}\t{e}\t{i}\n")
    return f"{prefix}_normalized.mtx", f"{prefix}.window.bed"


COMPONENT_BETAS = (0.5, 0.3, 0.2)       # A/B compartments + two weaker sub-compartment tracks


def device_dense_map(lens, seed=0, lam_inter=0.6, device="cuda", dtype=None,
                     betas=COMPONENT_BETAS):
    """Large maps: symmetric dense fp32 Poisson map built directly in HBM (torch RNG stream,
    `torch.Generator(device).manual_seed(seed)`), one chromosome-pair block at a time so the
    peak footprint stays ~2x the map. The rate is
        lambda = base(d) * (1 + sum_m beta_m c_m[i] c_m[j])
    with three compartment-like tracks of decreasing strength and smoothing scale, so that the
    top-3 eigenvectors of the correlation map are well separated (like sub-compartments in
    real maps). Returns (dense [N,N] fp32 'all'-type, seps)."""
    import torch
    dtype = dtype or torch.float32
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    n = int(sum(lens))
    seps = chr_seps_from_lens(lens)
    rng = np.random.default_rng(seed)
    comps = []
    for m, beta in enumerate(betas):
        track = np.concatenate([compartment_track(L, rng, max(3, L // (40 * (m + 1)))) for L in lens])
        comps.append(np.sqrt(beta) * track)
    comp = torch.from_numpy(np.stack(comps, axis=1)).to(device=device, dtype=dtype)   # [N, M]
    out = torch.empty((n, n), device=device, dtype=dtype)
    for a, (s1, e1) in enumerate(seps):
        ca = comp[s1:e1 + 1]
        for b, (s2, e2) in enumerate(seps):
            if b < a:
                continue
            cb = comp[s2:e2 + 1]
            cc = 1.0 + ca @ cb.T
            if a == b:
                L = e1 - s1 + 1
                i = torch.arange(L, device=device, dtype=dtype)
                d = (i[:, None] - i[None, :]).abs()
                lam = A_INTRA * (d + 1.0) ** (-ALPHA) * cc
                cnt = torch.poisson(torch.triu(lam), generator=gen)
                cnt = torch.triu(cnt) + torch.triu(cnt, 1).T
                out[s1:e1 + 1, s2:e2 + 1] = cnt
            else:
                cnt = torch.poisson(lam_inter * cc, generator=gen)
                out[s1:e1 + 1, s2:e2 + 1] = cnt
                out[s2:e2 + 1, s1:e1 + 1] = cnt.T
    return out, seps


def _hash_uniform(key, salt):
    """Counter-based uniform in (0, 1) from an int64 key (splitmix-like; wraps mod 2^64)."""
    z = key + salt
    z = (z ^ (z >> 30)) * -4658895280553007687          # 0xBF58476D1CE4E5B9 as int64
    z = (z ^ (z >> 27)) * -7723592293110705685          # 0x94D049BB133111EB
    z = z ^ (z >> 31)
    return ((z >> 11) & ((1 << 53) - 1)).double() * (1.0 / (1 << 53)) * (1 - 2e-16) + 1e-16


def device_count_rows(lens, row0, rows, seed, device, lam_inter=0.6, chunk=1024):
    """rows x n block of a SYMMETRIC synthetic count map (SURVEY 8d generator): the sample of cell
    (i, j) is a deterministic function of (min(i,j), max(i,j), seed), so every rank generates a
    consistent slice without any rank holding the whole map. Poisson by inverse CDF (lambda <= 12)
    or the rounded normal approximation (lambda > 12)."""
    import torch
    n = int(sum(lens))
    seps = chr_seps_from_lens(lens)
    rng = np.random.default_rng(seed)
    comps = np.stack([np.sqrt(b) * np.concatenate([compartment_track(L, rng, max(3, L // (40 * (m + 1))))
                                                   for L in lens]) for m, b in enumerate(COMPONENT_BETAS)], axis=1)
    comp = torch.from_numpy(comps).to(device=device, dtype=torch.float32)
    bin_chr = torch.from_numpy(np.concatenate([np.full(L, c, dtype=np.int64) for c, L in enumerate(lens)])).to(device)
    x = torch.empty((max(rows, 1), (n + 3) // 4 * 4), dtype=torch.float32, device=device)[:rows, :n]
    jj = torch.arange(n, device=device, dtype=torch.int64)
    for r in range(0, rows, chunk):
        r1 = min(rows, r + chunk)
        ii = torch.arange(row0 + r, row0 + r1, device=device, dtype=torch.int64)
        lo_ = torch.minimum(ii[:, None], jj[None, :])
        hi_ = torch.maximum(ii[:, None], jj[None, :])
        key = lo_ * n + hi_
        cc = 1.0 + comp[row0 + r:row0 + r1] @ comp.T
        same = bin_chr[row0 + r:row0 + r1, None] == bin_chr[None, :]
        d = (hi_ - lo_).float()
        lam = torch.where(same, A_INTRA * (d + 1.0) ** (-ALPHA) * cc, lam_inter * cc).double()
        u = _hash_uniform(key, seed * 7919 + 1)
        # small lambda: inverse CDF
        k = torch.zeros_like(lam)
        p = torch.exp(-lam)
        cdf = p.clone()
        for t in range(1, 40):
            k += (u > cdf).double()
            p = p * lam / t
            cdf += p
        # large lambda: normal approximation with a second hashed uniform (Box-Muller)
        u2 = _hash_uniform(key, seed * 7919 + 2)
        z = torch.sqrt(-2.0 * torch.log(u)) * torch.cos(2 * np.pi * u2)
        big = torch.clamp(torch.round(lam + torch.sqrt(lam) * z), min=0.0)
        x[r:r1] = torch.where(lam > 12.0, big, k).float()
    return x, seps
