"""CPU: numpy emulation of the loop structure of `oe_apply_upper_kernel` (csrc/hia_oe.cu: 64 x 64
tiles of the upper triangle, 256 threads as (tx = float4 column group, ty), 4 rows per thread, all
four 128-bit loads before any store, shared tile, transposed 128-bit write) against a direct
evaluation of the full symmetric map. Written from the kernel source: it guards the index logic
(diagonal tiles, ragged edge tiles and float4 groups that straddle the diagonal or the edge; the
lower triangle of the input is NaN-poisoned and must never reach the output)."""
import numpy as np
T = 64
FAST_TILES = []
def emulate(map_, n, ld, chr_start, bin_chr, inv_expected, inter_inv, n_chr, only_intra, do_log, fill):
    out=np.full((n,ld),np.nan,dtype=np.float32)
    nt=(n+T-1)//T
    for bi in range(nt):
        for bj in range(nt):
            if bj<bi: continue
            tile=np.zeros((T,T+1),dtype=np.float32)
            diag_tile = bi==bj
            # fast path (CTA-uniform): full tile strictly above the diagonal, rows in one chromosome,
            # columns in ONE OTHER chromosome: one factor, no per-cell look-ups or edge predicates
            if bi<bj and bj*T+T<=n:
                ra,rb=bin_chr[bi*T],bin_chr[bi*T+T-1]; ca,cbl=bin_chr[bj*T],bin_chr[bj*T+T-1]
                if ra==rb and ca==cbl and ra!=ca:
                    FAST_TILES.append((bi,bj))
                    f=np.float32(0) if only_intra else np.float32(inter_inv[ra*n_chr+ca])
                    for tid in range(256):
                        tx,ty=tid&15,tid>>4
                        c0=bj*T+4*tx
                        for k in range(4):
                            r=bi*T+ty+16*k
                            for j in range(4):
                                x=np.float32(np.float32(map_[r,c0+j])*f)
                                if do_log: x=np.float32(np.log(np.float32(1)+x)) if x>0 else fill
                                out[r,c0+j]=x; tile[ty+16*k,4*tx+j]=x
                    for tid in range(256):
                        tx,ty=tid&15,tid>>4
                        for k in range(4):
                            for j in range(4): out[bj*T+ty+16*k,bi*T+4*tx+j]=tile[4*tx+j,ty+16*k]
                    continue
            # phase 1
            for tid in range(256):
                tx,ty=tid&15,tid>>4
                c0=bj*T+4*tx
                cb=[bin_chr[min(c0+j,n-1)] for j in range(4)]
                for k in range(4):
                    r=bi*T+ty+16*k
                    v=[np.float32(0)]*4
                    if r<n and c0<n:
                        a=bin_chr[r]; s=chr_start[a]
                        assert c0+3<ld                      # the 128-bit load stays inside the padded row
                        raw=[np.float32(map_[r,c0+j]) for j in range(4)]
                        for j in range(4):
                            c=c0+j
                            if c<n and c>=r:
                                x=raw[j]
                                if cb[j]==a: x=np.float32(x*inv_expected[s+(c-r)])
                                else: x=np.float32(0) if only_intra else np.float32(x*inter_inv[a*n_chr+cb[j]])
                                if do_log: x=np.float32(np.log(np.float32(1)+x)) if x>0 else fill
                                v[j]=x
                        if c0>=r and c0+3<n:
                            for j in range(4): out[r,c0+j]=v[j]       # one float4 store
                        else:
                            for j in range(4):
                                if c0+j<n and c0+j>=r: out[r,c0+j]=v[j]
                    for j in range(4): tile[ty+16*k,4*tx+j]=v[j]
            # phase 2
            for tid in range(256):
                tx,ty=tid&15,tid>>4
                for k in range(4):
                    R=bj*T+ty+16*k; C0=bi*T+4*tx
                    if R>=n or C0>=n: continue
                    v=[tile[4*tx+j,ty+16*k] for j in range(4)]
                    if not diag_tile:
                        assert C0+3<n and C0+3<R
                        for j in range(4): out[R,C0+j]=v[j]
                    else:
                        for j in range(4):
                            if C0+j<R: out[R,C0+j]=v[j]
    return out
def direct(full, n, chr_start, bin_chr, inv_expected, inter_inv, n_chr, only_intra, do_log, fill):
    out=np.zeros((n,n),dtype=np.float32)
    for r in range(n):
        for c in range(n):
            v=np.float32(full[r,c]); a=bin_chr[r]; b=bin_chr[c]
            if a==b: v=np.float32(v*inv_expected[chr_start[a]+abs(c-r)])
            else: v=np.float32(0) if only_intra else np.float32(v*inter_inv[a*n_chr+b])
            if do_log: v=np.float32(np.log(np.float32(1)+v)) if v>0 else fill
            out[r,c]=v
    return out
def test_upper_apply_kernel_index_logic():
    rng = np.random.default_rng(0)
    for lens in ([70,45,20],[33],[31,1,64],[64,64],[130,2,61],[64,128,70]):
        n=sum(lens); ld=(n+3)//4*4
        chr_start=np.concatenate([[0],np.cumsum(lens)]).astype(int)
        bin_chr=np.repeat(np.arange(len(lens)),lens)
        up=np.triu(rng.poisson(3.0,(n,n)).astype(np.float32))
        full=up+np.triu(up,1).T
        poisoned=np.where(np.arange(n)[:,None]>np.arange(n)[None,:], np.nan, full).astype(np.float32)
        m=np.full((n,ld),np.nan,dtype=np.float32); m[:,:n]=poisoned
        inv_expected=rng.random(n).astype(np.float32)+0.1
        C=len(lens); ii=rng.random((C,C)).astype(np.float32)+0.1; ii=(ii+ii.T)/2; inter_inv=ii.reshape(-1)
        for only_intra in (False,True):
            for do_log in (False,True):
                fill=np.float32(0.123)
                got=emulate(m,n,ld,chr_start,bin_chr,inv_expected,inter_inv,C,only_intra,do_log,fill)[:,:n]
                ref=direct(full,n,chr_start,bin_chr,inv_expected,inter_inv,C,only_intra,do_log,fill)
                assert not np.isnan(got).any(), (lens,only_intra,do_log)
                assert np.array_equal(got,ref), (lens,only_intra,do_log,np.abs(got-ref).max())
    assert (0, 1) in FAST_TILES and (0, 3) in FAST_TILES       # [64,128,70]: tiles that take the fast path


# ---------------------------------------------------------------------------------------------
# symm_upper_kernel + symm_upper_reduce_kernel (csrc/hia_eig.cu): y = C v from the upper triangle
# only. The emulation follows the kernel's work decomposition -- (column block J, row segment) items,
# the early exit, r_end, the float4 groups of a lane with their three masks (below the diagonal,
# diagonal cell for the column part, beyond the edge), the partial buffers P[J][r] / Pc[seg][c] and
# the index ranges of the reduce kernel -- with NaN-poisoned partial buffers and a NaN-poisoned lower
# triangle: any cell read that the kernel must not read, or any partial read that was never
# written, shows up as NaN.
# ---------------------------------------------------------------------------------------------
UP_COLS, UP_SEG, UP_WARPS = 128, 1024, 8


def emulate_symm_upper(c, v):
    n, b = v.shape
    nJ, nSeg = -(-n // UP_COLS), -(-n // UP_SEG)
    prow = np.full((nJ, n, b), np.nan)
    pcol = np.full((nSeg, n, b), np.nan)
    for J in range(nJ):
        for seg in range(nSeg):
            c0, r_seg = J * UP_COLS, seg * UP_SEG
            if r_seg > c0 + UP_COLS - 1:
                continue
            r_end = min(n, r_seg + UP_SEG, c0 + UP_COLS)
            cacc = np.zeros((UP_COLS, b))
            for r0 in range(r_seg, r_end, 8):                       # a warp's group of 8 rows
                for r in range(r0, r0 + 8):
                    racc = np.zeros(b)
                    for lane in range(32):
                        gc = c0 + 4 * lane
                        if not (r < r_end and gc < n and gc + 3 >= r):
                            continue                                 # no load at all
                        e = np.array([c[r, gc + q] if gc + q < n else 0.0 for q in range(4)])
                        for q in range(4):
                            if gc + q < r:
                                e[q] = 0.0                           # below the diagonal (NaN-poisoned in the input)
                        for q in range(4):
                            if gc + q < n:
                                racc += e[q] * v[gc + q]
                        for q in range(4):
                            if gc + q == r:
                                e[q] = 0.0                           # the diagonal cell: row part only
                        for q in range(4):
                            cacc[4 * lane + q] += e[q] * v[r]
                    if r < r_end:
                        prow[J, r] = racc
            for cl in range(UP_COLS):
                if c0 + cl < n:
                    pcol[seg, c0 + cl] = cacc[cl]
    y = np.zeros((n, b))
    for r in range(n):
        for J in range(r // UP_COLS, nJ):
            y[r] += prow[J, r]
        for seg in range(0, min(r, (r // UP_COLS) * UP_COLS + UP_COLS - 1) // UP_SEG + 1):
            y[r] += pcol[seg, r]
    return y


def test_symm_upper_pass_index_logic():
    rng = np.random.default_rng(3)
    for n in (5, 127, 128, 129, 300, 1024 + 130, 2 * 1024 + 7):
        b = 4
        up = np.triu(rng.standard_normal((n, n)))
        full = up + np.triu(up, 1).T
        poisoned = np.where(np.arange(n)[:, None] > np.arange(n)[None, :], np.nan, full)
        v = rng.standard_normal((n, b))
        with np.errstate(invalid="ignore"):
            y = emulate_symm_upper(poisoned, v)
        assert not np.isnan(y).any(), n
        assert np.abs(y - full @ v).max() < 1e-9, n


# ---------------------------------------------------------------------------------------------
# Numerics of the similarity path on SPARSE high-resolution rows (csrc/hia_syrk.cu), emulated in numpy
# from the kernel source: (1) row_prep_kernel's statistics -- fp32 inside a float4, fp64 across float4s,
# shifted by a 32-cell mean estimate -- and its fp32 unit row with a two-float mean; (2) the epilogue's
# fixed-point fold of the K-chunk partials. Both guard regressions that cost > 1e-5 at 25 kb on the GPU
#  and cannot be seen on the small dense maps of the other tests.
# ---------------------------------------------------------------------------------------------
def _sparse_rows(n, seed, first):
    rng = np.random.default_rng(seed)
    base = rng.poisson(0.04, n) * (1.0 + rng.random(n))            # shared structure -> |rho| large
    out = []
    for _ in range(2):
        x = np.log1p(base + rng.poisson(0.01, n) * rng.random(n))
        x[x == 0] = np.log1p(0.0371)                                # the repeated background value
        x[0] = first                                                # the first cell is a count cell
        out.append(x.astype(np.float32))
    return out


def _prep_emulated(x, shift_mode):
    n = x.size
    f32 = np.float32
    if shift_mode == "sampled":
        idx = (np.arange(32, dtype=np.int64) * n) >> 5
        shift = f32(np.float32(x[idx].astype(np.float32).sum(dtype=np.float32)) * f32(1.0 / 32.0))
    else:
        shift = x[0]
    a = (x - shift).astype(f32)
    k4 = n // 4
    g = a[:k4 * 4].reshape(-1, 4)
    p1 = ((g[:, 0] + g[:, 1]).astype(f32) + (g[:, 2] + g[:, 3]).astype(f32)).astype(f32)
    sq = (g * g).astype(f32)
    p2 = ((sq[:, 0] + sq[:, 1]).astype(f32) + (sq[:, 2] + sq[:, 3]).astype(f32)).astype(f32)
    tail = a[k4 * 4:]
    t1 = p1.astype(np.float64).sum() + tail.astype(np.float64).sum()
    t2 = p2.astype(np.float64).sum() + (tail * tail).astype(f32).astype(np.float64).sum()
    ms = t1 / n
    ss = t2 - 2.0 * ms * t1 + n * ms * ms
    mean = ms + float(shift)
    inv = 1.0 / np.sqrt(ss)
    mh = f32(mean)
    ml = f32(mean - float(mh))
    invs = f32(inv * 32768.0)
    xs = (((x - mh).astype(f32) - ml).astype(f32) * invs).astype(f32)
    hi = xs.astype(np.float16)
    lo = (xs - hi.astype(f32)).astype(f32).astype(np.float16)
    return hi.astype(np.float64), lo.astype(np.float64)


def _exact_unit(x):
    d = x.astype(np.float64)
    d = d - d.mean()
    return d / np.sqrt((d * d).sum())


def test_row_prep_numerics_on_sparse_rows():
    n = 123541
    worst = {"sampled": 0.0, "first_cell": 0.0}
    for first in (2.5, 1.9, 3.3, 4.1, 2.2, 2.9):                    # the rounding of (background - shift)^2 is a lottery
        xa, xb = _sparse_rows(n, 3, first)
        rho = float(_exact_unit(xa) @ _exact_unit(xb))
        assert abs(rho) > 0.5
        for mode in worst:
            (ha, la), (hb, lb) = _prep_emulated(xa, mode), _prep_emulated(xb, mode)
            got = float((ha @ hb + ha @ lb + la @ hb) * 2.0 ** -30)  # exact products of the planes
            worst[mode] = max(worst[mode], abs(got - rho))
    assert worst["sampled"] <= 2e-7, worst                          # GPU, 25 kb: 2.7e-8 mean / 1.0e-7 max
    assert worst["first_cell"] >= 3 * worst["sampled"], worst       # what the first version did (GPU: 8e-7 / 3.6e-6)


def test_fixed_point_fold_keeps_the_small_partials():
    """The chunk partials of one cell between neighbouring rows of a sparse 25 kb map (k_chunk 128 ... 512 at
    K = 123 541): a few chunks near the diagonal carry almost all of the sum, hundreds of small, nearly
    equal background partials follow. A sequential fp32 fold rounds each of those away (they are below half
    an ulp of the sum); the int32 fold in units of 2^-30 that the epilogue does keeps them."""
    rng = np.random.default_rng(0)
    worst32, worst_fix = 0.0, 0.0
    for rho, n_chunks, small in ((0.93, 965, 14.8), (0.61, 965, 20.8), (-0.77, 241, -25.2), (0.66, 241, 31.6)):
        big = np.array([0.2, 0.5, 0.3]) * rho * 2.0 ** 30             # the planes' 2^15 x 2^15 scale
        part = np.concatenate([big, small * (1.0 + 0.02 * rng.standard_normal(n_chunks - 3))])
        exact = part.sum() * 2.0 ** -30
        acc = np.float32(0.0)
        fx = 0
        for p in part.astype(np.float32):
            acc = np.float32(acc + p)
            fx += int(np.rint(p))                                    # __float2int_rn(partial): exact integer adds
        worst32 = max(worst32, abs(float(acc) * 2.0 ** -30 - exact))
        worst_fix = max(worst_fix, abs(float(np.float32(fx)) * 2.0 ** -30 - exact))
    assert worst_fix <= 3.5e-7, worst_fix          # <= half a unit of 2^-30 per partial when they all round the same way
    assert worst32 >= 3e-6, worst32                # the stagnation the first version of the epilogue had


# ---------------------------------------------------------------------------------------------
# fill_rows_fast_kernel (csrc/hia_scatter.cu): the index logic of one (row, segment) CTA, emulated from the
# kernel source -- the entries of the row split into a scalar head up to the first 16-byte aligned entry,
# aligned groups of four and a scalar tail; plain stores while the columns increase strictly, the segment
# redone with adds otherwise; the segment buffer shifted by the 16-byte phase of the output segment and
# written out in aligned float4 groups with masked edges. NaN-poisoned output: a cell the kernel must
# write and does not, or an entry it reads outside its row, shows up.
# ---------------------------------------------------------------------------------------------
SEG = 16384


def emulate_fill_rows_fast(cols, vals, row_ptr, n, ld, triu, seg=SEG, threads=512):
    out = np.full((n, ld), np.nan, dtype=np.float32)
    flat = out.reshape(-1)
    segs = -(-n // seg)
    for block in range(n * segs):
        r = block // segs
        c_lo = r if triu else 0
        seg_lo = c_lo + (block - r * segs) * seg
        if seg_lo >= n:
            continue
        seg_n = min(seg, n - seg_lo)
        e0, e1 = int(row_ptr[r]), int(row_ptr[r + 1])
        off = (r * ld + seg_lo) & 3
        nq = (off + seg_n + 3) >> 2
        buf = np.zeros(4 * nq, dtype=np.float32)
        rc, rv = cols[e0:e1], vals[e0:e1]
        length = e1 - e0
        h = min(length, (4 - (e0 & 3)) & 3)
        ng = (length - h) >> 2
        t0 = h + 4 * ng
        bad = False
        seen = np.zeros(length, dtype=bool)                     # every entry of the row is visited exactly once

        def put(prev, c, v):
            nonlocal bad
            bad |= c <= prev
            c -= seg_lo
            if 0 <= c < seg_n:
                buf[off + c] = v

        def before(e):
            return int(rc[e - 1]) if e > 0 else -1

        for tid in range(min(threads, 4)):                      # head: tid < h
            if tid < h:
                put(before(tid), int(rc[tid]), rv[tid]); seen[tid] = True
        for g in range(ng):                                     # body (any thread / iteration order: distinct cells)
            ea = h + 4 * g
            assert (e0 + ea) % 4 == 0                           # the 128-bit loads are aligned
            prev = before(ea)
            for j in range(4):
                put(prev, int(rc[ea + j]), rv[ea + j]); seen[ea + j] = True
                prev = int(rc[ea + j])
        for tid in range(min(threads, 4)):                      # tail: t0 + tid < len
            if t0 + tid < length:
                put(before(t0 + tid), int(rc[t0 + tid]), rv[t0 + tid]); seen[t0 + tid] = True
        assert seen.all() and length - t0 < 4
        if bad:
            buf[:] = 0
            for e in range(length):
                c = int(rc[e]) - seg_lo
                if 0 <= c < seg_n:
                    buf[off + c] += rv[e]
        base = r * ld + seg_lo - off
        assert base % 4 == 0                                    # aligned 128-bit stores
        for q in range(nq):
            i0 = 4 * q - off
            for j in range(4):
                if 0 <= i0 + j < seg_n:
                    flat[base + 4 * q + j] = buf[4 * q + j]
    return out


def test_fill_rows_fast_index_logic():
    rng = np.random.default_rng(1)
    for n, ld, seg in ((37, 40, 16), (64, 64, 16), (50, 52, 64), (33, 36, 8)):
        for triu in (True, False):
            dense = np.zeros((n, n), dtype=np.float32)
            rows_l, cols_l, vals_l = [], [], []
            for r in range(n):
                k = int(rng.integers(0, n))
                cs = np.sort(rng.choice(n, size=k, replace=False)) if r % 5 else rng.integers(0, n, size=k)   # some rows unsorted / with duplicates
                for c in cs:
                    v = np.float32(rng.integers(1, 9))
                    rows_l.append(r); cols_l.append(int(c)); vals_l.append(v)
                    if not triu or c >= r:
                        dense[r, c] += v
            cols = np.array(cols_l, dtype=np.int32); vals = np.array(vals_l, dtype=np.float32)
            row_ptr = np.searchsorted(np.array(rows_l), np.arange(n + 1)).astype(np.int64)
            got = emulate_fill_rows_fast(cols, vals, row_ptr, n, ld, triu, seg=seg)
            lo = np.arange(n)[:, None] > np.arange(n)[None, :]
            if triu:                                            # a 'triu' row starts on the diagonal: cells left of it are not written
                assert np.isnan(got[:, :n][lo]).all()
                assert np.array_equal(got[:, :n][~lo], dense[~lo]), (n, ld, seg)
            else:
                assert np.array_equal(got[:, :n], dense), (n, ld, seg)
            assert np.isnan(got[:, n:]).all()                   # the padding columns are never touched


# ---------------------------------------------------------------------------------------------
# dwt2_fwd_kernel / dwt2_inv_kernel (csrc/hia_wavelet.cu): the tile index logic, emulated from the kernel
# source against oracle/wavelet_oracle.py -- forward: a 36 x 36 input tile at (2 k0 - 4) with the symmetric
# index reflection (repeated for blocks shorter than the halo), tile row 2 kl + 5 - j for tap j; inverse:
# 18 x 18 coefficients per band from (n0 / 2), taps 4 + par - 2 q for the three coefficients of an output
# sample, coefficients beyond the band = 0, output cut to the forward size of the level.
# ---------------------------------------------------------------------------------------------
def _sym_index(i, n):
    p = 2 * n
    i = i % p                                    # the kernel adds p when C's % went negative
    return i if i < n else p - 1 - i


def emulate_dwt2_fwd(x):
    from oracle import wavelet_oracle as wo
    rows, cols = x.shape
    orows, ocols = wo.dwt_len(rows), wo.dwt_len(cols)
    bands = [np.full((orows, ocols), np.nan) for _ in range(4)]
    FT, FIN = 16, 36
    lo, hi = wo.DB3_DEC_LO, wo.DB3_DEC_HI
    for by in range(-(-orows // FT)):
        for bx in range(-(-ocols // FT)):
            k0r, k0c = by * FT, bx * FT
            tile = np.array([[x[_sym_index(2 * k0r - 4 + tr, rows), _sym_index(2 * k0c - 4 + tc, cols)]
                              for tc in range(FIN)] for tr in range(FIN)])
            s_lo = np.zeros((FT, FIN)); s_hi = np.zeros((FT, FIN))
            for kl in range(FT):
                for j in range(6):
                    s_lo[kl] += lo[j] * tile[2 * kl + 5 - j]
                    s_hi[kl] += hi[j] * tile[2 * kl + 5 - j]
            for kr in range(FT):
                for kc in range(FT):
                    gr, gc = k0r + kr, k0c + kc
                    if gr < orows and gc < ocols:
                        l = s_lo[kr, [2 * kc + 5 - j for j in range(6)]]
                        h = s_hi[kr, [2 * kc + 5 - j for j in range(6)]]
                        bands[0][gr, gc] = lo @ l; bands[1][gr, gc] = hi @ l
                        bands[2][gr, gc] = lo @ h; bands[3][gr, gc] = hi @ h
    return dict(zip(("aa", "ad", "da", "dd"), bands))


def emulate_dwt2_inv(c, rows, cols):
    from oracle import wavelet_oracle as wo
    crows, ccols = c["aa"].shape
    out = np.full((rows, cols), np.nan)
    IT, IK = 32, 18
    g, gh = wo.DB3_REC_LO, wo.DB3_REC_HI
    names = ("aa", "ad", "da", "dd")
    for by in range(-(-rows // IT)):
        for bx in range(-(-cols // IT)):
            n0r, n0c = by * IT, bx * IT
            s = np.zeros((4, IK, IK))
            for tr in range(IK):
                for tc in range(IK):
                    gr, gc = n0r // 2 + tr, n0c // 2 + tc
                    if gr < crows and gc < ccols:
                        for b in range(4):
                            s[b, tr, tc] = c[names[b]][gr, gc]
            s_lo = np.zeros((IK, IT)); s_hi = np.zeros((IK, IT))
            for tr in range(IK):
                for nl in range(IT):
                    p, par = nl >> 1, nl & 1
                    for q in range(3):
                        t = 4 + par - 2 * q
                        s_lo[tr, nl] += s[0, tr, p + q] * g[t] + s[1, tr, p + q] * gh[t]
                        s_hi[tr, nl] += s[2, tr, p + q] * g[t] + s[3, tr, p + q] * gh[t]
            for ml in range(IT):
                for nl in range(IT):
                    gr, gc = n0r + ml, n0c + nl
                    if gr >= rows or gc >= cols:
                        continue
                    p, par = ml >> 1, ml & 1
                    out[gr, gc] = sum(s_lo[p + q, nl] * g[4 + par - 2 * q] + s_hi[p + q, nl] * gh[4 + par - 2 * q]
                                      for q in range(3))
    return out


def test_wavelet_tile_index_logic():
    from oracle import wavelet_oracle as wo
    rng = np.random.default_rng(2)
    for shape in ((37, 50), (64, 33), (3, 41), (1, 9), (35, 2)):
        x = rng.standard_normal(shape)
        got = emulate_dwt2_fwd(x)
        ref = wo.dwt2(x)
        for k in ref:
            assert not np.isnan(got[k]).any(), (shape, k)
            assert np.abs(got[k] - ref[k]).max() < 1e-12, (shape, k)
        back = emulate_dwt2_inv(ref, *shape)              # the level's approximation, cut to its forward size
        assert not np.isnan(back).any(), shape
        assert np.abs(back - x).max() < 1e-12, shape      # = waverecn's trim + perfect reconstruction
