"""CPU: numpy emulation of the loop structure of `oe_apply_upper_kernel` (csrc/hia_oe.cu: 32 x 32
tiles of the upper triangle, 256 threads as (tx, ty), 4 rows per thread, shared tile, transposed
write) against a direct evaluation of the full symmetric map. Written from the kernel source: it
guards the index logic (diagonal tiles, ragged edge tiles, the lower triangle of the input is
NaN-poisoned and must never be read) of a kernel that had not run on a GPU when it was written."""
import numpy as np
def emulate(map_, n, ld, chr_start, bin_chr, inv_expected, inter_inv, n_chr, only_intra, do_log, fill):
    out=np.full((n,ld),np.nan,dtype=np.float32)
    nt=(n+31)//32
    for bi in range(nt):
        for bj in range(nt):
            if bj<bi: continue
            tile=np.zeros((32,33),dtype=np.float32)
            # phase 1
            for tid in range(256):
                tx,ty=tid&31,tid>>5
                c=bj*32+tx
                cb=bin_chr[min(c,n-1)]
                for k in range(4):
                    r=bi*32+ty+8*k
                    v=np.float32(0)
                    if r<n and c<n and c>=r:
                        v=np.float32(map_[r,c])
                        a=bin_chr[r]
                        if a==cb: v=np.float32(v*inv_expected[chr_start[a]+(c-r)])
                        else: v=np.float32(0) if only_intra else np.float32(v*inter_inv[a*n_chr+cb])
                        if do_log: v=np.float32(np.log(np.float32(1)+v)) if v>0 else fill
                        out[r,c]=v
                    tile[ty+8*k,tx]=v
            # phase 2
            for tid in range(256):
                tx,ty=tid&31,tid>>5
                for k in range(4):
                    r=bj*32+ty+8*k; cc=bi*32+tx
                    if r<n and cc<n and r>cc: out[r,cc]=tile[tx,ty+8*k]
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
    for lens in ([70,45,20],[33],[31,1,64]):
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
