"""CPU: f2 -- the native .mtx reader / writer against pandas.read_table(+dropna) and numpy.savetxt,
i.e. against exactly what the reference calls (publish/new_hic_class.py:2580-2581, :2290-2297)."""
import os

import numpy as np
import pandas as pd
import pytest

from hiarch_b200 import mtx_io


def _pandas_read(path):
    df = pd.read_table(path, names=['row', 'col', 'value'])
    df.dropna(inplace=True)
    return df['row'].values.astype(np.int64), df['col'].values.astype(np.int64), df['value'].values.astype(np.float64)


def _same_as_pandas(path):
    r, c, v = mtx_io.read_mtx(path)
    pr, pc, pv = _pandas_read(path)
    assert np.array_equal(r, pr) and np.array_equal(c, pc)
    assert np.array_equal(v, pv)                      # bit for bit, not allclose
    return r, c, v


def test_reader_matches_pandas_on_pipeline_files(tmp_path):
    rng = np.random.default_rng(0)
    n = 300_000                                        # several thread slices
    rows = rng.integers(0, 30000, n)
    cols = rng.integers(0, 30000, n)
    vals = np.concatenate([rng.gamma(2.0, 30.0, n // 2), rng.normal(0, 3, n - n // 2)])
    path = str(tmp_path / "a.mtx")
    np.savetxt(path, np.vstack([rows, cols, vals]).T, fmt="%d\t%d\t%.2f")
    r, c, v = _same_as_pandas(path)
    assert len(r) == n
    # integer-valued third column (out_int files), and a file without a trailing newline
    path2 = str(tmp_path / "b.mtx")
    np.savetxt(path2, np.vstack([rows[:1000], cols[:1000], np.floor(vals[:1000])]).T, fmt="%d\t%d\t%d")
    with open(path2, "rb+") as fh:
        fh.seek(-1, os.SEEK_END)
        fh.truncate()
    _same_as_pandas(path2)


def test_reader_edge_cases_like_pandas(tmp_path):
    path = str(tmp_path / "edge.mtx")
    with open(path, "w", newline="") as fh:
        fh.write("0\t1\t2.50\n"
                 "\n"                                   # blank line: skipped
                 "3\t4\t\n"                             # missing value -> NaN -> dropna
                 "5\t6\n"                               # two fields -> NaN -> dropna
                 "7\t8\tnan\n"                          # explicit NaN -> dropna
                 "9\t10\t-0.00\n"
                 "11\t12\t1e-05\n"                      # exponent
                 "13\t14\t123456789.125\r\n"            # CRLF
                 "15\t16\t0.1\n"
                 "17\t18\t179769313486231570000000000000000000000.0\n"
                 "19\t20\t-1234567.89\n"
                 "21\t22\t.5\n"
                 "23\t24\t5.\n"
                 "25\t26\t+3.25\n"
                 "27\t28\t0.300000000000004\n"           # 15 significant digits: still exact in both
                 "29\t30\t1.7976931348623157e308\n"
                 "31\t32\t4.9e-324")                    # no trailing newline
    r, c, v = _same_as_pandas(path)
    assert list(r) == [0, 9, 11, 13, 15, 17, 19, 21, 23, 25, 27, 29, 31]
    assert np.signbit(v[1]) and v[1] == 0.0
    # a field that is not a number is an error, not a silent NaN
    bad = str(tmp_path / "bad.mtx")
    open(bad, "w").write("0\t1\t2.5\n0\tx\t1.0\n")
    with pytest.raises(Exception):
        mtx_io.read_mtx(bad)
    with pytest.raises(Exception):
        mtx_io.read_mtx(str(tmp_path / "missing.mtx"))
    # literals with more than 15 significant digits: pandas' default parser (xstrtod) is NOT correctly
    # rounded there (0.30000000000000004 -> 0.3); ours is (== Python float()); never more than 1 ulp apart.
    # The pipeline's own files are "%.2f": far inside the range where both are exact.
    lits = ["0.30000000000000004", "123456789.123456789", "3.141592653589793238", "2.2250738585072014e-308"]
    longp = str(tmp_path / "long.mtx")
    open(longp, "w").write("".join(f"{i}\t{i}\t{l}\n" for i, l in enumerate(lits)))
    _, _, v = mtx_io.read_mtx(longp)
    pv = _pandas_read(longp)[2]
    assert [float(x) for x in v] == [float(l) for l in lits]
    assert np.all(np.abs(v - pv) <= np.spacing(np.abs(v)))
    empty = str(tmp_path / "empty.mtx")
    open(empty, "w").close()
    r, c, v = mtx_io.read_mtx(empty)
    assert len(r) == 0


def test_writer_is_byte_identical_to_savetxt(tmp_path):
    rng = np.random.default_rng(1)
    n = 400_000
    rows = rng.integers(0, 123541, n)
    cols = rng.integers(0, 123541, n)
    vals = np.concatenate([
        rng.gamma(2.0, 30.0, n // 4), rng.normal(0, 3, n // 4),
        rng.integers(-100000, 100000, n // 4) / 8.0,            # exact binary ties at the 3rd decimal: x.xx5
        rng.integers(-10**7, 10**7, n - 3 * (n // 4)) / 1000.0])  # decimal .xx5 inputs (not ties in binary)
    special = np.array([0.0, -0.0, -0.001, 0.005, 0.015, 0.025, 0.125, 0.375, 2.675, 1e15, -1e15, 1e22,
                        123456789012.345, 0.994999999999, 0.995, 0.9950000000000001, 5e-324, 99999999999.995,
                        1e11, 1e11 - 0.005, np.inf, -np.inf, np.nan])
    vals = np.concatenate([vals, special])
    rows = np.concatenate([rows, np.arange(len(special))])
    cols = np.concatenate([cols, np.arange(len(special))])
    ours, ref = str(tmp_path / "ours.mtx"), str(tmp_path / "ref.mtx")
    mtx_io.write_mtx(ours, rows, cols, vals)
    np.savetxt(ref, np.vstack([rows, cols, vals]).T, fmt="%d\t%d\t%.2f")
    assert open(ours, "rb").read() == open(ref, "rb").read()
    # out_int: '%d' % float truncates toward zero
    fin = np.isfinite(vals)
    mtx_io.write_mtx(ours, rows[fin], cols[fin], vals[fin], out_int=True)
    np.savetxt(ref, np.vstack([rows[fin], cols[fin], vals[fin]]).T, fmt="%d\t%d\t%d")
    assert open(ours, "rb").read() == open(ref, "rb").read()


def test_write_then_read_round_trip_and_pinned32(tmp_path):
    rng = np.random.default_rng(2)
    n = 100_000
    rows, cols = rng.integers(0, 5000, n), rng.integers(0, 5000, n)
    vals = np.round(rng.gamma(2.0, 30.0, n), 2)
    path = str(tmp_path / "rt.mtx")
    mtx_io.write_mtx(path, rows, cols, vals)
    r, c, v, (r32, c32, v32) = mtx_io.read_mtx(path, pinned32=True)
    assert np.array_equal(r, rows) and np.array_equal(c, cols) and np.array_equal(v, vals)
    import torch
    assert r32.is_pinned() == torch.cuda.is_available()
    assert np.array_equal(r32.numpy(), rows.astype(np.int32)) and np.array_equal(c32.numpy(), cols.astype(np.int32))
    assert np.array_equal(v32.numpy(), vals.astype(np.float32))


def test_dropped_lines_across_slices_direct_and_staged_paths(tmp_path):
    """Blank / incomplete / NaN lines sprinkled over a file of several thread slices: the direct
    parse (capacity = lines of the file) closes the gaps the dropped lines leave in every slice; a
    capacity of exactly the entry count (hia_mtx_scan) takes the staged path. Both equal pandas."""
    import ctypes
    from hiarch_b200 import _lib
    rng = np.random.default_rng(3)
    n = 200_000
    rows, cols = rng.integers(0, 5000, n), rng.integers(0, 5000, n)
    vals = np.round(rng.normal(0, 50, n), 2)
    kind = rng.choice(5, n, p=[.9, .03, .03, .02, .02])
    path = str(tmp_path / "gaps.mtx")
    with open(path, "w") as fh:
        for r, c, v, k in zip(rows, cols, vals, kind):
            if k == 0:
                fh.write("%d\t%d\t%.2f\n" % (r, c, v))
            elif k == 1:
                fh.write("\n")
            elif k == 2:
                fh.write("%d\t%d\n" % (r, c))
            elif k == 3:
                fh.write("%d\t%d\tnan\n" % (r, c))
            else:
                fh.write("%d\t%d\t%.3e\r\n" % (r, c, v))              # slow path: exponent + CRLF
    pr, pc, pv = _pandas_read(path)
    for threads in (1, 3, 8):
        r, c, v = mtx_io.read_mtx(path, threads=threads)                  # direct path
        assert np.array_equal(r, pr) and np.array_equal(c, pc) and np.array_equal(v, pv)
    lib = _lib.load()
    mr, mc, neg = ctypes.c_int64(0), ctypes.c_int64(0), ctypes.c_int32(0)
    cnt = lib.hia_mtx_scan(path.encode(), 4, ctypes.addressof(mr), ctypes.addressof(mc), ctypes.addressof(neg))
    assert cnt == len(pr) and mr.value == pr.max() and mc.value == pc.max() and neg.value == 1
    assert cnt < lib.hia_mtx_count_lines(path.encode(), 4)
    r = np.empty(cnt, np.int64); c = np.empty(cnt, np.int64); v = np.empty(cnt, np.float64)
    r32 = np.empty(cnt, np.int32); v32 = np.empty(cnt, np.float32)
    got = ctypes.c_int64(0)
    st = lib.hia_mtx_parse(path.encode(), 4, cnt, r.ctypes.data, c.ctypes.data, v.ctypes.data,
                           r32.ctypes.data, None, v32.ctypes.data, ctypes.addressof(got))     # staged path
    assert st == 0 and got.value == cnt
    assert np.array_equal(r, pr) and np.array_equal(c, pc) and np.array_equal(v, pv)
    assert np.array_equal(r32, pr.astype(np.int32)) and np.array_equal(v32, pv.astype(np.float32))
    st = lib.hia_mtx_parse(path.encode(), 4, cnt - 1, r.ctypes.data, c.ctypes.data, v.ctypes.data,
                           None, None, None, ctypes.addressof(got))
    assert st == -5                                                        # HIA_ERR_WORKSPACE


def test_read_sps_file_csr_fast_path_equals_scipy(tmp_path):
    """read_sps_file builds the CSR directly when the file is in canonical order (what to_output
    writes) and falls back to scipy's coo -> csr otherwise: identical arrays either way."""
    import scipy.sparse as sps
    from hiarch_b200.new_hic_class import read_sps_file, _csr_from_sorted_coo
    rng = np.random.default_rng(9)
    n = 400
    dense = np.triu(np.round(rng.gamma(1.0, 2.0, (n, n)) * (rng.random((n, n)) < .3), 2))
    dense[5, :] = 0                                               # an empty row
    dense[7, 9] = 0
    r, c = np.nonzero(dense)
    v = dense[r, c]
    v[3] = 0.0                                                    # an explicit zero is a stored entry
    sorted_path, shuffled_path, dup_path = (str(tmp_path / f"{k}.mtx") for k in ("s", "u", "d"))
    np.savetxt(sorted_path, np.vstack([r, c, v]).T, fmt="%d\t%d\t%.2f")
    perm = rng.permutation(len(r))
    np.savetxt(shuffled_path, np.vstack([r[perm], c[perm], v[perm]]).T, fmt="%d\t%d\t%.2f")
    np.savetxt(dup_path, np.vstack([np.r_[r, r[:10]], np.r_[c, c[:10]], np.r_[v, v[:10]]]).T, fmt="%d\t%d\t%.2f")
    vv = np.array([float("%.2f" % x) for x in v])
    ref = sps.coo_matrix((vv, (r, c)), shape=(n, n)).tocsr()
    assert _csr_from_sorted_coo(r[perm], c[perm], vv[perm], (n, n)) is None
    assert _csr_from_sorted_coo(np.r_[r, r[-1:]], np.r_[c, c[-1:]], np.r_[vv, vv[-1:]], (n, n)) is None
    for path in (sorted_path, shuffled_path):
        m = read_sps_file(path).matrix                           # shape from the largest indices
        assert m.shape == (r.max() + 1, c.max() + 1)
    bed = str(tmp_path / "w.bed")
    with open(bed, "w") as fh:
        for i in range(n):
            fh.write(f"chr1\t{i * 1000}\t{(i + 1) * 1000}\t{i}\n")
    for path in (sorted_path, shuffled_path):
        m = read_sps_file(path, bed).matrix
        assert m.shape == ref.shape and m.dtype == ref.dtype
        assert np.array_equal(m.indptr, ref.indptr) and np.array_equal(m.indices, ref.indices)
        assert np.array_equal(m.data, ref.data) and m.indices.dtype == ref.indices.dtype
        assert m.nnz == len(r)                                    # the explicit zero is kept
    d = read_sps_file(dup_path, bed).matrix                       # duplicates are summed by scipy
    ref_d = sps.coo_matrix((np.r_[vv, vv[:10]], (np.r_[r, r[:10]], np.r_[c, c[:10]])), shape=(n, n)).tocsr()
    assert (d != ref_d).nnz == 0 and d.nnz == ref_d.nnz


def test_reader_vs_pandas_by_literal_form(tmp_path):
    """Differential check over many literal forms: plain decimals of <= 15 digits (every file the
    pipeline writes) are bit-identical to pandas; longer literals and exponent forms are parsed
    correctly rounded (strtod) where pandas' fast xstrtod is a few ulp off -- so they may differ from
    pandas, but never from Python's float()."""
    import re
    rng = np.random.default_rng(7)
    forms = ["%.2f", "%.10f", "%.17g", "%d", "+%.3f", "%.3e", "00%.1f", "%.15f"]
    n = 40000
    path = str(tmp_path / "forms.mtx")
    texts = []
    with open(path, "w") as fh:
        for _ in range(n):
            x = rng.normal(0, 100) * 10.0 ** int(rng.integers(-6, 6))
            f = forms[int(rng.integers(0, len(forms)))]
            t = f % (int(x) if f == "%d" else (abs(x) if f[0] in "+0" else x))
            texts.append(t)
            fh.write("%d\t%d\t%s\n" % (rng.integers(0, 10 ** 6), rng.integers(0, 10 ** 6), t))
    r, c, v = mtx_io.read_mtx(path, threads=3)
    pr, pc, pv = _pandas_read(path)
    assert np.array_equal(r, pr) and np.array_equal(c, pc)
    exact = np.array([float(t) for t in texts])
    assert np.array_equal(v, exact)                               # correctly rounded, every form
    short = np.array([("e" not in t) and len(re.sub(r"[^0-9]", "", t)) <= 15 for t in texts])
    assert short.sum() > n // 3
    assert np.array_equal(v[short], pv[short])                    # == pandas wherever pandas is exact
    assert np.all(np.abs(v - pv) <= 1e-11 * np.abs(v))
