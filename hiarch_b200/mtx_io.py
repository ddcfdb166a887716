"""f2: the 3-column text wire format (`%d\\t%d\\t%.2f`) through the native reader / writer.

read_mtx   replaces pd.read_table(sps_file, names=['row', 'col', 'value']) + dropna
           (read_sps_file, publish/new_hic_class.py:2580-2581)
write_mtx  replaces np.savetxt(out_file, vstack([rows, cols, values]).T, fmt="%d\\t%d\\t%.2f")
           (SpsHiCMtx.to_output, publish/new_hic_class.py:2290-2297)

Both run in libhiarch_b200.so (csrc/hia_io.cu: mmap + one slice of lines per host thread).
"""
import ctypes

import numpy as np

from . import _lib


def read_mtx(path, threads=0, pinned32=False):
    """-> (rows int64, cols int64, vals float64) numpy arrays in file order (blank lines skipped,
    lines with a missing field dropped). pinned32=True additionally returns (rows32, cols32, vals32)
    as PINNED torch tensors filled in the same pass (int32 / int32 / fp32: what the scatter kernel
    takes), ready for an asynchronous H2D copy."""
    lib = _lib.load()
    cap = lib.hia_mtx_count_lines(path.encode(), threads)         # upper bound: lines of the file
    if cap < 0:
        raise _lib.HiaError(f"hia_mtx_count_lines({path!r}): cannot open the file")
    rows = np.empty(cap, dtype=np.int64)
    cols = np.empty(cap, dtype=np.int64)
    vals = np.empty(cap, dtype=np.float64)
    extra = None
    p32 = [None, None, None]
    if pinned32:
        import torch
        pin = torch.cuda.is_available()          # page-locked only where there is a device to copy to
        extra = tuple(torch.empty(cap, dtype=dt, pin_memory=pin) for dt in (torch.int32, torch.int32, torch.float32))
        p32 = [t.data_ptr() for t in extra]
    got = ctypes.c_int64(0)
    st = lib.hia_mtx_parse(path.encode(), threads, cap, rows.ctypes.data, cols.ctypes.data, vals.ctypes.data,
                           p32[0], p32[1], p32[2], ctypes.addressof(got))
    if st != 0:
        raise _lib.HiaError(f"hia_mtx_parse({path!r}): a field is neither a number nor a missing-value marker")
    n = got.value
    rows, cols, vals = rows[:n], cols[:n], vals[:n]
    if pinned32:
        return rows, cols, vals, tuple(t[:n] for t in extra)
    return rows, cols, vals


def write_mtx(path, rows, cols, vals, out_int=False, threads=0):
    """np.savetxt(path, [rows, cols, vals].T, fmt='%d\\t%d\\t%.2f' | '%d\\t%d\\t%d'), byte for byte."""
    lib = _lib.load()
    rows = np.ascontiguousarray(rows, dtype=np.int64)
    cols = np.ascontiguousarray(cols, dtype=np.int64)
    vals = np.ascontiguousarray(vals, dtype=np.float64)
    assert rows.shape == cols.shape == vals.shape and rows.ndim == 1
    st = lib.hia_mtx_write(path.encode(), rows.size, rows.ctypes.data, cols.ctypes.data, vals.ctypes.data,
                           1 if out_int else 0, threads)
    _lib.check(st, "hia_mtx_write")
