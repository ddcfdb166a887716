"""CPU: the per-(row, centre) arithmetic of the prefix-sum intra-x kernels
(hiarch_b200/csrc/hia_intra_x_core.cuh, `__host__ __device__`) compiled for the HOST with g++
(tests/host/intra_x_host.cpp) and checked against the real reference's get_intra_x output
(golden) and the oracle -- the device kernels run exactly these functions."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from oracle import hia_oracle as O

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


@pytest.fixture(scope="module")
def host_lib(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("ix") / "intra_x_host.so")
    subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-o", so,
                           os.path.join(ROOT, "tests", "host", "intra_x_host.cpp")])
    lib = ctypes.CDLL(so)
    lib.intra_x_prefix_host.restype = ctypes.c_int
    lib.intra_x_prefix_host.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_int,
                                        ctypes.c_void_p]
    return lib


def _run(lib, a32, k_len=None):
    n = a32.shape[0]
    a32 = np.ascontiguousarray(a32, dtype=np.float32)
    k_len = max(3, int(n * .05)) if k_len is None else k_len
    out = np.zeros(n)
    assert lib.intra_x_prefix_host(a32.ctypes.data, n, n, k_len, out.ctypes.data) == 0
    return out


@pytest.mark.parametrize("n", [1, 2, 3, 7, 20, 61, 150, 257])
def test_prefix_core_matches_oracle_on_unsymmetric_blocks(host_lib, n):
    rng = np.random.default_rng(n)
    a = rng.standard_normal((n, n)).astype(np.float32)           # get_intra_x does not assume symmetry
    ref = O.intra_x(a.astype(np.float64))
    got = _run(host_lib, a)
    assert np.all(np.abs(got - ref) <= 1e-12 + 1e-12 * np.abs(ref)), np.abs(got - ref).max()


def test_prefix_core_matches_the_real_reference(host_lib, golden):
    g = golden("g1chr")
    filt = g["filt"].astype(np.float32)
    got = _run(host_lib, filt)
    ref = g["intra_x"]                                           # get_intra_x of the reference, fp64 input
    assert np.all(np.abs(got - ref) <= 1e-6 * np.abs(ref) + 1e-7)   # fp32 storage of the map
    ref32 = O.intra_x(filt.astype(np.float64))
    assert np.all(np.abs(got - ref32) <= 1e-12 + 1e-12 * np.abs(ref32))


def test_prefix_core_large_offsets_and_counts(host_lib):
    """Count-like values (large positive sums: the cancellation case of the prefix differences)."""
    rng = np.random.default_rng(5)
    n = 400
    d = np.abs(np.arange(n)[:, None] - np.arange(n)[None, :])
    a = rng.poisson(2000.0 * (d + 1.0) ** -1.08 + 3.0).astype(np.float32)
    ref = O.intra_x(a.astype(np.float64))
    got = _run(host_lib, a)
    assert np.all(np.abs(got - ref) <= 1e-11 * np.abs(ref)), (np.abs(got - ref) / np.abs(ref)).max()
