"""CPU: the fp64 box-filter line function of hia_used_filter_f64
(hiarch_b200/csrc/hia_box64_core.cuh, `__host__ __device__`) compiled for the HOST and compared BIT
FOR BIT with scipy.ndimage.uniform_filter (mode 'reflect') -- GF_S1's used_filter,
publish/confor/src/S1_preprocess/S1_1_filtering.py:13-16 via oe_map/S3_denoise.py:54-57. The
device kernel runs exactly this function (thread per line)."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
from scipy.ndimage import uniform_filter

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


@pytest.fixture(scope="module")
def host_lib(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("box") / "box64_host.so")
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", so,
                           os.path.join(ROOT, "tests", "host", "box64_host.cpp")])
    lib = ctypes.CDLL(so)
    lib.box64_host.restype = ctypes.c_int
    lib.box64_host.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
    return lib


@pytest.mark.parametrize("shape,size", [((120, 100), 12), ((37, 53), 5), ((12, 40), 30), ((64, 64), 2),
                                        ((200, 9), 7), ((9, 200), 8), ((33, 33), 33), ((50, 70), 49),
                                        ((1, 25), 3), ((25, 1), 4), ((300, 300), 30), ((17, 90), 64)])
def test_box64_line_is_bit_exact_vs_scipy(host_lib, shape, size):
    rng = np.random.default_rng(shape[0] * 1000 + size)
    a = np.round(rng.standard_normal(shape) * 3, 2)              # %.2f-quantised values (ties)
    out = np.empty(shape)
    a = np.ascontiguousarray(a)
    assert host_lib.box64_host(a.ctypes.data, shape[0], shape[1], size, out.ctypes.data) == 0
    ref = uniform_filter(a, size=size, mode="reflect")
    assert np.array_equal(out, ref)
