"""One process, one CUDA context: the GPU parity tests of the kernels touched since the last full
GPU run, then tools/probe_c4.py. For a short gpurun call (the torch import / context creation is
paid once). Logs go to gpurun_out/."""
import contextlib
import io
import os
import runpy
import sys
import time

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
os.chdir(ROOT)
sys.path.insert(0, ROOT)
os.makedirs("gpurun_out", exist_ok=True)
t0 = time.time()
import pytest

K = ("not similarity and not topk_eig and not scatter and not test_oe_ and not hot_path "
     "and not preprocess and not remap and not row_sharded")
with open("gpurun_out/v_tests.log", "w") as f, contextlib.redirect_stdout(f), contextlib.redirect_stderr(f):
    rc = pytest.main(["tests", "-m", "gpu", "-q", "-p", "no:cacheprovider", "-k", K] + sys.argv[1:])
    print("pytest rc", int(rc), "at", round(time.time() - t0, 1), "s")
print("pytest rc", int(rc), "at", round(time.time() - t0, 1), "s", flush=True)
sys.argv = ["probe_c4.py"]
with open("gpurun_out/v_probe.log", "w") as f, contextlib.redirect_stdout(f), contextlib.redirect_stderr(f):
    try:
        runpy.run_path(os.path.join(ROOT, "tools", "probe_c4.py"), run_name="__main__")
    except BaseException as exc:                      # noqa: BLE001 -- keep the log
        import traceback
        traceback.print_exc()
print("probe done at", round(time.time() - t0, 1), "s", flush=True)
