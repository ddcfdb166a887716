"""CPU: the `publish/`-shaped import root (compat/publish). With it first on the module search path
the reference's own, UNCHANGED step scripts (publish/HiArch/*.py) resolve every import to the
CUDA-backed mirrors in hiarch_b200 -- so `one_click_pipeline.sh`, which only launches those
scripts, runs on the B200 path. Needs the reference checkout (absent on the GPU box: skipped
there); no kernel is launched, the scripts are loaded without running their `__main__` blocks."""
import os
import runpy
import subprocess
import sys

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
COMPAT = os.path.join(ROOT, "compat", "publish")
REF = os.environ.get("HIARCH_REFERENCE", "/root/reference")
SCRIPTS = os.path.join(REF, "publish", "HiArch")

# script -> {global name the script imports: module it must come from}
EXPECT = {
    "NormDis.py": {"read_sps_file": "hiarch_b200.new_hic_class", "oe_map_io": "hiarch_b200.oe_map"},
    "CorrectMap.py": {"read_sps_file": "hiarch_b200.new_hic_class", "heatmap_plot": "iutils.plot_heatmap"},
    "Checkerboard.py": {"local_io": "hiarch_b200.complex"},
    "GF_S1_get_center.py": {"filter_io": "hiarch_b200.confor", "main_find_anchor": "hiarch_b200.confor"},
    "GF_S2_get_score.py": {"to_train_mtx": "hiarch_b200.confor", "main_to_confor": "hiarch_b200.confor"},
}

CHECK = r'''
import runpy, sys, json
script, compat, root = sys.argv[1:4]
sys.path[:0] = [compat, root]
g = runpy.run_path(script, run_name="loaded_not_run")
out = {k: getattr(v, "__module__", None) for k, v in g.items() if callable(v)}
import iutils.plot_heatmap, new_hic_class, oe_map.main_get_oe, complex.src.main_local
out["__files__"] = [m.__file__ for m in (iutils.plot_heatmap, new_hic_class, oe_map.main_get_oe, complex.src.main_local)]
print("REC " + json.dumps(out))
'''


@pytest.mark.skipif(not os.path.isdir(SCRIPTS), reason="reference checkout not present")
@pytest.mark.parametrize("script", sorted(EXPECT))
def test_reference_step_script_resolves_to_the_cuda_mirrors(script, tmp_path):
    import json
    chk = tmp_path / "chk.py"
    chk.write_text(CHECK)
    out = subprocess.run([sys.executable, str(chk), os.path.join(SCRIPTS, script), COMPAT, ROOT],
                         capture_output=True, text=True, timeout=300, cwd=str(tmp_path))
    assert out.returncode == 0, out.stderr[-2000:]
    rec = json.loads([l for l in out.stdout.splitlines() if l.startswith("REC ")][0][4:])
    for name, module in EXPECT[script].items():
        assert rec.get(name) == module, (name, rec.get(name))
    for f in rec["__files__"]:                                  # nothing came from the reference's own publish/
        assert f.startswith(COMPAT), f


def test_compat_root_covers_the_reference_layout():
    """Every module path the five step scripts import exists in compat/publish."""
    for rel in ("new_hic_class.py", "run_hic_dataset.py", "iutils/read_matrix.py", "iutils/plot_heatmap.py",
                "oe_map/main_get_oe.py", "oe_map/S1_preprocess.py", "oe_map/S2_oe.py", "oe_map/S3_denoise.py",
                "oe_map/S4_norm_value.py", "complex/src/main_local.py", "complex/src/S1_get_adj_mtx.py",
                "confor/src/S1_preprocess/S1_1_filtering.py", "confor/src/S1_preprocess/S1_2_find_anchor_main.py",
                "confor/src/S2_to_train_mtx/to_train_mtx.py", "confor/src/S5_sim_to_pat/to_confor.py"):
        assert os.path.isfile(os.path.join(COMPAT, rel)), rel
