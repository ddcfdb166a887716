"""CPU: the Python drop-in boundary keeps the reference's call signatures --
same parameter names in the same order with the same defaults (misspellings included). The
reference side is a committed fixture (tests/golden/ref_signatures.json, dumped from the real
reference by tests/golden/gen_signatures.py). The mirror may append extra keyword parameters after
the reference's (device, workspaces ...), never reorder or rename."""
import importlib
import inspect
import json
import os

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
REF = json.load(open(os.path.join(ROOT, "tests", "golden", "ref_signatures.json")))

# reference "module:name" -> (mirror module, mirror name)
MIRROR = {
    "new_hic_class": "hiarch_b200.new_hic_class",
    "iutils.read_matrix": "hiarch_b200.new_hic_class",
    "oe_map.S1_preprocess": "hiarch_b200.preprocess",
    "oe_map.S2_oe": "hiarch_b200.oe_map",
    "oe_map.main_get_oe": "hiarch_b200.oe_map",
    "oe_map.S3_denoise": "hiarch_b200.oe_map",
    "NormDis": "hiarch_b200.oe_map",
    "GF_S1_get_center": "hiarch_b200.confor",
    "GF_S2_get_score": "hiarch_b200.confor",
    "oe_map.S4_norm_value": "hiarch_b200.oe_map",
    "complex.src.S1_get_adj_mtx": "hiarch_b200.complex",
    "complex.src.main_local": "hiarch_b200.complex",
    "confor.src.S1_preprocess.S1_1_filtering": "hiarch_b200.confor",
    "confor.src.S1_preprocess.S1_2_find_anchor_main": "hiarch_b200.confor",
    "confor.src.S2_to_train_mtx.to_train_mtx": "hiarch_b200.confor",
    "confor.src.S5_sim_to_pat.to_confor": "hiarch_b200.confor",
    "run_hic_dataset": "hiarch_b200.run_hic_dataset",
    "CorrectMap": "hiarch_b200.preprocess",
}

# parameters whose DEFAULT legitimately differs (names and order are still checked):
# reference key -> {parameter: reason}
DEFAULT_DIFFERS = {
    "confor.src.S5_sim_to_pat.to_confor:main_to_confor": {
        "intra_ave_file": "install path of the reference's .mat; None = the templates shipped in data/ave_maps.npz",
        "inter_ave_file": "same",
    },
}


def _resolve(key):
    mod_name, name = key.split(":")
    mod = importlib.import_module(MIRROR[mod_name])
    obj = mod
    for part in name.split("."):
        obj = getattr(obj, part)
    return obj


@pytest.mark.parametrize("key", sorted(REF))
def test_mirror_keeps_reference_signature(key):
    try:
        obj = _resolve(key)
    except AttributeError:
        pytest.fail(f"{key} has no mirror in {MIRROR[key.split(':')[0]]}")
    ours = [(p.name, None if p.default is inspect.Parameter.empty else repr(p.default))
            for p in inspect.signature(obj).parameters.values()
            if p.kind not in (inspect.Parameter.VAR_KEYWORD, inspect.Parameter.VAR_POSITIONAL)]
    ref = [(n, d) for n, d, kind in REF[key] if kind not in ("VAR_KEYWORD", "VAR_POSITIONAL")]
    skip = DEFAULT_DIFFERS.get(key, {})
    ours_cmp = [(n, None if n in skip else d) for n, d in ours[:len(ref)]]
    ref = [(n, None if n in skip else d) for n, d in ref]
    assert ours_cmp == ref, f"{key}\n  reference: {ref}\n  mirror:    {ours}"
    for name, default in ours[len(ref):]:
        assert default is not None, f"{key}: extra mirror parameter {name!r} must have a default"
