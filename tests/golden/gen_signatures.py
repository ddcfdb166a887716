"""Dump the call signatures of the reference's drop-in boundary into
tests/golden/ref_signatures.json: parameter names in order + repr of the defaults, read with
`inspect.signature` from the REAL reference imported through oracle/ref_shim.py.

    python tests/golden/gen_signatures.py          (needs /root/reference; run in the authoring container)
"""
import importlib
import inspect
import json
import os
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)
from oracle import ref_shim

# reference module -> names (functions, or Class.method)
BOUNDARY = {
    "new_hic_class": ["read_sps_file", "concat_mtxes", "GenomeIndex.__init__", "GenomeIndex.condense",
                      "GenomeIndex.keep_chros", "GenomeIndex.to_output",
                      "HiCMtx.obs_d_exp", "HiCMtx.yield_by_chro", "ArrayHiCMtx.__init__", "ArrayHiCMtx.log",
                      "ArrayHiCMtx.to_dense", "ArrayHiCMtx.to_sps", "ArrayHiCMtx.to_all", "ArrayHiCMtx.to_triu",
                      "ArrayHiCMtx.to_output", "SpsHiCMtx.__init__", "SpsHiCMtx.to_dense", "SpsHiCMtx.to_sps",
                      "SpsHiCMtx.to_all", "SpsHiCMtx.to_triu", "SpsHiCMtx.to_output"],
    "iutils.read_matrix": ["read_matrix"],
    "oe_map.S1_preprocess": ["preprocess_mtx"],
    "oe_map.S2_oe": ["oe_map_core"],
    "oe_map.main_get_oe": ["oe_map_main", "oe_map_io"],
    "NormDis": ["norm_dis"],
    "GF_S1_get_center": ["anchors_io"],
    "GF_S2_get_score": ["get_confor"],
    "oe_map.S3_denoise": ["denoise_one_method", "tid_off_outliers", "main_denoise"],
    "oe_map.S4_norm_value": ["norm_to_zero_mean", "remove_outliers", "norm_de_mtx"],
    "complex.src.S1_get_adj_mtx": ["get_adj_mtx"],
    "complex.src.main_local": ["get_local", "main_local", "local_io"],
    "confor.src.S1_preprocess.S1_1_filtering": ["used_filter", "main_filter", "filter_io"],
    "confor.src.S1_preprocess.S1_2_find_anchor_main": ["main_find_anchor"],
    "confor.src.S2_to_train_mtx.to_train_mtx": ["to_train_mtx"],
    "confor.src.S5_sim_to_pat.to_confor": ["to_confor", "main_to_confor"],
    "run_hic_dataset": ["ParallelFunc.__init__", "ParallelFunc.vali_parallel"],
    "CorrectMap": ["correct_map"],
}


def describe(obj):
    sig = inspect.signature(obj)
    return [[p.name, None if p.default is inspect.Parameter.empty else repr(p.default), p.kind.name]
            for p in sig.parameters.values()]


def main():
    ref_shim.install()
    out = {}
    for mod_name, names in BOUNDARY.items():
        try:
            mod = importlib.import_module(mod_name)
        except Exception as exc:                        # noqa: BLE001
            print(f"skip {mod_name}: {type(exc).__name__}: {exc}")
            continue
        for name in names:
            obj = mod
            try:
                for part in name.split("."):
                    obj = getattr(obj, part)
            except AttributeError:
                print(f"missing in reference: {mod_name}.{name}")
                continue
            out[f"{mod_name}:{name}"] = describe(obj)
    path = os.path.join(ROOT, "tests", "golden", "ref_signatures.json")
    json.dump(out, open(path, "w"), indent=1, sort_keys=True)
    print(f"{len(out)} signatures -> {path}")


if __name__ == "__main__":
    main()
