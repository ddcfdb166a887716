"""TEST INFRASTRUCTURE ONLY -- import shim for the *real* HiArch reference.

Makes `/root/reference/publish` importable in the authoring container so that
`tests/golden/gen_golden.py` can generate golden vectors from the reference's own
code and `tests/test_oracle_vs_reference.py` can pin `oracle/hia_oracle.py` against it.
`/root/reference` does not exist on the GPU box: nothing in the product path, in
`-m gpu` tests, in `smoke()` or in `bench.py` imports this module.

What the shim does:
  1. puts `<ref>/publish` and `<ref>/publish/HiArch` on sys.path;
  2. registers stub modules for the plot-only / absent packages the reference imports
     at module top (`publish/new_hic_class.py:8,10`: seaborn, matplotlib;
     `publish/oe_map/S3_denoise.py:5`: skimage.restoration);
  3. no-ops the unconditional plot calls.
The wavelet step (`skimage.restoration.denoise_wavelet`, S3_denoise.py:45-47) has no
implementation here (scikit-image/PyWavelets absent, unpinned) -> the stub is the identity by
default and every fixture that crosses that call says so ("wavelet=identity");
`set_wavelet("db3")` swaps in oracle/wavelet_oracle.py's written-down restatement of the two
libraries' published algorithm (fixtures then say "wavelet=db3 (oracle/wavelet_oracle.py)": the
reference's own code around a stand-in for the one call that cannot run here).
"""
import os
import sys
import types

REF_ROOT = os.environ.get("HIARCH_REFERENCE", "/root/reference")
PUBLISH = os.path.join(REF_ROOT, "publish")


def available():
    return os.path.isfile(os.path.join(PUBLISH, "new_hic_class.py"))


class _Anything:
    """Attribute sink: any attribute access / call returns another sink."""

    def __getattr__(self, name):
        return _Anything()

    def __call__(self, *a, **k):
        return _Anything()

    def __iter__(self):
        return iter(())


def _stub(name, **attrs):
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    mod.__getattr__ = lambda attr: _Anything()  # PEP 562
    sys.modules[name] = mod
    return mod


_installed = False


def _identity_wavelet(arr, **kw):
    return arr


def _oracle_wavelet(arr, **kw):
    """Stand-in for skimage.restoration.denoise_wavelet, only for the reference's own call."""
    from oracle import wavelet_oracle
    if kw != {"wavelet": "db3", "method": "VisuShrink"}:
        raise NotImplementedError(f"denoise_wavelet stand-in: only the reference's call is restated, got {kw}")
    return wavelet_oracle.denoise_wavelet_db3_visushrink(arr)


def set_wavelet(kind):
    """'identity' (default) or 'db3': what the stubbed skimage.restoration.denoise_wavelet does."""
    fn = {"identity": _identity_wavelet, "db3": _oracle_wavelet}[kind]
    sys.modules["skimage.restoration"].denoise_wavelet = fn


def install():
    """Idempotently make the reference importable. Returns the `new_hic_class` module."""
    global _installed
    if not available():
        raise RuntimeError(f"reference not present at {REF_ROOT}")
    if not _installed:
        try:            # torch must be imported BEFORE the stub modules exist (its own import
            import torch  # noqa: F401  probes optional packages and trips over the stubs)
        except Exception:  # pragma: no cover
            pass
        for p in (os.path.join(PUBLISH, "HiArch"), PUBLISH):
            if p not in sys.path:
                sys.path.insert(0, p)
        for name in ("seaborn", "matplotlib", "matplotlib.pyplot", "matplotlib.patches",
                     "matplotlib.colors", "matplotlib.gridspec"):
            if name not in sys.modules:
                _stub(name)
        if "skimage" not in sys.modules:
            sk = _stub("skimage")
            rest = _stub("skimage.restoration",
                         denoise_wavelet=_identity_wavelet,        # see set_wavelet
                         denoise_bilateral=lambda arr, **kw: arr)
            sk.restoration = rest
        _installed = True
    import new_hic_class  # noqa: E402  (the reference's own module)

    _noop = lambda *a, **k: None
    # unconditional plot calls
    import iutils.plot_heatmap as ph
    ph.heatmap_plot = _noop
    import oe_map.main_get_oe as mg
    mg.heatmap_plot = _noop
    mg.histplot_oe_values = _noop
    try:
        import confor.src.S1_preprocess.S1_2_find_anchor_main as fam
        fam.plot_anchors = _noop
    except Exception:  # pragma: no cover
        pass
    try:
        import confor.src.S2_to_train_mtx.to_train_mtx as ttm
        ttm.plot_pre_mtx = _noop
    except Exception:  # pragma: no cover
        pass
    return new_hic_class
