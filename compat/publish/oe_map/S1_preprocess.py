"""publish/oe_map/S1_preprocess.py -> hiarch_b200.preprocess."""
from hiarch_b200.preprocess import preprocess_mtx  # noqa: F401
