"""publish/complex/src/S1_get_adj_mtx.py -> hiarch_b200.complex."""
from hiarch_b200.complex import get_adj_mtx  # noqa: F401
