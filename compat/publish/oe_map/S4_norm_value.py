"""publish/oe_map/S4_norm_value.py -> hiarch_b200.oe_map."""
from hiarch_b200.oe_map import norm_de_mtx, norm_to_zero_mean, remove_outliers  # noqa: F401
