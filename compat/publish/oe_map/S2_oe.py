"""publish/oe_map/S2_oe.py -> hiarch_b200.oe_map."""
from hiarch_b200.oe_map import counts_must_decrese, k, mode, oe_map_core  # noqa: F401
