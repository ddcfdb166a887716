"""publish/oe_map/main_get_oe.py -> hiarch_b200.oe_map."""
from hiarch_b200.oe_map import oe_map_io, oe_map_main  # noqa: F401
