"""publish/confor/src/S1_preprocess/S1_1_filtering.py -> hiarch_b200.confor."""
from hiarch_b200.confor import filter_io, main_filter, used_filter  # noqa: F401
