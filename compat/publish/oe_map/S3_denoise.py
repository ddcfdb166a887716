"""publish/oe_map/S3_denoise.py -> hiarch_b200.oe_map."""
from hiarch_b200.oe_map import denoise_one_method, main_denoise, tid_off_outliers, used_denoise  # noqa: F401
