"""publish/confor/src/S1_preprocess/S1_2_find_anchor_main.py -> hiarch_b200.confor."""
from hiarch_b200.confor import main_find_anchor  # noqa: F401
