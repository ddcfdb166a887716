"""publish/confor/src/S5_sim_to_pat/to_confor.py -> hiarch_b200.confor."""
from hiarch_b200.confor import get_n_norm_ave_maps, change_map_name, main_to_confor, to_confor  # noqa: F401
