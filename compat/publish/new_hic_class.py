"""publish/new_hic_class.py -> hiarch_b200.new_hic_class (CUDA-backed mirror)."""
from hiarch_b200.new_hic_class import *  # noqa: F401,F403
from hiarch_b200.new_hic_class import (ArrayHiCMtx, GenomeIndex, HiCMtx, SpsHiCMtx, change_suffix,  # noqa: F401
                                       concat_mtxes, read_matrix, read_sps_file)
from hiarch_b200.sgd import ScaledGenomicDistance  # noqa: F401
