"""publish/run_hic_dataset.py -> hiarch_b200.run_hic_dataset."""
from hiarch_b200.run_hic_dataset import *  # noqa: F401,F403
from hiarch_b200.run_hic_dataset import CaptureParas, ParallelFunc, read_mtx, yield_mp_idxes  # noqa: F401
