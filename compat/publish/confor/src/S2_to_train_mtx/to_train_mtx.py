"""publish/confor/src/S2_to_train_mtx/to_train_mtx.py -> hiarch_b200.confor."""
from hiarch_b200.confor import to_train_mtx  # noqa: F401
