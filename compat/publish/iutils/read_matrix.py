"""publish/iutils/read_matrix.py -> hiarch_b200.new_hic_class (read_matrix :21-30)."""
from hiarch_b200.new_hic_class import read_matrix  # noqa: F401
