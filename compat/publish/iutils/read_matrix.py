"""publish/iutils/read_matrix.py -> hiarch_b200 (read_matrix :21-30; the anchors table :195-342)."""
from hiarch_b200.new_hic_class import read_matrix  # noqa: F401
from hiarch_b200.anchors import AnchorTable as Anchors  # noqa: F401
