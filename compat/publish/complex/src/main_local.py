"""publish/complex/src/main_local.py -> hiarch_b200.complex."""
from hiarch_b200.complex import get_local, local_io, main_local  # noqa: F401
