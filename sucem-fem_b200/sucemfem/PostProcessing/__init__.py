"""Post-processing functionals (mirror of /root/reference/sucemfem/PostProcessing/__init__.py:22)."""
from .PostProcessing import *  # noqa: F401,F403
