"""Interpolation onto the boundary dofs (mirror of /root/reference/sucemfem/Interpolation/__init__.py)."""
from .Interpolation import *  # noqa: F401,F403
