"""Interpolation onto the boundary dofs (mirror of /root/reference/sucemfem/Interpolation/__init__.py)."""
# Interface (module, class, method and argument names) after the file of cemagg/sucem-fem cited above:
# Copyright (C) 2011 Stellenbosch University, authors Neilen Marais; GNU GPL v3 or later
# (see NOTICE.md at the repository root).  The implementation below is this repository's own.
from .Interpolation import *  # noqa: F401,F403
