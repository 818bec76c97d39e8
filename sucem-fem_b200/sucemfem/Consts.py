"""Physical constants (mirror of /root/reference/sucemfem/Consts.py:27-31; the values are definitions)."""
# Interface (module, class, method and argument names) after the file of cemagg/sucem-fem cited above:
# Copyright (C) 2011 Stellenbosch University, authors Neilen Marais; GNU GPL v3 or later
# (see NOTICE.md at the repository root).  The implementation below is this repository's own.
import numpy as N

pi = N.pi
e = N.e
mu0 = N.pi * 4 * 1e-7
c0 = 299792458
eps0 = 10 ** 7 / 4 / N.pi / c0 / c0
Z0 = mu0 * c0
