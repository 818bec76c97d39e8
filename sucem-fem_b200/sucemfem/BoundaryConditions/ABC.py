"""First-order absorbing boundary condition (mirror of /root/reference/sucemfem/BoundaryConditions/ABC.py:26-58):
bilinear form <n x v, n x u> over the whole exterior boundary (region numbers are ignored, ABC.py:32-33)."""
# Interface (module, class, method and argument names) after the file of cemagg/sucem-fem cited above:
# Copyright (C) 2011 Stellenbosch University, authors Neilen Marais; GNU GPL v3 or later
# (see NOTICE.md at the repository root).  The implementation below is this repository's own.
from .. import Forms
from .base_boundary_condition import BoundaryCondition


class ABCBoundaryCondition(BoundaryCondition):
    def get_bilinear_form(self, test_function=None, trial_function=None):
        return Forms.BilinearForm('abc', self.function_space)
