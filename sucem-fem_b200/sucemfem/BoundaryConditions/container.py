"""Collection of boundary conditions (mirror of /root/reference/sucemfem/BoundaryConditions/container.py:24-69)."""
# Interface (module, class, method and argument names) after the file of cemagg/sucem-fem cited above:
# Copyright (C) 2011 Stellenbosch University, authors Neilen Marais; GNU GPL v3 or later
# (see NOTICE.md at the repository root).  The implementation below is this repository's own.
from .. import Forms


class BoundaryConditions(object):
    def __init__(self):
        self.boundary_conditions = {}

    def add_boundary_condition(self, boundary_condition):
        bc_num = boundary_condition.region_number
        # boundary condition numbers have to be unique
        assert(bc_num not in self.boundary_conditions)
        self.boundary_conditions[bc_num] = boundary_condition

    def apply_essential(self, A, b=None):
        """Apply essential boundary conditions to system matrix A and optional RHS b"""
        for bc_num, bc in self.boundary_conditions.items():
            apply_fn = bc.get_essential_application_func()
            if b:
                apply_fn(A, b)
            else:
                apply_fn(A)

    def get_essential_row_flags(self, function_space):
        """Union of the constrained-dof flags of all essential conditions (device uint8 [ndof]) or None."""
        flags = None
        for bc_num, bc in self.boundary_conditions.items():
            flags = bc.get_essential_row_flags(function_space, flags)
        return flags

    def get_linear_form(self):
        lin_form = Forms.NullForm()
        for bc_num, bc in self.boundary_conditions.items():
            lin_form = lin_form + bc.get_linear_form()
        return lin_form

    def get_bilinear_form(self):
        bilin_form = Forms.NullForm()
        for bc_num, bc in self.boundary_conditions.items():
            bilin_form = bilin_form + bc.get_bilinear_form()
        return bilin_form

    def set_function_space(self, function_space):
        self.function_space = function_space
        for bc in self.boundary_conditions.values():
            bc.set_function_space(function_space)
