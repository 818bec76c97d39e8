"""Boundary condition base class with nil contributions
(mirror of /root/reference/sucemfem/BoundaryConditions/base_boundary_condition.py:24-112)."""
# Interface (module, class, method and argument names) after the file of cemagg/sucem-fem cited above:
# Copyright (C) 2011 Stellenbosch University, authors Neilen Marais, Evan Lezar; GNU GPL v3 or later
# (see NOTICE.md at the repository root).  The implementation below is this repository's own.
from .. import Forms


class BoundaryCondition(object):
    mesh_function = None
    function_space = None
    boundary_value_expression = None

    def set_boundary_value_expression(self, boundary_value_expression):
        self.boundary_value_expression = boundary_value_expression

    def set_mesh_function(self, mesh_function):
        self.mesh_function = mesh_function

    def set_region_number(self, region_number):
        self.region_number = region_number

    def set_function_space(self, function_space):
        self.function_space = function_space

    def get_essential_application_func(self, function_space=None):
        return lambda x: None

    def get_essential_row_flags(self, function_space, flags=None):
        """Device uint8 flags of the dofs this condition constrains (None: no essential part)."""
        return flags

    def get_linear_form(self, test_function=None):
        return Forms.NullForm()

    def get_bilinear_form(self, test_function=None, trial_function=None):
        return Forms.NullForm()
