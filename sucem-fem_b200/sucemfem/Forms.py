"""Form descriptors with the reference's interface (/root/reference/sucemfem/Forms.py:25-116).

The reference builds UFL expressions that FFC compiles; here a bilinear form is a small descriptor
{kind, function_space, coefficient} that sucemfem.dolfin_compat.assemble / SystemMatrices hand to the CUDA
kernels.  NullForm is the additive identity, as in the reference.
"""
# Interface (module, class, method and argument names) after the file of cemagg/sucem-fem cited above:
# Copyright (C) 2011 Stellenbosch University, authors Neilen Marais; GNU GPL v3 or later
# (see NOTICE.md at the repository root).  The implementation below is this repository's own.


class NullForm(object):
    """Null form that can be added to a form while making a zero contribution (Forms.py:25-31)."""
    _integrals = []

    def __add__(self, other):
        return other


class BilinearForm(object):
    """kind in {'mass', 'stiffness', 'abc'}; coefficient: a Materials.MaterialFunction or None (= 1)."""

    def __init__(self, kind, function_space, coefficient=None):
        self.kind, self.function_space, self.coefficient = kind, function_space, coefficient

    def __add__(self, other):
        if isinstance(other, NullForm):
            return self
        raise NotImplementedError('sums of bilinear forms are assembled matrix by matrix')

    __radd__ = __add__


class GalerkinInteriorForms(object):
    """Basic Galerkin system forms for the interior part of the problem (Forms.py:34-63)."""

    def set_material_functions(self, material_functions):
        self.material_functions = material_functions

    def set_function_space(self, function_space):
        """Set function space used to solve the problem"""
        self.function_space = function_space

    def get_mass_form(self):
        raise NotImplementedError()

    def get_stiffness_form(self):
        raise NotImplementedError()


class EMGalerkinInteriorForms(GalerkinInteriorForms):
    """mass eps_r*inner(v,u)*dx and stiffness dot(curl(v),curl(u))/mu_r*dx (Forms.py:65-85)."""

    def get_mass_form(self):
        return BilinearForm('mass', self.function_space, self.material_functions['eps_r'])

    def get_stiffness_form(self):
        return BilinearForm('stiffness', self.function_space, self.material_functions['mu_r'])


class CombineGalerkinForms(object):
    """Base class for problem-specific form combination logic (Forms.py:88-116)."""

    def set_interior_forms(self, interior_forms):
        self.interior_forms = interior_forms

    def set_boundary_conditions(self, boundary_conditions):
        self.boundary_conditions = boundary_conditions

    def get_combined_forms(self):
        raise NotImplementedError("Use a concrete class")
