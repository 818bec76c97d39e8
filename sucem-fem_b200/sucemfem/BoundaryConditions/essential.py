"""Essential (Dirichlet / PEC) boundary conditions
(mirror of /root/reference/sucemfem/BoundaryConditions/essential.py:25-99).

The reference builds dolfin.DirichletBC(V, zero, mesh_function, region) and returns its .apply (essential.py:77-81);
here the constrained dofs (closure of the marked facets) are flagged on the device and applied as row-only ident,
either as a separate step (apply) or fused into the value kernels (SystemMatrices).  Only the zero boundary value
(PEC) acts on matrices, which is all the matrix path of the reference uses.
"""
# Interface (module, class, method and argument names) after the file of cemagg/sucem-fem cited above:
# Copyright (C) 2011 Stellenbosch University, authors Neilen Marais, Evan Lezar; GNU GPL v3 or later
# (see NOTICE.md at the repository root).  The implementation below is this repository's own.
import numpy as np
import torch

from .. import dolfin_compat as dolfin
from .base_boundary_condition import BoundaryCondition


class EssentialBoundaryCondition(BoundaryCondition):
    def init_with_meshfunction(self, mesh_function, region_number):
        self.set_region_number(region_number)
        self.set_mesh_function(mesh_function)

    def set_PEC_expression(self):
        self.set_boundary_value_expression(('0.0',) * self.function_space.mesh().geometry().dim())

    def get_essential_row_flags(self, function_space, flags=None):
        space = function_space.device_space
        markers = torch.from_numpy(self.mesh_function.array().astype(np.int32)).to(space.mesh.device)
        return space.pec_row_flags(markers, self.region_number, flags)

    def get_essential_application_func(self, function_space=None):
        if function_space is not None:
            self.set_function_space(function_space)
        if self.boundary_value_expression is None:
            self.set_PEC_expression()
        flags = self.get_essential_row_flags(self.function_space)

        def apply(A, b=None):
            A.ident(flags)
            if b is not None:
                b[flags.cpu().numpy().astype(bool)] = 0.0
        return apply


class PECWallsBoundaryCondition(EssentialBoundaryCondition):
    """Essential boundary condition that models PEC walls on the whole exterior (essential.py:83-99)."""

    def init_with_mesh(self, mesh):
        mesh_function = dolfin.MeshFunction('uint', mesh, mesh.topology().dim() - 1)
        mesh_function.set_all(0)
        walls = dolfin.DomainBoundary()
        walls.mark(mesh_function, 999)
        self.init_with_meshfunction(mesh_function, 999)
