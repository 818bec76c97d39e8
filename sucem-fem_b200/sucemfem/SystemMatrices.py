"""The drop-in boundary of the path (mirror of /root/reference/sucemfem/SystemMatrices.py:26-58).

calc_system_matrices() keeps the reference's contract -- for every named form: NullForm -> None, otherwise a
MatrixClass() instance assembled from the form with the essential boundary conditions applied -- but when both the
mass and the stiffness form of one function space are requested it fills them with ONE fused launch of the S+T
kernels, PEC rows included, instead of two assemblies followed by two ident passes.
"""
# Interface (module, class, method and argument names) after the file of cemagg/sucem-fem cited above:
# Copyright (C) 2011 Stellenbosch University, authors Neilen Marais; GNU GPL v3 or later
# (see NOTICE.md at the repository root).  The implementation below is this repository's own.
from . import Forms
from . import dolfin_compat as dolfin
from . import _native as N


class SystemMatrices(object):
    MatrixClass = dolfin.PETScMatrix

    def set_matrix_class(self, matrix_class):
        """Overrides the default matrix class set in the class definition"""
        self.MatrixClass = matrix_class

    def set_matrix_forms(self, matrix_forms):
        """Set matrix_forms with a dict mapping matrix names to bilinear forms"""
        self.matrix_forms = matrix_forms

    def set_boundary_conditions(self, boundary_conditions):
        """Set boundary_conditions with instance of BoundaryConditions"""
        self.boundary_conditions = boundary_conditions

    def calc_system_matrices(self):
        """Calculate and return system matrices in a dict"""
        system_matrices = dict()
        forms = dict(self.matrix_forms)
        mass = [(k, f) for k, f in forms.items() if isinstance(f, Forms.BilinearForm) and f.kind == 'mass']
        stiff = [(k, f) for k, f in forms.items() if isinstance(f, Forms.BilinearForm) and f.kind == 'stiffness']
        abc = [(k, f) for k, f in forms.items() if isinstance(f, Forms.BilinearForm) and f.kind == 'abc']
        same_space = (len(mass) == 1 and len(stiff) == 1 and mass[0][1].function_space is stiff[0][1].function_space
                      and all(f.function_space is mass[0][1].function_space for _, f in abc))
        if same_space and issubclass(self.MatrixClass, dolfin.HostCSRMatrix):
            # host matrix class: stream row blocks to page-locked host arrays while they are assembled
            (mname, mform), (sname, sform) = mass[0], stiff[0]
            space = mform.function_space.device_space
            flags = self.boundary_conditions.get_essential_row_flags(mform.function_space)
            eps = None if mform.coefficient is None else mform.coefficient.device_values()
            mu = None if sform.coefficient is None else sform.coefficient.device_values()
            want = ('cols', 'S', 'T') + (('S0',) if len(abc) == 1 else ())
            host = space.assemble_to_host(eps, mu, flags, want=want)
            named = [(sname, 'S'), (mname, 'T')] + ([(abc[0][0], 'S0')] if len(abc) == 1 else [])
            for name, key in named:
                mat = self.MatrixClass()
                mat.init_host(space.ndof, host['rowptr'], host['cols'], host[key])
                system_matrices[name] = mat
                del forms[name]
        elif (len(mass) == 1 and len(stiff) == 1 and mass[0][1].function_space is stiff[0][1].function_space
                and issubclass(self.MatrixClass, dolfin.DeviceCSRMatrix)):
            (mname, mform), (sname, sform) = mass[0], stiff[0]
            space = mform.function_space.device_space
            pat = space.pattern()
            flags = self.boundary_conditions.get_essential_row_flags(mform.function_space)
            eps = None if mform.coefficient is None else mform.coefficient.device_values()
            mu = None if sform.coefficient is None else sform.coefficient.device_values()
            S, T = pat.assemble_ST(eps, mu, flags, kernel=N.KERNEL_AUTO)
            for name, vals in ((sname, S), (mname, T)):
                mat = self.MatrixClass()
                mat.init(pat, vals)
                system_matrices[name] = mat
                del forms[name]
        for matname, form in forms.items():
            mat = self.MatrixClass()
            if isinstance(form, Forms.NullForm):
                mat = None
            else:
                dolfin.assemble(form, tensor=mat)
                self.boundary_conditions.apply_essential(mat)
            system_matrices[matname] = mat
        return system_matrices
