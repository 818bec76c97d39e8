"""Voltage along a straight line through a computed field (mirror of /root/reference/sucemfem/Sources/PostProcess.py:27-63;
reference test: sucemfem/Sources/tests/test_postprocessing.py:30-77).

The reference wraps the dof vector in a dolfin.Function and integrates E . l_hat with scipy's Romberg rule.  Here the
field is evaluated on the host from the nodal basis of the cell that contains each point (PostProcessing.Reconstruct)
and integrated with adaptive Gauss-Kronrod quadrature: the integrand is piecewise polynomial with kinks where the line
crosses cell faces.  O(#quadrature points) host work, outside the assembly hot path.
"""
import numpy as np
import scipy.integrate

from ..Utilities.Geometry import unit_vector, vector_length


class FieldFunction(object):
    """Callable stand-in for dolfin.Function(function_space, vector): fn(x, y, z) -> field vector at the point."""

    def __init__(self, function_space, dofs):
        from ..PostProcessing import Reconstruct
        self._rec = Reconstruct(function_space)
        self._rec.set_dof_values(np.asarray(dofs))

    def __call__(self, *point):
        return self._rec.reconstruct_points(np.array([point], dtype=np.float64))[0]


class VoltageAlongLine(object):
    """Measure voltage along a straight line between two points"""
    epsabs, epsrel, limit = 1e-12, 1e-10, 400

    def __init__(self, field_function):
        self.field_function = field_function

    def calculate_voltage(self, start_pt, end_pt):
        fn = self.field_function
        start_pt, end_pt = np.asarray(start_pt, dtype=np.float64), np.asarray(end_pt, dtype=np.float64)
        delta = end_pt - start_pt
        l_hat = unit_vector(delta)
        # E . l_hat over the unit interval, then de-normalised with the length of the path
        eval_fn = lambda l: float(np.dot(l_hat, np.real(fn(*(start_pt + delta * l)))))
        intg, _ = scipy.integrate.quad(eval_fn, 0.0, 1.0, epsabs=self.epsabs, epsrel=self.epsrel, limit=self.limit)
        return intg * vector_length(delta)


class ComplexVoltageAlongLine(object):
    """Measure complex voltage along a straight line between two points"""

    def __init__(self, function_space):
        self.function_space = function_space

    def set_dofs(self, dofs):
        self.dofs = np.asarray(dofs)
        self.E_r = FieldFunction(self.function_space, self.dofs.real)
        self.E_i = FieldFunction(self.function_space, self.dofs.imag)
        self.real_voltage = VoltageAlongLine(self.E_r)
        self.imag_voltage = VoltageAlongLine(self.E_i)

    def calculate_voltage(self, start_pt, end_pt):
        return (self.real_voltage.calculate_voltage(start_pt, end_pt) +
                1j * self.imag_voltage.calculate_voltage(start_pt, end_pt))
