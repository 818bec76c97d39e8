"""High level current filament source (mirror of /root/reference/sucemfem/Sources/fillament_source.py:24-55)."""
from .fillament_current_source import FillamentCurrentSource
from ..Utilities.Geometry import unit_vector, vector_length


class FillamentSource(object):
    def __init__(self, function_space):
        self.function_space = function_space

    def set_source_parameters(self, source_parameters):
        """dict with keys I (current in amperes) and endpoints (2x3 array: start, end)"""
        self.source_parameters = source_parameters
        endpoints = source_parameters['endpoints']
        delta = endpoints[1] - endpoints[0]
        self.direction = unit_vector(delta)
        self.length = vector_length(delta)

    def get_current_source(self):
        cs = FillamentCurrentSource()
        cs.set_function_space(self.function_space)
        cs.set_source_endpoints(self.source_parameters['endpoints'])
        cs.set_value(self.source_parameters['I'])
        return cs
