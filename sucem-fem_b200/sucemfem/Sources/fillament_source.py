"""Current filament described by a parameter dict -> FillamentCurrentSource.

Interface of cemagg/sucem-fem's sucemfem/Sources/fillament_source.py (/root/reference/sucemfem/Sources/fillament_source.py:24-55;
Copyright (C) 2011 Stellenbosch University, GPLv3, author Neilen Marais): `FillamentSource(function_space)`,
`set_source_parameters({'I': amperes, 'endpoints': 2x3 array})`, attributes `direction` / `length`,
`get_current_source()`.  Implementation is this repository's own.
"""
import numpy as np

from .fillament_current_source import FillamentCurrentSource


class FillamentSource(object):
    def __init__(self, function_space):
        self.function_space = function_space
        self.source_parameters = None
        self.direction = None
        self.length = None

    def set_source_parameters(self, source_parameters):
        start, end = np.asarray(source_parameters['endpoints'], dtype=np.float64)
        self.source_parameters = source_parameters
        self.length = float(np.linalg.norm(end - start))
        self.direction = (end - start) / self.length

    def get_current_source(self):
        src = FillamentCurrentSource()
        src.set_function_space(self.function_space)
        src.set_source_endpoints(self.source_parameters['endpoints'])
        src.set_value(self.source_parameters['I'])
        return src
