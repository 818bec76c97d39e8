"""Current sources container and abstract base class
(mirror of /root/reference/sucemfem/Sources/current_source.py:26-65)."""
import collections

import numpy as N


class CurrentSources(object):
    def __init__(self):
        self.sources = []

    def set_function_space(self, function_space):
        self.function_space = function_space

    def add_source(self, source):
        self.sources.append(source)

    def init_sources(self):
        for src in self.sources:
            src.set_function_space(self.function_space)

    def get_source_contributions(self):
        """Get and return the RHS contribution of the current sources"""
        contribs = collections.OrderedDict()
        for src in self.sources:
            dofnos, values = src.get_contribution()
            for dn, v in zip(dofnos, values):
                contribs[int(dn)] = contribs.get(int(dn), 0.) + v
        dofnos = N.array(list(contribs.keys()), dtype=N.uint)
        rhs_contribs = N.array(list(contribs.values()))
        return dofnos, rhs_contribs


class CurrentSource(object):
    """Abstract base class for current sources"""

    def set_function_space(self, function_space):
        self.function_space = function_space

    def get_contribution(self):
        raise NotImplementedError('User subclass should implement get_contribution()')
