"""Current sources: the container the driven problem asks for its RHS, and the source base class.

Interface of cemagg/sucem-fem's sucemfem/Sources/current_source.py (/root/reference/sucemfem/Sources/current_source.py:26-65;
Copyright (C) 2011 Stellenbosch University, GPLv3, author Neilen Marais) kept so that reference drivers run unchanged:
class and method names, the (dofnos, rhs_contribs) return convention and the order in which dofs first appear.
The bodies are this repository's own (numpy accumulation instead of a per-entry dict loop).
"""
import numpy as np


def merge_contributions(dof_lists, value_lists):
    """Sum values that hit the same dof; dofs are returned in order of first appearance (what an insertion-ordered
    accumulation gives), values in the common dtype of the inputs."""
    if not dof_lists:
        return np.zeros(0, dtype=np.uint), np.zeros(0)
    dofs = np.concatenate([np.asarray(d, dtype=np.int64).ravel() for d in dof_lists])
    vals = np.concatenate([np.asarray(v).ravel() for v in value_lists])
    uniq, first, inv = np.unique(dofs, return_index=True, return_inverse=True)
    order = np.argsort(first, kind='stable')
    rank = np.empty_like(order)
    rank[order] = np.arange(order.size)
    out = np.zeros(uniq.size, dtype=vals.dtype)
    np.add.at(out, rank[inv], vals)          # in-order accumulation per dof, like the reference's running sum
    return uniq[order].astype(np.uint), out


class CurrentSource(object):
    """Base class: a source knows its function space and returns (dofnos, rhs_contribs)."""
    function_space = None

    def set_function_space(self, function_space):
        self.function_space = function_space

    def get_contribution(self):
        raise NotImplementedError('User subclass should implement get_contribution()')


class CurrentSources(object):
    def __init__(self):
        self.sources = []
        self.function_space = None

    def set_function_space(self, function_space):
        self.function_space = function_space

    def add_source(self, source):
        self.sources.append(source)

    def init_sources(self):
        for source in self.sources:
            source.set_function_space(self.function_space)

    def get_source_contributions(self):
        """(dofnos, rhs_contribs) of all sources together: RHS[dofnos] += rhs_contribs."""
        parts = [source.get_contribution() for source in self.sources]
        return merge_contributions([p[0] for p in parts], [p[1] for p in parts])
