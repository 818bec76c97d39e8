"""Test meshes (mirror of /root/reference/sucemfem/Testing/Meshes.py:25-61)."""
import numpy as N

from .. import dolfin_compat as dolfin


class InscribedTet(object):
    def __init__(self):
        self.coordinates = N.array(
            [[-1. / 2, 1. / 2, -1. / 2], [1. / 2, 1. / 2, 1. / 2], [1. / 2, -1. / 2, -1. / 2],
             [-1. / 2, -1. / 2, 1. / 2], [1. / 6, 1. / 6, -1. / 6], [-1. / 6, 1. / 6, 1. / 6],
             [-1. / 6, -1. / 6, -1. / 6], [1. / 6, -1. / 6, 1. / 6]], N.float64)
        self.element_nodes = N.array(
            [[1, 4, 6, 7], [1, 3, 5, 7], [3, 4, 7, 8], [1, 2, 5, 6], [2, 3, 5, 8], [2, 4, 6, 8],
             [4, 6, 7, 8], [1, 5, 6, 7], [3, 5, 7, 8], [2, 5, 6, 8], [5, 6, 7, 8]], N.uint32) - 1

    def get_dolfin_mesh(self):
        return dolfin.Mesh(coordinates=self.coordinates, cells=self.element_nodes)
