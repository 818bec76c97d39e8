"""Mesh-entity marking helpers (mirror of /root/reference/sucemfem/Geometry.py:24-118): boundary edges and the cells
connected to them.  The reference walks DOLFIN's connectivity with Python iterators; here the entity numbering comes
from the device (K1) and the marking is a couple of array operations."""
# Interface (module, class, method and argument names) after the file of cemagg/sucem-fem cited above:
# Copyright (C) 2011 Stellenbosch University, authors Neilen Marais; GNU GPL v3 or later
# (see NOTICE.md at the repository root).  The implementation below is this repository's own.
import numpy as np
import torch

from . import dolfin_compat as dolfin


class BoundaryEdges(object):
    def __init__(self, mesh, boundary_facefun=None, boundary_value=1):
        """Find and mark the edges that occur on a set of boundary faces (default: all exterior faces)."""
        self.mesh = mesh
        if boundary_facefun is None:
            boundary_facefun = dolfin.FaceFunction('uint', mesh)
            boundary_facefun.set_all(0)
            dolfin.DomainBoundary().mark(boundary_facefun, boundary_value)
        self.boundary_facefun = boundary_facefun
        self.boundary_value = boundary_value

    def mark(self, edge_meshfun, value):
        sp = self.mesh._numbering()          # order-1 space: one dof per edge, so PEC row flags are edge flags
        markers = torch.from_numpy(self.boundary_facefun.array().astype(np.int32)).to(sp.mesh.device)
        flags = sp.pec_row_flags(markers, self.boundary_value).cpu().numpy().astype(bool)
        edge_meshfun.array()[flags] = value


class BoundaryEdgeCells(object):
    """Find and mark cells connected the domain boundary through edges"""

    def __init__(self, mesh):
        self.edge_boundary_value = 1
        boundary_edgefun = dolfin.EdgeFunction('uint', mesh)
        boundary_edgefun.set_all(0)
        BoundaryEdges(mesh).mark(boundary_edgefun, self.edge_boundary_value)
        self.boundary_edgefun = boundary_edgefun

    def mark(self, cell_meshfun, value):
        CellsConnected2Edges(self.boundary_edgefun, self.edge_boundary_value).mark(cell_meshfun, value)


class CellsConnected2Edges(object):
    def __init__(self, boundary_edgefun, boundary_value):
        self.boundary_edgefun = boundary_edgefun
        self.boundary_value = boundary_value
        self.mesh = boundary_edgefun.mesh

    def mark(self, cell_meshfun, value):
        cell_edges = self.mesh._numbering().cell_entities()[0].cpu().numpy()
        hit = (self.boundary_edgefun.array() == self.boundary_value)[cell_edges].any(axis=1)
        cell_meshfun.array()[hit] = value
