"""Mesh generators (mirror of /root/reference/sucemfem/Utilities/MeshGenerators.py:26-53)."""
# Interface (module, class, method and argument names) after the file of cemagg/sucem-fem cited above:
# Copyright (C) 2011 Stellenbosch University, authors Neilen Marais; GNU GPL v3 or later
# (see NOTICE.md at the repository root).  The implementation below is this repository's own.
import numpy as N

from .. import dolfin_compat as dolfin


def get_centred_cube(domain_size, max_edge_len, centred_element_coordinate=None):
    """Generate a cube mesh centred around [0,0,0]; optionally translate it slightly such that
    centred_element_coordinate is at the centroid of an element (the lowest-indexed cell containing it)."""
    domain_size = N.atleast_1d(N.asarray(domain_size, dtype=N.float64))
    if domain_size.size == 1:
        domain_size = N.repeat(domain_size, 3)
    domain_subdivisions = N.array(N.ceil(N.sqrt(2) * domain_size / max_edge_len), N.uint)
    mesh = dolfin.UnitCube(*domain_subdivisions)
    mesh.coordinates()[:] *= domain_size
    mesh.coordinates()[:] -= domain_size / 2
    if centred_element_coordinate is not None:
        pt = N.asarray(centred_element_coordinate, dtype=N.float64)
        x = mesh.coordinates()[mesh.cells()]                      # [C,4,3]
        J = N.swapaxes(x[:, 1:, :] - x[:, :1, :], 1, 2)
        lam = N.linalg.solve(J, (pt[None, :] - x[:, 0, :])[:, :, None])[:, :, 0]
        inside = (lam.min(axis=1) >= -1e-12) & (1 - lam.sum(axis=1) >= -1e-12)
        hits = N.nonzero(inside)[0]
        mid = x[hits].mean(axis=1)
        closest = hits[N.argmin(N.sqrt(((mid - pt) ** 2).sum(axis=1)))]
        mesh.coordinates()[:] -= x[closest].mean(axis=0)
    return mesh
