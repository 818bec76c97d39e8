"""Material regions -> per-cell coefficient functions (mirror of /root/reference/sucemfem/Materials.py:28-150).

The reference expands region numbers with a Python list comprehension over all cells (Materials.py:145-146); here
the look-up runs on the device (sfem_expand_material) and the result stays in HBM for the assembly kernels.
"""
# Interface (module, class, method and argument names) after the file of cemagg/sucem-fem cited above:
# Copyright (C) 2011 Stellenbosch University, authors Neilen Marais; GNU GPL v3 or later
# (see NOTICE.md at the repository root).  The implementation below is this repository's own.
import numpy as np
import torch

from . import Consts, device

mu0 = Consts.mu0
eps0 = Consts.eps0


class MaterialProperties(object):
    allowed_properties_defaults = {'eps_r': 1.0, 'mu_r': 1.0}

    def __init__(self):
        for prop, val in self.allowed_properties_defaults.items():
            setattr(self, prop, val)

    def init_values(self, **kwargs):
        for prop, val in kwargs.items():
            if prop not in self.allowed_properties_defaults:
                raise ValueError('Unknown material property')
            setattr(self, prop, val)

    def set_eps_r(self, eps_r):
        self.eps_r = eps_r

    def set_mu_r(self, mu_r):
        self.mu_r = mu_r

    def get_eps_r(self):
        return self.eps_r

    def get_mu_r(self):
        return self.mu_r

    def get_mu_r_inv(self):
        return 1 / self.get_mu_r()

    def get_eps(self):
        return eps0 * self.eps_r

    def get_mu(self):
        return mu0 * self.mu_r


class MaterialPropertiesFactory(object):
    def __init__(self, material_regions):
        """material_regions: {region_no: {'eps_r': .., 'mu_r': ..}}; None -> region 0 with default properties."""
        self.material_properties = {}
        if material_regions is None:
            self.material_properties[0] = MaterialProperties()
        else:
            for region_no, property_values in material_regions.items():
                matprop = MaterialProperties()
                matprop.init_values(**property_values)
                self.material_properties[region_no] = matprop

    def get_material_properties(self):
        return self.material_properties


class MaterialFunction(object):
    """Element-wise constant (DG0) coefficient; values resident on the device."""

    def __init__(self, values):
        self._values = values

    def device_values(self):
        return self._values

    def vector(self):
        return self._values.cpu().numpy()


class MaterialFunctionFactory(object):
    def __init__(self, region_material_properties, region_meshfunction, mesh):
        self.region_material_properties = region_material_properties
        self.region_meshfunction = region_meshfunction      # None: all cells in region 0
        self.mesh = mesh

    def get_material_functions(self, *property_names):
        mat_fns = {}
        mats = self.region_material_properties
        if self.region_meshfunction is None:
            regions = torch.zeros(self.mesh.num_cells(), dtype=torch.int32, device='cuda')
        else:
            regions = torch.from_numpy(self.region_meshfunction.array().astype(np.int32)).cuda()
        for pname in property_names:
            valfunc = 'get_' + pname
            region_valuemap = dict((int(k), float(getattr(mat, valfunc)())) for k, mat in mats.items())
            # raises ValueError('Material number %d not found') like Materials.py:147-148
            mat_fns[pname] = MaterialFunction(device.expand_material(regions, region_valuemap))
        return mat_fns
