"""Current filament as an N-point sum of point sources
(mirror of /root/reference/sucemfem/Sources/fillament_current_source.py:29-82)."""
import collections

import numpy as np

from .current_source import CurrentSource
from .point_source import calc_pointsource_contrib
from ..Utilities.Geometry import unit_vector, vector_length


class FillamentCurrentSource(CurrentSource):
    no_integration_points = 100

    def __init__(self, *names, **kwargs):
        self._dirty = True

    def set_no_integration_points(self, no_integration_points):
        self.no_integration_points = no_integration_points

    def set_source_endpoints(self, source_endpoints):
        """2x3 array with the start and end point; conventional current flows from start to end."""
        self.source_endpoints = source_endpoints
        self._dirty = True

    def set_value(self, value):
        """Set line current value in Amperes"""
        self.value = value
        self._dirty = True

    def _update(self):
        if self._dirty:
            self.source_start, self.source_end = self.source_endpoints
            self.source_delta = self.source_end - self.source_start
            self.vector_value = unit_vector(self.source_delta) * self.value
        self._dirty = False

    def get_contribution(self):
        self._update()
        source_len = vector_length(self.source_delta)
        no_pts = self.no_integration_points
        if no_pts == 1:
            intg_pts = [self.source_start + self.source_delta * 0.5]
        else:
            intg_pts = self.source_start + self.source_delta * np.linspace(0, 1, no_pts)[:, np.newaxis]
        contribs = collections.OrderedDict()
        point_magnitude = self.vector_value * source_len / no_pts
        for pt in intg_pts:
            dnos, vals = calc_pointsource_contrib(self.function_space, pt, point_magnitude)
            for dn, v in zip(dnos, vals):
                contribs[int(dn)] = contribs.get(int(dn), 0.) + v
        dofnos = np.array(list(contribs.keys()), dtype=np.uint)
        rhs_contribs = np.array(list(contribs.values()))
        return dofnos, rhs_contribs
