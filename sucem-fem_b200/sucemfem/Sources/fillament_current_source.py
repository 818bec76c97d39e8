"""Current filament as an N-point sum of point dipoles.

Interface of cemagg/sucem-fem's sucemfem/Sources/fillament_current_source.py
(/root/reference/sucemfem/Sources/fillament_current_source.py:29-82; Copyright (C) 2011 Stellenbosch University, GPLv3,
author Neilen Marais): setters, `no_integration_points = 100`, `get_contribution() -> (dofnos, rhs_contribs)`.
Quadrature as in the reference (:62-82): N equally spaced points INCLUDING both end points, each weighted
I * l_hat * length / N (one point: the mid point).  Implementation is this repository's own.
"""
import numpy as np

from .current_source import CurrentSource, merge_contributions
from .point_source import calc_pointsource_contrib


class FillamentCurrentSource(CurrentSource):
    no_integration_points = 100

    def __init__(self, *names, **kwargs):
        self.source_endpoints = None
        self.value = None

    def set_no_integration_points(self, no_integration_points):
        self.no_integration_points = no_integration_points

    def set_source_endpoints(self, source_endpoints):
        """2x3 array: start and end point; conventional current flows from start to end."""
        self.source_endpoints = source_endpoints

    def set_value(self, value):
        """Line current in amperes."""
        self.value = value

    def _points_and_weight(self):
        start, end = np.asarray(self.source_endpoints, dtype=np.float64)
        delta = end - start
        n = int(self.no_integration_points)
        t = np.array([0.5]) if n == 1 else np.linspace(0.0, 1.0, n)
        # (I * l_hat) * length / n  ==  I * delta / n
        return start + t[:, None] * delta, delta / np.linalg.norm(delta) * self.value * np.linalg.norm(delta) / n

    def get_contribution(self):
        points, weight = self._points_and_weight()
        parts = [calc_pointsource_contrib(self.function_space, pt, weight) for pt in points]
        return merge_contributions([p[0] for p in parts], [p[1] for p in parts])
