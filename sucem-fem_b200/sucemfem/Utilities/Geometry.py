"""Vector helpers (mirror of /root/reference/sucemfem/Utilities/Geometry.py)."""
import numpy as np


def unit_vector(vec):
    return vec / vector_length(vec)


def vector_length(vec):
    return np.sqrt(np.sum(vec ** 2))
