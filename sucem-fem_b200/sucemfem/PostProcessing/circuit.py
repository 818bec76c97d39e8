"""Circuit quantities (mirror of /root/reference/sucemfem/PostProcessing/circuit.py:5-12)."""
import numpy as np


def S11(Zl, Z0):
    """Reflection coefficient of a load Zl on a line of characteristic impedance Z0"""
    Zl = np.asarray(Zl)
    return (Zl - Z0) / (Zl + Z0)
