"""Direction vectors of the far-field point (mirror of /root/reference/sucemfem/PostProcessing/ntff_expressions.py:24-39)."""
# Interface (module, class, method and argument names) after the file of cemagg/sucem-fem cited above:
# Copyright (C) 2011 Stellenbosch University, authors Neilen Marais; GNU GPL v3 or later
# (see NOTICE.md at the repository root).  The implementation below is this repository's own.
import numpy as np


def r_hat(theta, phi):
    return np.array([np.sin(theta) * np.cos(phi), np.sin(theta) * np.sin(phi), np.cos(theta)])


def theta_hat(theta, phi):
    return np.array([np.cos(theta) * np.cos(phi), np.cos(theta) * np.sin(phi), -np.sin(theta)])


def phi_hat(theta, phi):
    return np.array([-np.sin(phi), np.cos(phi), 0.])


def phase(k0, rprime, rhat):
    return k0 * np.dot(rprime, rhat)
