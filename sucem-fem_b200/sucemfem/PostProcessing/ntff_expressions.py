"""Direction vectors of the far-field point (mirror of /root/reference/sucemfem/PostProcessing/ntff_expressions.py:24-39)."""
import numpy as np


def r_hat(theta, phi):
    return np.array([np.sin(theta) * np.cos(phi), np.sin(theta) * np.sin(phi), np.cos(theta)])


def theta_hat(theta, phi):
    return np.array([np.cos(theta) * np.cos(phi), np.cos(theta) * np.sin(phi), -np.sin(theta)])


def phi_hat(theta, phi):
    return np.array([-np.sin(phi), np.cos(phi), 0.])


def phase(k0, rprime, rhat):
    return k0 * np.dot(rprime, rhat)
