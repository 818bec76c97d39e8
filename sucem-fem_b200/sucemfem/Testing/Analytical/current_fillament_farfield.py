"""Far field of a straight current filament with uniform current (mirror of
/root/reference/sucemfem/Testing/Analytical/current_fillament_farfield.py:27-44, which evaluates the same expression
symbolically with sympy):

    E_theta = j w mu0 I0 L exp(-j beta r) / (4 pi r) * sin(theta) * sinc_u,   u = (beta L / 2) cos(theta),
    sinc_u = sin(u) / u   (-> 1 at theta = pi/2, where the symbolic form is 0/0).
"""
import numpy as np

from ...Consts import mu0, c0


def eval_E_theta(freq, L_val, I_val, theta_val, r_val=1):
    """theta component of the radiated field at distance r_val (default 1 m); theta_val in radians, scalar or array."""
    theta = np.asarray(theta_val, dtype=np.float64)
    omega = 2 * np.pi * freq
    beta = omega / c0
    u = 0.5 * beta * L_val * np.cos(theta)
    pattern = np.sin(theta) * np.sinc(u / np.pi)          # numpy's sinc(x) = sin(pi x) / (pi x)
    val = 1j * omega * mu0 * I_val * L_val * np.exp(-1j * beta * r_val) / (4 * np.pi * r_val) * pattern
    return complex(val) if val.ndim == 0 else val
