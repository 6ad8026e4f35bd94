"""Manufactured Dirichlet data (perphil.utils.manufactured_solutions, utils/manufactured_solutions.py:7-94).

Same formulas as the reference, as pointwise numpy expressions of the node coordinates (K10 in
SURVEY 2.2: tiny host-side work that produces the values handed to dpp_set_dirichlet).
"""
from __future__ import annotations

import math

import numpy as np

from .mesh import Expression, Function
from .parameters import DPPParameters


def _pressures(prm: DPPParameters):
    k1, k2, beta, mu, eta = float(prm.k1), float(prm.k2), float(prm.beta), float(prm.mu), prm.eta
    pi = math.pi

    def common(X):
        if X.shape[1] == 2:  # :39-41
            return (mu / pi) * np.exp(pi * X[:, 0]) * np.sin(pi * X[:, 1])
        return (mu / pi) * np.exp(pi * X[:, 0]) * (np.sin(pi * X[:, 1]) + np.sin(pi * X[:, 2]))  # :82-88

    def e(X):
        if X.shape[1] == 2:
            return np.exp(eta * X[:, 1])
        return np.exp(eta * X[:, 1]) + np.exp(eta * X[:, 2])

    p1 = Expression(lambda X: common(X) - (mu / (beta * k1)) * e(X))
    p2 = Expression(lambda X: common(X) + (mu / (beta * k2)) * e(X))
    # closed-form tag: lets the error-norm kernel evaluate the expression (and its gradient) on the device
    p1.manufactured, p2.manufactured = (prm, 0), (prm, 1)
    return p1, p2


def exact_expressions(mesh, dpp_params: DPPParameters):
    """(u1, p1, u2, p2); velocities are post-processing (out of scope) and returned as None."""
    p1, p2 = _pressures(dpp_params)
    return None, p1, None, p2


def exact_expressions_3d(mesh, dpp_params: DPPParameters):
    return exact_expressions(mesh, dpp_params)


def interpolate_exact(mesh, velocity_space, pressure_space, dpp_params: DPPParameters):
    _, p1, _, p2 = exact_expressions(mesh, dpp_params)
    return (None, Function(pressure_space, name="p1_exact").interpolate(p1), None,
            Function(pressure_space, name="p2_exact").interpolate(p2))
