"""Manufactured Dirichlet data (perphil.utils.manufactured_solutions, utils/manufactured_solutions.py:7-94).

Same formulas as the reference, as pointwise numpy expressions of the node coordinates (K10 in
SURVEY 2.2: tiny host-side work that produces the values handed to dpp_set_dirichlet).  The exact
Darcy velocities u_i = -(k_i/mu) grad p_i (:21-37 in 2-D, :72-81 in 3-D) are `VectorExpression`s: what
the projected velocity of `calculate_darcy_velocity_from_pressure` converges to.
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


class VectorExpression:
    """Pointwise vector-valued expression: coords [n, dim] -> values [n, dim] (fd.as_vector of the reference)."""

    def __init__(self, fn):
        self._fn = fn

    def __call__(self, coords: np.ndarray) -> np.ndarray:
        return np.asarray(self._fn(np.atleast_2d(np.asarray(coords, dtype=float))), dtype=float)


def _velocities(prm: DPPParameters):
    k1, k2, beta, mu, eta = float(prm.k1), float(prm.k2), float(prm.beta), float(prm.mu), prm.eta
    pi = math.pi

    def make(k, sign):
        # 2-D (:21-37): -k (e^{pi x} sin(pi y), e^{pi x} cos(pi y) -+ (eta/(beta k)) e^{eta y})
        # 3-D (:72-81): -(k/mu) grad p with p of :82-88
        def u(X):
            ex = np.exp(pi * X[:, 0])
            if X.shape[1] == 2:
                return np.stack([-k * ex * np.sin(pi * X[:, 1]),
                                 -k * (ex * np.cos(pi * X[:, 1]) + sign * (eta / (beta * k)) * np.exp(eta * X[:, 1]))],
                                axis=1)
            s = np.sin(pi * X[:, 1]) + np.sin(pi * X[:, 2])
            gx = mu * ex * s
            gy = mu * ex * np.cos(pi * X[:, 1]) + sign * (mu * eta / (beta * k)) * np.exp(eta * X[:, 1])
            gz = mu * ex * np.cos(pi * X[:, 2]) + sign * (mu * eta / (beta * k)) * np.exp(eta * X[:, 2])
            return -(k / mu) * np.stack([gx, gy, gz], axis=1)

        return VectorExpression(u)

    return make(k1, -1.0), make(k2, +1.0)


def exact_expressions(mesh, dpp_params: DPPParameters):
    """(u1, p1, u2, p2) of utils/manufactured_solutions.py:7-51 (2-D) / :54-94 (3-D, picked by the coordinates)."""
    p1, p2 = _pressures(dpp_params)
    u1, u2 = _velocities(dpp_params)
    return u1, p1, u2, p2


def exact_expressions_3d(mesh, dpp_params: DPPParameters):
    return exact_expressions(mesh, dpp_params)


def interpolate_exact(mesh, velocity_space, pressure_space, dpp_params: DPPParameters):
    """utils/manufactured_solutions.py:97-135.  `velocity_space` None: the vector version of the pressure space
    (what calculate_darcy_velocity_from_pressure projects into)."""
    from .postprocessing import VectorFunction

    u1, p1, u2, p2 = exact_expressions(mesh, dpp_params)
    Vv = velocity_space if velocity_space is not None else pressure_space
    X = Vv.node_coordinates
    return (VectorFunction(Vv, u1(X), name="u1_exact"), Function(pressure_space, name="p1_exact").interpolate(p1),
            VectorFunction(Vv, u2(X), name="u2_exact"), Function(pressure_space, name="p2_exact").interpolate(p2))
