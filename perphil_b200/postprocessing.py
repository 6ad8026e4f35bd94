"""Post-processing of a DPP solution (perphil.utils.postprocessing, utils/postprocessing.py:6-124):
`split_dpp_solution`, `calculate_darcy_velocity_from_pressure`, `l2_error`, `h1_seminorm_error`.  The error
integrals (csrc/error_norms.cu, `dpp_error_norms`) and the Darcy-velocity projection
(csrc/darcy.cu, `dpp_darcy_velocity`; item 3) run on the GPU; `slice_along_x` (a plotting helper: point evaluation
along a line) is host numpy."""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np

from .mesh import Expression, Function, _Dat


def split_dpp_solution(dpp_solution: Function) -> Tuple[Function, Function]:
    """utils/postprocessing.py:6-31."""
    W = dpp_solution.function_space()
    if not hasattr(W, "num_sub_spaces") or W.num_sub_spaces() != 2:
        raise ValueError(f"Expected a 2-field MixedFunctionSpace, got {type(W)}")
    p1 = Function(W.sub(0), name="p1_h", val=np.array(dpp_solution.sub(0).dat.data, copy=True))
    p2 = Function(W.sub(1), name="p2_h", val=np.array(dpp_solution.sub(1).dat.data, copy=True))
    return p1, p2


class VectorFunction:
    """What `fd.Function(VectorFunctionSpace(mesh, "CG", p))` is to the callers of the projection: nodal
    values `dat.data` of shape [n_nodes, dim] (Firedrake's layout for vector CG spaces)."""

    def __init__(self, space, values: np.ndarray, name: str = "velocity"):
        self._space = space
        self.dat = _Dat(values)
        self._name = name
        self.cg_iterations = None

    def function_space(self):
        return self._space

    def name(self):
        return self._name

    def sub(self, i: int) -> Function:
        return Function(self._space, name=f"{self._name}[{i}]", val=np.array(self.dat.data[:, i], copy=True))


def calculate_darcy_velocity_from_pressure(pressure_field: Function, conductivity, velocity_space=None,
                                           degree: Optional[int] = None, rtol: float = 1e-8) -> VectorFunction:
    """utils/postprocessing.py:34-63: project u = -k grad(p_h) into the vector CG space.  `conductivity` is a
    float / Constant.  The velocity space is the vector version of the pressure's own Lagrange space (the
    reference's default `degree=1` with its default degree-1 pressure space); another degree raises."""
    from .solver import handle_for

    V = pressure_field.function_space()
    W = getattr(V, "parent", None)
    if W is None:
        raise ValueError("the pressure must live on W.sub(i) (use split_dpp_solution)")
    h = handle_for(W)
    p_deg = int(h.degree)
    if degree is not None and int(degree) != p_deg:
        raise NotImplementedError(f"velocity degree {degree} != pressure degree {p_deg}: the B200 projection uses "
                                  "the pressure's own Lagrange space")
    if velocity_space is not None and velocity_space is not V:
        raise NotImplementedError("pass velocity_space=None (vector version of the pressure space)")
    vel, its = h.darcy_velocity(float(conductivity), p=np.asarray(pressure_field.dat.data, dtype=np.float64), rtol=rtol)
    out = VectorFunction(V, np.ascontiguousarray(vel.T), name="velocity")
    out.cg_iterations = [int(i) for i in its]
    return out


def slice_along_x(scalar_field: Function, x_value: float) -> Tuple[np.ndarray, np.ndarray]:
    """utils/postprocessing.py:66-86: sample a scalar field along the vertical line x = x_value, at the distinct
    y-coordinates of the space's nodes (`np.unique` of the interpolated y coordinate there) -- the
    `scalar_field.at((x, y))` point evaluation of the reference, done on the host for the 2-D tensor-product
    spaces of this package (a plotting helper, not part of the GPU path)."""
    V = scalar_field.function_space()
    mesh = V.mesh()
    if mesh.dim != 2:
        raise NotImplementedError("slice_along_x samples 2-D fields (the reference uses it for the 2-D notebooks)")
    p = int(V.degree)
    ax, ay = mesh.local_axes(p)
    nx, ny = V.grid_nodes
    u = np.asarray(scalar_field.dat.data, dtype=float).reshape(nx, ny)
    xv = ax[::p]                                   # vertex coordinates along x
    if not (xv[0] - 1e-14 <= x_value <= xv[-1] + 1e-14):
        raise ValueError(f"x = {x_value} is outside the mesh")
    c = int(min(max(np.searchsorted(xv, x_value, side="right") - 1, 0), xv.size - 2))
    xi = (x_value - xv[c]) / (xv[c + 1] - xv[c])
    nodes = np.linspace(0.0, 1.0, p + 1)
    vals = np.zeros(ny)
    for a in range(p + 1):
        N = 1.0
        for m in range(p + 1):
            if m != a:
                N *= (xi - nodes[m]) / (nodes[a] - nodes[m])
        vals += N * u[p * c + a, :]
    return np.asarray(ay, dtype=float).copy(), vals


def _norms(numerical: Function, exact, nq: int):
    from .solver import handle_for

    V = numerical.function_space()
    W = getattr(V, "parent", None)
    field = getattr(V, "index", None)
    if W is None or field is None:
        raise ValueError("error norms need a Function on W.sub(i) (use split_dpp_solution)")
    h = handle_for(W)
    n = h.n_nodes
    u = np.zeros(2 * n)
    u[field * n:(field + 1) * n] = numerical.dat.data
    man = getattr(exact, "manufactured", None)
    if man is not None:  # closed form evaluated on the device with the parameters the expression was built from
        prm, f_expr = man
        if f_expr != field:
            raise ValueError("the exact expression belongs to the other pressure field")
        h.set_params(float(prm.k1), float(prm.k2), float(prm.beta), float(prm.mu))
        out = h.error_norms(u, None, nq)
    elif isinstance(exact, Function):
        e = np.zeros(2 * n)
        e[field * n:(field + 1) * n] = exact.dat.data
        out = h.error_norms(u, e, nq)
    else:
        raise NotImplementedError("exact must be a manufactured expression (exact_expressions) or a Function of the "
                                  "same space; arbitrary host callables cannot be evaluated on the device")
    return out[field], out[2 + field]


def l2_error(numerical: Function, exact_expr, quadrature_points: int = 6) -> float:
    """||numerical - exact||_L2 (utils/postprocessing.py:89-105)."""
    return float(_norms(numerical, exact_expr, quadrature_points)[0])


def h1_seminorm_error(numerical: Function, exact_expr, quadrature_points: int = 6) -> float:
    """|numerical - exact|_H1 (utils/postprocessing.py:108-124)."""
    return float(_norms(numerical, exact_expr, quadrature_points)[1])
