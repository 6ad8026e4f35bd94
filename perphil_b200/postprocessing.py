"""Post-processing of a DPP solution (perphil.utils.postprocessing, utils/postprocessing.py:6-124):
`split_dpp_solution`, `l2_error`, `h1_seminorm_error`.  The error integrals run on the GPU
(csrc/error_norms.cu, `dpp_error_norms`); SURVEY 8(f) item 1.  The Darcy-velocity projection and the
slicing helper of the reference stay out of scope (SURVEY 8f items 3+)."""
from __future__ import annotations

from typing import Tuple

import numpy as np

from .mesh import Expression, Function


def split_dpp_solution(dpp_solution: Function) -> Tuple[Function, Function]:
    """utils/postprocessing.py:6-31."""
    W = dpp_solution.function_space()
    if not hasattr(W, "num_sub_spaces") or W.num_sub_spaces() != 2:
        raise ValueError(f"Expected a 2-field MixedFunctionSpace, got {type(W)}")
    p1 = Function(W.sub(0), name="p1_h", val=np.array(dpp_solution.sub(0).dat.data, copy=True))
    p2 = Function(W.sub(1), name="p2_h", val=np.array(dpp_solution.sub(1).dat.data, copy=True))
    return p1, p2


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
