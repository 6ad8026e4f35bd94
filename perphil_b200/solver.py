"""`solve_dpp` / `solve_dpp_nonlinear` / `Solution`: the drop-in boundary of the hot path
(perphil.solvers.solver, solvers/solver.py:14-128), served by libdppb200 on a B200.

Routing: a `solver_parameters` dict that carries ``"dpp_backend": "b200"`` runs on the GPU;
anything else is handed, untouched, to the reference implementation when perphil + Firedrake are
importable (that is routing, not a fallback: nothing of the B200 path ever runs on the CPU).
"""
from __future__ import annotations

import weakref
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import numpy as np

from . import _lib as L
from .backend import PINNED, DppHandle
from .forms import dpp_form, dpp_splitted_form
from .mesh import Function, MixedFunctionSpace
from .parameters import B200_BACKEND, B200_BACKEND_KEY, DPPParameters
from .provider import bc_data, is_mixed, space_data


class ConvergenceError(RuntimeError):
    """The Krylov / Picard iteration stopped with a negative converged reason.  The reference's
    ``solver.solve()`` (solvers/solver.py:71, :121: Firedrake Linear/NonlinearVariationalSolver) raises
    ``firedrake.exceptions.ConvergenceError`` in that case; when Firedrake is importable this class derives
    from it, so ``except ConvergenceError`` written against the reference keeps working."""

    def __init__(self, message: str, reason: int = 0, iterations: int = 0, residual_norm: float = float("nan")):
        super().__init__(message)
        self.reason, self.iterations, self.residual_norm = int(reason), int(iterations), float(residual_norm)


try:  # pragma: no cover - Firedrake is not installable here
    from firedrake.exceptions import ConvergenceError as _FdConvergenceError

    class ConvergenceError(ConvergenceError, _FdConvergenceError):  # type: ignore[no-redef]
        pass
except Exception:
    pass

_REASONS = {-3: "DIVERGED_MAX_IT", -4: "DIVERGED_DTOL", -5: "DIVERGED_BREAKDOWN", -9: "DIVERGED_NANORINF",
            -10: "DIVERGED_INDEFINITE_MAT", -100: "DIVERGED_COMM_TIMEOUT"}


@dataclass(frozen=True)
class Solution:
    """solvers/solver.py:14-27 -- same three attributes."""

    solution: object
    iteration_number: int
    residual_error: float


_HANDLES: Dict[int, Tuple[weakref.ref, DppHandle]] = {}


def handle_for(W, device: Optional[int] = None) -> DppHandle:
    """One device-resident handle per function space (mesh topology is uploaded once, like
    Firedrake caches its maps on the mesh)."""
    key = id(W)
    hit = _HANDLES.get(key)
    if hit is not None and hit[0]() is W:
        return hit[1]
    sd = space_data(W)
    comm = getattr(W.mesh(), "comm", None)
    dev = device if device is not None else (comm.device if comm is not None else 0)
    if sd.node_coords is not None:
        # arbitrarily numbered (Firedrake/DMPlex) mesh: re-number to lexicographic if it is a tensor grid
        h = DppHandle.from_mesh_arrays(sd.dim, sd.degree, sd.cell_node_map, sd.node_coords, sd.coords,
                                       sd.coord_cell_node_map, n_nodes=sd.n_nodes, device=dev)
    else:
        h = DppHandle(sd.dim, sd.degree, sd.cell_node_map, sd.coords, sd.coord_cell_node_map, n_nodes=sd.n_nodes,
                      device=dev)
    if comm is not None and comm.size > 1:
        comm.attach(h, sd, W.sub(0) if is_mixed(W) else W)
    try:
        ref = weakref.ref(W, lambda _r, k=key: _drop(k))
    except TypeError:
        ref = (lambda w=W: w)
    _HANDLES[key] = (ref, h)
    return h


def _drop(key):
    hit = _HANDLES.pop(key, None)
    if hit is not None:
        hit[1].close()


def release_handles():
    for key in list(_HANDLES):
        _drop(key)


_KSP = {"cg": L.KSP_CG, "gmres": L.KSP_GMRES}
_PC = {"none": L.PC_NONE, "jacobi": L.PC_JACOBI, "pbjacobi": L.PC_PBJACOBI, "fieldsplit": L.PC_FIELDSPLIT}


def _block_options(opt: L.DppOptions, params: Dict):
    b0, b1 = params.get("fieldsplit_0"), params.get("fieldsplit_1")
    if b0 is None and b1 is None:
        return
    b0 = b0 if b0 is not None else b1
    b1 = b1 if b1 is not None else b0
    if b0 != b1:
        raise NotImplementedError("fieldsplit_0 and fieldsplit_1 must use the same block solver")
    ksp, pc = b0.get("ksp_type", "preonly"), b0.get("pc_type", "jacobi")
    if pc in ("lu", "ilu", "cholesky"):
        raise NotImplementedError(
            f"block pc_type={pc!r} (MUMPS LU / ILU, K8) is not built for the B200 path; use a Jacobi-CG block "
            "solver (B200_*_FIELDSPLIT_PARAMS) or run the preset on the reference path")
    if ksp not in ("cg", "preonly") or pc not in ("jacobi", "none"):
        raise NotImplementedError(f"unsupported block solver ksp_type={ksp!r} pc_type={pc!r}")
    opt.inner_ksp_type = L.INNER_CG if ksp == "cg" else L.INNER_PREONLY
    opt.inner_pc_type = L.PC_JACOBI if pc == "jacobi" else L.PC_NONE
    opt.inner_rtol = float(b0.get("ksp_rtol", 1e-5))   # PETSc default when unset
    opt.inner_atol = float(b0.get("ksp_atol", 1e-50))
    opt.inner_max_it = int(b0.get("ksp_max_it", 10000))


def options_from_petsc(handle: DppHandle, params: Dict, nonlinear: bool = False) -> L.DppOptions:
    """Translate PETSc-style option names (solvers/parameters.py) into dpp_options."""
    opt = handle.default_options()
    mat_type = params.get("mat_type", "matfree")
    if mat_type not in ("matfree", "aij"):
        raise NotImplementedError(f"mat_type={mat_type!r}")
    opt.operator_mode = L.OP_ASSEMBLED if mat_type == "aij" else L.OP_MATRIX_FREE
    if nonlinear or "snes_type" in params:
        st = params.get("snes_type", "picard_split")
        if st != "picard_split":
            # PICARD_*_SOLVER_PARAMS (solvers/parameters.py:71-95) ask for PETSc's pointwise SNES NGS /
            # nrichardson sweeps, whose iteration counts (92 / 5135 in the stored runs) a block method does not
            # reproduce: refuse instead of silently answering with 6 block-Picard iterations
            raise NotImplementedError(
                f"snes_type={st!r} is not built for the B200 path; the scale-splitting block iteration is selected "
                'with the explicit key "snes_type": "picard_split" (B200_PICARD_SPLIT_PARAMS)')
        opt.ksp_type = L.KSP_PICARD
        opt.rtol = float(params.get("snes_rtol", 1e-8))
        opt.atol = float(params.get("snes_atol", 1e-12))
        opt.max_it = int(params.get("snes_max_it", 50000))
        opt.inner_rtol, opt.inner_atol, opt.inner_max_it = 1e-10, 1e-50, 10000
        _block_options(opt, {k: v for k, v in params.items() if k.startswith("fieldsplit_")
                             and isinstance(v, dict) and v.get("pc_type") not in ("lu", "ilu")})
        return opt
    ksp = params.get("ksp_type", "gmres")
    if ksp == "preonly":
        raise NotImplementedError("ksp_type=preonly + pc_type=lu (MUMPS, K8) is not built for the B200 path")
    if ksp not in _KSP:
        raise NotImplementedError(f"ksp_type={ksp!r}")
    pc = params.get("pc_type", "none")
    if pc not in _PC:
        raise NotImplementedError(f"pc_type={pc!r} is not built for the B200 path (available: {sorted(_PC)})")
    opt.ksp_type, opt.pc_type = _KSP[ksp], _PC[pc]
    # unset keys: PETSc's defaults, except ksp_rtol, for which Firedrake's variational solvers inject 1e-7
    # (solving_utils DEFAULT_KSP_PARAMETERS) -- what the reference's solve_dpp therefore runs with
    opt.rtol = float(params.get("ksp_rtol", 1e-7))
    opt.atol = float(params.get("ksp_atol", 1e-50))
    opt.dtol = float(params.get("ksp_divtol", 1e4))
    opt.max_it = int(params.get("ksp_max_it", 10000))
    opt.gmres_restart = int(params.get("ksp_gmres_restart", 30))
    if ksp == "gmres" and not 1 <= opt.gmres_restart <= 30:
        raise NotImplementedError(f"ksp_gmres_restart={opt.gmres_restart}: the B200 GMRES keeps its Hessenberg / "
                                  "Givens state for restart lengths 1..30 (PETSc's default is 30)")
    if pc == "fieldsplit":
        fs = params.get("pc_fieldsplit_type", "multiplicative")
        if fs not in ("additive", "multiplicative"):
            raise NotImplementedError(f"pc_fieldsplit_type={fs!r}")
        opt.fieldsplit_type = L.FS_ADDITIVE if fs == "additive" else L.FS_MULTIPLICATIVE
        opt.inner_ksp_type, opt.inner_pc_type = L.INNER_PREONLY, L.PC_JACOBI
        _block_options(opt, params)
    if "b200_check_every" in params:
        opt.check_every = int(params["b200_check_every"])
    return opt


def _is_b200(params: Dict) -> bool:
    return params.get(B200_BACKEND_KEY) == B200_BACKEND


def _reference_solve(name: str, *args, **kw):
    try:
        from perphil.solvers import solver as ref  # type: ignore
    except Exception as exc:  # pragma: no cover - Firedrake is not installable here
        raise RuntimeError(
            f'solver_parameters without "{B200_BACKEND_KEY}": "{B200_BACKEND}" are routed to the reference '
            f"perphil/Firedrake path, which is not importable here ({exc}). The B200 path has no CPU fallback.")
    return getattr(ref, name)(*args, **kw)


def _new_function(W):
    if isinstance(W, MixedFunctionSpace):
        return Function(W)
    import firedrake as fd  # real Firedrake space

    return fd.Function(W)


def configure_handle(h: DppHandle, W, model_params, bcs) -> None:
    """Upload DPPParameters and Dirichlet data (DppHandle skips uploads that repeat what it already holds, so
    repeated solves of one problem keep the diagonal, the boundary classification and the CUDA graphs)."""
    h.set_params(float(model_params.k1), float(model_params.k2), float(model_params.beta), float(model_params.mu))
    got = {f: (n, v) for f, n, v in bc_data(W, bcs)}
    for f in (0, 1):
        n, v = got.get(f, (np.zeros(0, np.int32), np.zeros(0)))
        h.set_dirichlet(f, n, v)


def _raise_if_diverged(info, what: str):
    if info.converged_reason < 0:
        name = _REASONS.get(int(info.converged_reason), "DIVERGED")
        raise ConvergenceError(
            f"{what} failed to converge after {info.iterations} iterations with reason: {name} "
            f"({info.converged_reason}), residual norm {info.residual_norm:.6e}",
            info.converged_reason, info.iterations, info.residual_norm)


def _run(W, model_params, bcs, params, nonlinear, fields=None, monitor=None):
    h = handle_for(W)
    configure_handle(h, W, model_params, bcs)
    opt = options_from_petsc(h, params, nonlinear)
    hist_cap = int(params.get("b200_history", 0)) or (min(opt.max_it + 1, 1 << 16) if "ksp_monitor" in params else 0)
    n = h.n_nodes
    if fields is None and isinstance(W, MixedFunctionSpace):
        # the result Function is backed by a page-locked buffer: D2H lands in it directly
        sol = Function(W, buffer=PINNED.take(2 * n))
        _, info = h.solve(opt, want_solution=True, history=hist_cap, out=sol.vector)
    else:
        u, info = h.solve(opt, want_solution=True, history=hist_cap)
        sol = fields if fields is not None else _new_function(W)
        sol.sub(0).dat.data[:] = u[:n]
        sol.sub(1).dat.data[:] = u[n:]
    if "ksp_monitor" in params or "snes_monitor" in params:
        tag = "SNES Function norm" if nonlinear else "KSP Residual norm"
        for i, r in enumerate(info.history):
            print(f"  {i:3d} {tag} {r:.12e}")
    _LAST_INFO[0] = info
    if not params.get("b200_error_if_not_converged", True):
        return sol, info
    _raise_if_diverged(info, "Nonlinear solve" if nonlinear else "Linear solve")
    return sol, info


def solve_dpp(W, model_params: DPPParameters, bcs: List, solver_parameters: Dict = {},
              options_prefix: str = "dpp") -> Solution:
    """solvers/solver.py:30-76.  Raises ValueError unless W is a 2-field mixed space (:61-62)."""
    if not hasattr(W, "num_sub_spaces") or W.num_sub_spaces() != 2:
        raise ValueError(f"Expected a 2-field MixedFunctionSpace, got {type(W)}")
    if not _is_b200(solver_parameters):
        return _reference_solve("solve_dpp", W, model_params, bcs, solver_parameters, options_prefix)
    dpp_form(W, model_params)  # same structural check as the reference (:64)
    sol, info = _run(W, model_params, bcs, solver_parameters, nonlinear=False)
    result = Solution(sol, int(info.iterations), float(info.residual_norm))
    _LAST_INFO[0] = info
    return result


def solve_dpp_nonlinear(W, model_params: DPPParameters, bcs: List, solver_parameters: Dict = {},
                        options_prefix: str = "dpp_nonlinear") -> Solution:
    """solvers/solver.py:79-128 with the block Picard (scale-splitting) iteration in place of
    PETSc SNES NGS.  `solution` is the Function holding (p1, p2); iteration_number
    = outer iterations, residual_error = monolithic residual 2-norm (snes.getFunctionNorm())."""
    if not hasattr(W, "num_sub_spaces") or W.num_sub_spaces() != 2:
        raise ValueError(f"Expected a 2-field MixedFunctionSpace, got {type(W)}")
    if not _is_b200(solver_parameters):
        return _reference_solve("solve_dpp_nonlinear", W, model_params, bcs, solver_parameters, options_prefix)
    _, fields = dpp_splitted_form(W, model_params)
    sol, info = _run(W, model_params, bcs, solver_parameters, nonlinear=True, fields=fields)
    _LAST_INFO[0] = info
    return Solution(sol, int(info.iterations), float(info.residual_norm))


_LAST_INFO = [None]


def last_solve_info():
    """SolveInfo (timings, converged reason, history) of the most recent B200 solve."""
    return _LAST_INFO[0]
