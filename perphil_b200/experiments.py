"""Approach routing and the performance harness of perphil.experiments, for the B200 path
(experiments/iterative_bench.py:31-48,110-131,157-252 and experiments/petsc_profiling_3d.py:43-230;
SURVEY 8(f) item 2).  Rows carry the column names of the reference's stored CSVs
(`approach,nx,ny,dofs,num_cells,iterations,residual,time_total,time_KSPSolve,time_MatMult,...`) so that
B200 rows can be concatenated with `notebooks/results-conforming-3d/petsc_profiling/*.csv`; times come
from CUDA events inside libdppb200 instead of PETSc's event log (`backend = "cuda-events"`).
"""
from __future__ import annotations

import time
from dataclasses import dataclass
from enum import Enum
from typing import Any, Dict, Iterable, List, Optional, Tuple

from . import parameters as P
from .manufactured import exact_expressions_3d
from .mesh import Constant, DirichletBC, UnitCubeMesh, create_function_spaces
from .parameters import DPPParameters
from .solver import handle_for, last_solve_info, solve_dpp, solve_dpp_nonlinear


class Approach(str, Enum):
    """Solver approaches (values = the `approach` strings of the reference's CSVs where one exists)."""

    PLAIN_GMRES = "GMRES"                                   # iterative_bench.py:42
    GMRES_JACOBI = "GMRES + Jacobi PC"
    SS_GMRES = "Scale-Splitting GMRES"                      # :44 (multiplicative fieldsplit; blocks: Jacobi-CG, not LU)
    SS_GMRES_ADDITIVE = "Scale-Splitting GMRES (additive)"
    PICARD = "Scaling-Splitting Picard"                     # :46 with block (not pointwise NGS) sweeps
    CG_JACOBI = "CG + Jacobi PC"
    CG_FIELDSPLIT = "CG + block-Jacobi fieldsplit"
    # reference approaches that need MUMPS LU / ILU(0) (K8: not built) stay on the reference path
    GMRES_ILU = "GMRES + ILU PC"
    SS_GMRES_ILU = "Scale-Splitting GMRES + ILU PC"
    MONOLITHIC_MUMPS = "Monolithic LU with MUMPS"


_PRESETS = {
    Approach.PLAIN_GMRES: P.B200_GMRES_PARAMS,
    Approach.GMRES_JACOBI: P.B200_GMRES_JACOBI_PARAMS,
    Approach.SS_GMRES: P.B200_GMRES_FIELDSPLIT_PARAMS,
    Approach.SS_GMRES_ADDITIVE: P.B200_GMRES_FIELDSPLIT_ADDITIVE_PARAMS,
    Approach.PICARD: P.B200_PICARD_SPLIT_PARAMS,
    Approach.CG_JACOBI: P.B200_CG_JACOBI_PARAMS,
    Approach.CG_FIELDSPLIT: P.B200_CG_FIELDSPLIT_PARAMS,
}


def params_for(approach: Approach) -> Dict:
    """iterative_bench.py:157-188."""
    approach = Approach(approach)
    if approach not in _PRESETS:
        raise NotImplementedError(f"{approach.value!r} needs MUMPS LU / ILU(0), which the B200 path does not build; "
                                  "run it with perphil.experiments.iterative_bench on the reference path")
    return dict(_PRESETS[approach])


def default_bcs(W) -> List[DirichletBC]:
    """Homogeneous Dirichlet data on both pressures (iterative_bench.py:110-121)."""
    return [DirichletBC(W.sub(0), Constant(0.0), "on_boundary"), DirichletBC(W.sub(1), Constant(0.0), "on_boundary")]


def default_model_params() -> DPPParameters:
    """k1 = beta = mu = 1, k2 = 1e-2 (iterative_bench.py:124-131)."""
    return DPPParameters(k1=1.0, k2=1.0 / 1e2, beta=1.0, mu=1.0)


@dataclass(frozen=True)
class SolveResult:
    """iterative_bench.py:51-76."""

    approach: Approach
    nx: int
    ny: int
    iteration_number: int
    residual_error: float
    fields: Optional[Tuple[Any, Any]] = None


def solve_on_mesh(W, approach: Approach, params: Optional[DPPParameters] = None, bcs: Optional[List] = None) -> SolveResult:
    """iterative_bench.py:191-252."""
    approach = Approach(approach)
    params = params or default_model_params()
    bcs = bcs or default_bcs(W)
    sp = params_for(approach)
    solve = solve_dpp_nonlinear if approach == Approach.PICARD else solve_dpp
    sol = solve(W, params, bcs=bcs, solver_parameters=sp)
    fields = tuple(sol.solution.split()) if hasattr(sol.solution, "split") else None
    return SolveResult(approach, -1, -1, int(sol.iteration_number), float(sol.residual_error), fields)


CSV_COLUMNS = ("approach", "nx", "ny", "dofs", "num_cells", "iterations", "residual", "time_total", "time_total_repeats",
               "time_PCSetUp", "time_PCApply", "time_KSPSolve", "time_SNESSolve", "time_MatMult", "backend", "repeats")


def run_perf_once_3d(nx: int, approach: Approach, repeats: int = 3, degree: int = 1, comm=None) -> Dict[str, Any]:
    """One row of the 3-D timing sweep (petsc_profiling_3d.py:43-200) on hexahedra: one warm-up solve, then
    `repeats` timed solves of the manufactured-BC problem; `time_total` = mean wall seconds per solve,
    event columns = device seconds SUMMED over the repeats like PETSc's log."""
    approach = Approach(approach)
    mesh = UnitCubeMesh(nx, nx, nx, comm=comm)
    _, V = create_function_spaces(mesh, pressure_deg=degree)
    W = V * V
    params = default_model_params()
    _, p1, _, p2 = exact_expressions_3d(mesh, params)
    bcs = [DirichletBC(W.sub(0), p1, "on_boundary"), DirichletBC(W.sub(1), p2, "on_boundary")]
    sp = params_for(approach)
    solve = solve_dpp_nonlinear if approach == Approach.PICARD else solve_dpp
    solve(W, params, bcs=bcs, solver_parameters=sp)  # warm-up (petsc_profiling.py:697-699)
    h = handle_for(W)
    apply_s = h.time_apply(reps=10, warmup=2) * 1e-3
    wall = ksp = setup = matmult = 0.0
    sol = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        sol = solve(W, params, bcs=bcs, solver_parameters=sp)
        wall += time.perf_counter() - t0
        info = last_solve_info()
        ksp += info.solve_ms * 1e-3
        setup += info.setup_ms * 1e-3
        matmult += info.apply_count * apply_s   # operator applications x the stand-alone apply time
    row = {
        "approach": approach.value, "nx": nx, "ny": nx, "dofs": 2 * (degree * nx + 1) ** 3, "num_cells": nx ** 3,
        "iterations": int(sol.iteration_number), "residual": float(sol.residual_error),
        "time_total": wall / repeats, "time_total_repeats": wall, "time_PCSetUp": setup,
        "time_PCApply": max(ksp - matmult, 0.0) if approach in (Approach.SS_GMRES, Approach.SS_GMRES_ADDITIVE,
                                                                 Approach.CG_FIELDSPLIT) else 0.0,
        "time_KSPSolve": ksp, "time_SNESSolve": ksp + setup, "time_MatMult": matmult,
        "backend": "cuda-events", "repeats": repeats,
    }
    return row


def run_perf_sweep_3d(nx_list: Iterable[int], approaches: Iterable[Approach], repeats: int = 3, degree: int = 1):
    """pandas DataFrame with the reference's column names (petsc_profiling_3d.py:203-230)."""
    import pandas as pd

    rows = [run_perf_once_3d(nx, a, repeats=repeats, degree=degree) for nx in nx_list for a in approaches]
    return pd.DataFrame(rows, columns=list(CSV_COLUMNS))
