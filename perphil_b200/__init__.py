"""perphil_b200: the B200 (sm_100a) implementation of perphil's DPP assemble+solve hot path.

Public surface = the reference's own names for that path:
    solve_dpp, solve_dpp_nonlinear, Solution          (perphil.solvers.solver)
    dpp_form, dpp_delayed_form, dpp_splitted_form      (perphil.forms.dpp)
    DPPParameters                                      (perphil.models.dpp.parameters)
    get_matrix_data_from_form, calculate_condition_number   (perphil.solvers.conditioning)
    split_dpp_solution, calculate_darcy_velocity_from_pressure, l2_error, h1_seminorm_error   (perphil.utils.postprocessing)
    exact_expressions, exact_expressions_3d            (perphil.utils.manufactured_solutions)
plus the B200_* solver-parameter presets and Firedrake-shaped synthetic meshes/spaces.
"""
from .conditioning import (MatrixData, SpectrumEstimate, assemble_bilinear_form, calculate_condition_number,
                           condition_number_matrix_free, get_matrix_data_from_form)
from .forms import dpp_delayed_form, dpp_form, dpp_splitted_form
from .manufactured import exact_expressions, exact_expressions_3d, interpolate_exact
from .mesh import (Constant, DirichletBC, Function, FunctionSpace, MixedFunctionSpace, UnitCubeMesh, UnitSquareMesh,
                   create_function_spaces, create_mesh)
from .parameters import (B200_BACKEND, B200_BACKEND_KEY, B200_CG_FIELDSPLIT_PARAMS, B200_CG_JACOBI_AIJ_PARAMS,
                         B200_CG_JACOBI_PARAMS, B200_CG_PARAMS, B200_CG_PBJACOBI_PARAMS,
                         B200_GMRES_FIELDSPLIT_ADDITIVE_PARAMS, B200_GMRES_FIELDSPLIT_PARAMS, B200_GMRES_JACOBI_PARAMS,
                         B200_GMRES_PARAMS, B200_PICARD_SPLIT_PARAMS, DPPParameters)
from .postprocessing import (VectorFunction, calculate_darcy_velocity_from_pressure, h1_seminorm_error, l2_error,
                             slice_along_x, split_dpp_solution)
from .solver import ConvergenceError, Solution, handle_for, last_solve_info, release_handles, solve_dpp, solve_dpp_nonlinear

__all__ = [n for n in dir() if not n.startswith("_")]
