"""Zero-patch hook (style 2): a petsc4py "python" KSP that lets the UNMODIFIED
perphil.solvers.solver.solve_dpp (solvers/solver.py:30-76) run its Krylov solve on libdppb200.

    params = {"mat_type": "matfree", "ksp_type": "python",
              "ksp_python_type": "perphil_b200.petsc_plugin.B200DPPKSP", "pc_type": "none",
              "b200": perphil_b200.B200_CG_JACOBI_PARAMS}      # inner preset, optional
    perphil_b200.petsc_plugin.register(model_params)            # DPPParameters are not recoverable from UFL
    sol = perphil.solvers.solver.solve_dpp(W, model_params, bcs, solver_parameters=params)

Firedrake builds the matfree Mat (python context: the form `a` and the DirichletBCs) and the lifted
right-hand side; PETSc calls `B200DPPKSP.solve(ksp, b, x)`.  solve_dpp reads back
`ksp.getIterationNumber()` / `ksp.getResidualNorm()` (solver.py:73-74), so both are set here.
petsc4py / Firedrake are not installable in this build environment: this module is import-safe without
them and is exercised only by its argument checks (tests/test_host_layer.py); see INTEGRATION.md.
"""
from __future__ import annotations

import numpy as np

from .parameters import B200_CG_JACOBI_PARAMS

_REGISTERED = {"params": None, "preset": None}


def register(model_params, preset=None):
    """Tell the plug-in which DPPParameters (and which B200_* preset) the next solves use."""
    _REGISTERED["params"], _REGISTERED["preset"] = model_params, preset


class B200DPPKSP:
    """petsc4py python-KSP context (create/solve protocol of PETSc.KSP.Type.PYTHON)."""

    def create(self, ksp):
        self._work = None

    def solve(self, ksp, b, x):
        from .solver import _run  # late: keeps module import light

        if _REGISTERED["params"] is None:
            raise RuntimeError("perphil_b200.petsc_plugin.register(model_params) must be called before the solve")
        A, _ = ksp.getOperators()
        ctx = A.getPythonContext()                      # Firedrake ImplicitMatrixContext
        W = ctx.a.arguments()[0].function_space()       # the MixedFunctionSpace of dpp_form (forms/dpp.py:116-117)
        bcs = list(getattr(ctx, "row_bcs", None) or getattr(ctx, "bcs", ()))
        preset = dict(_REGISTERED["preset"] or B200_CG_JACOBI_PARAMS)
        # Firedrake hands PETSc the lifted system A_bc d = b with d = 0 on constrained rows: solve for the
        # full field with the BC values and return the increment (u - u0), which is what SNES ksponly adds.
        sol, info = _run(W, _REGISTERED["params"], bcs, preset, nonlinear=False)
        n = W.sub(0).dim()
        u = np.concatenate([np.asarray(sol.sub(0).dat.data_ro), np.asarray(sol.sub(1).dat.data_ro)])
        u0 = np.zeros_like(u)
        for bc in bcs:
            f = bc.function_space().index
            g = bc.function_arg
            nodes = np.asarray(bc.nodes)
            u0[f * n + nodes] = np.asarray(g.dat.data_ro)[nodes] if hasattr(g, "dat") else float(g)
        x.setArray(u - u0)
        ksp.setIterationNumber(int(info.iterations))
        ksp.setResidualNorm(float(info.residual_norm))
        ksp.setConvergedReason(int(info.converged_reason))
