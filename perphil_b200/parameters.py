"""DPPParameters and solver-parameter presets.

`DPPParameters` mirrors perphil.models.dpp.parameters (models/dpp/parameters.py:5-53) without the
Firedrake dependency.  The B200_* presets are NEW presets in the style of
perphil.solvers.parameters (solvers/parameters.py:1-102): PETSc option names plus one marker key,
``"dpp_backend": "b200"``, that routes `solve_dpp` to libdppb200.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional

from .mesh import Constant

B200_BACKEND_KEY = "dpp_backend"
B200_BACKEND = "b200"
_MAX_ITERATION_NUMBER = 50000  # solvers/parameters.py:1


@dataclass
class DPPParameters:
    k1: float | Constant = 1.0
    k2: Optional[float | Constant] = None
    beta: float | Constant = 1.0
    mu: float | Constant = 1.0
    scale_contrast: float = 1e2

    def __post_init__(self):
        if not isinstance(self.k1, Constant):
            self.k1 = Constant(self.k1)
        if self.k2 is None:  # models/dpp/parameters.py:35-36
            self.k2 = Constant(float(self.k1) / self.scale_contrast)
        if not isinstance(self.k2, Constant):
            self.k2 = Constant(self.k2)
        if not isinstance(self.beta, Constant):
            self.beta = Constant(self.beta)
        if not isinstance(self.mu, Constant):
            self.mu = Constant(self.mu)

    @property
    def eta(self) -> float:  # models/dpp/parameters.py:44-53
        k1, k2, beta = float(self.k1), float(self.k2), float(self.beta)
        return math.sqrt(beta * (k1 + k2) / (k1 * k2))


_KSP_TOLS = {"ksp_rtol": 1.0e-8, "ksp_atol": 1.0e-12, "ksp_max_it": _MAX_ITERATION_NUMBER}

# Jacobi-preconditioned CG on the matrix-free operator: the headline preset (BASELINE.json)
B200_CG_JACOBI_PARAMS: dict = {
    B200_BACKEND_KEY: B200_BACKEND, "mat_type": "matfree", "ksp_type": "cg", "pc_type": "jacobi", **_KSP_TOLS,
}
B200_CG_PARAMS: dict = {**B200_CG_JACOBI_PARAMS, "pc_type": "none"}
B200_CG_PBJACOBI_PARAMS: dict = {**B200_CG_JACOBI_PARAMS, "pc_type": "pbjacobi"}
# same solver on the assembled CSR matrix (K1/K3a)
B200_CG_JACOBI_AIJ_PARAMS: dict = {**B200_CG_JACOBI_PARAMS, "mat_type": "aij"}

# GMRES(30) presets = PLAIN_GMRES_PARAMS / GMRES_JACOBI_PARAMS (solvers/parameters.py:21,24) on the B200
B200_GMRES_PARAMS: dict = {
    B200_BACKEND_KEY: B200_BACKEND, "mat_type": "matfree", "ksp_type": "gmres", "pc_type": "none", **_KSP_TOLS,
}
B200_GMRES_JACOBI_PARAMS: dict = {**B200_GMRES_PARAMS, "pc_type": "jacobi"}

_BLOCK_CG_JACOBI = {"ksp_type": "cg", "pc_type": "jacobi", "ksp_rtol": 1.0e-10, "ksp_atol": 1.0e-50,
                    "ksp_max_it": 10000}
_BLOCK_JACOBI = {"ksp_type": "preonly", "pc_type": "jacobi"}

# fieldsplit presets: same option layout as FIELDSPLIT_LU_PARAMS (solvers/parameters.py:30-37) with
# iterative (Jacobi-CG) block solves instead of MUMPS LU (K8 is not built)
B200_GMRES_FIELDSPLIT_PARAMS: dict = {
    **B200_GMRES_PARAMS,
    "pc_type": "fieldsplit", "pc_fieldsplit_type": "multiplicative",
    "pc_fieldsplit_0_fields": "0", "pc_fieldsplit_1_fields": "1",
    "fieldsplit_0": _BLOCK_CG_JACOBI, "fieldsplit_1": _BLOCK_CG_JACOBI,
}
# block-Jacobi (additive) fieldsplit under CG: BASELINE.json configs[1]
B200_CG_FIELDSPLIT_PARAMS: dict = {
    **B200_CG_JACOBI_PARAMS,
    "pc_type": "fieldsplit", "pc_fieldsplit_type": "additive",
    "pc_fieldsplit_0_fields": "0", "pc_fieldsplit_1_fields": "1",
    "fieldsplit_0": _BLOCK_CG_JACOBI, "fieldsplit_1": _BLOCK_CG_JACOBI,
}
B200_GMRES_FIELDSPLIT_ADDITIVE_PARAMS: dict = {**B200_GMRES_FIELDSPLIT_PARAMS, "pc_fieldsplit_type": "additive"}

# scale-splitting block Picard on the dpp_delayed_form split (forms/dpp.py:135-205); tolerances as
# PICARD_*_SOLVER_PARAMS (solvers/parameters.py:71-95)
B200_PICARD_SPLIT_PARAMS: dict = {
    B200_BACKEND_KEY: B200_BACKEND, "mat_type": "matfree", "snes_type": "picard_split",
    "snes_rtol": 1e-8, "snes_atol": 1e-12, "snes_max_it": _MAX_ITERATION_NUMBER,
    "fieldsplit_0": _BLOCK_CG_JACOBI, "fieldsplit_1": _BLOCK_CG_JACOBI,
}
