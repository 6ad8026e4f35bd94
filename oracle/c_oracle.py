"""ctypes binding of oracle/libdpporacle.so (dpp_oracle_c.c) -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

The C/OpenMP restatement of the reference's CPU path (assembled AIJ matrix, SeqAIJ MatMult,
KSPCG + PCJACOBI) on uniform Q1/Q2 tensor grids.  Used (i) by tests/test_oracle_c.py, which pins it
to oracle/dpp_oracle.py and through it to the reference's stored numbers, and (ii) by bench.py as
the timed CPU baseline (``cpu_baseline`` and ``--impl reference``) with all host threads.
perphil_b200 never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def load(build: bool = True):
    global _LIB
    if _LIB is not None:
        return _LIB
    path = os.path.join(_HERE, "libdpporacle.so")
    if not os.path.exists(path) and build:
        subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
    lib = C.CDLL(path)
    lib.orc_build.restype = C.c_void_p
    lib.orc_build.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_int), C.c_double, C.c_double, C.c_double, C.c_double,
                              C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
    lib.orc_destroy.argtypes = [C.c_void_p]
    lib.orc_n_dof.restype = C.c_int64
    lib.orc_n_dof.argtypes = [C.c_void_p]
    lib.orc_nnz.restype = C.c_int64
    lib.orc_nnz.argtypes = [C.c_void_p]
    lib.orc_export.argtypes = [C.c_void_p] * 6
    lib.orc_spmv.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    lib.orc_cg.restype = C.c_int
    lib.orc_cg.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_double, C.c_int, C.c_void_p,
                           C.POINTER(C.c_double), C.POINTER(C.c_int), C.c_void_p, C.c_int, C.POINTER(C.c_double)]
    lib.orc_gmres.restype = C.c_int
    lib.orc_gmres.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int,
                              C.c_double, C.c_double, C.c_int, C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int),
                              C.c_void_p, C.c_int, C.POINTER(C.c_int64)]
    lib.orc_get_b.argtypes = [C.c_void_p, C.c_void_p]
    lib.orc_num_threads.restype = C.c_int
    lib.orc_set_num_threads.argtypes = [C.c_int]
    _LIB = lib
    return lib


@dataclass
class CgResult:
    u: Optional[np.ndarray]
    iteration_number: int
    residual_error: float
    reason: int
    history: List[float]
    spmv_seconds: float


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class COracleSystem:
    """Assembled, Dirichlet-eliminated DPP system on a uniform unit square / cube tensor grid."""

    def __init__(self, cells: Sequence[int], degree: int, k1: float, k2: float, beta: float, mu: float,
                 bc_nodes0, bc_vals0, bc_nodes1, bc_vals1):
        lib = load()
        cells = tuple(int(c) for c in cells)
        arr = (C.c_int * len(cells))(*cells)
        n0 = np.ascontiguousarray(bc_nodes0, dtype=np.int32)
        v0 = np.ascontiguousarray(bc_vals0, dtype=np.float64)
        n1 = np.ascontiguousarray(bc_nodes1, dtype=np.int32)
        v1 = np.ascontiguousarray(bc_vals1, dtype=np.float64)
        self._lib = lib
        self._h = lib.orc_build(len(cells), degree, arr, k1, k2, beta, mu, n0.size, _ptr(n0), _ptr(v0), n1.size,
                                _ptr(n1), _ptr(v1))
        self.n_dof = int(lib.orc_n_dof(self._h))
        self.nnz = int(lib.orc_nnz(self._h))

    def close(self):
        if self._h:
            self._lib.orc_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def export(self):
        """(indptr, indices, data, b, u0): full element pattern with explicit zeros on eliminated entries."""
        indptr = np.empty(self.n_dof + 1, np.int64)
        indices = np.empty(self.nnz, np.int32)
        data = np.empty(self.nnz)
        b = np.empty(self.n_dof)
        u0 = np.empty(self.n_dof)
        self._lib.orc_export(self._h, _ptr(indptr), _ptr(indices), _ptr(data), _ptr(b), _ptr(u0))
        return indptr, indices, data, b, u0

    def spmv(self, x: np.ndarray) -> np.ndarray:
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.empty_like(x)
        self._lib.orc_spmv(self._h, _ptr(x), _ptr(y))
        return y

    def cg(self, pc: str = "jacobi", rtol=1e-8, atol=1e-12, dtol=1e4, max_it=50000, want_solution=True,
           history=0) -> CgResult:
        u = np.empty(self.n_dof) if want_solution else None
        hist = np.zeros(max(history, 1))
        rn, reason, tmv = C.c_double(), C.c_int(), C.c_double()
        its = self._lib.orc_cg(self._h, 1 if pc == "jacobi" else 0, rtol, atol, dtol, max_it, _ptr(u), C.byref(rn),
                               C.byref(reason), _ptr(hist) if history else None, history, C.byref(tmv))
        return CgResult(u, int(its), float(rn.value), int(reason.value), list(hist[:min(history, its + 1)]),
                        float(tmv.value))

    def rhs(self) -> np.ndarray:
        """Lifted right-hand side b = -(A u0) on unconstrained rows (solver.py:66-71)."""
        b = np.empty(self.n_dof)
        self._lib.orc_get_b(self._h, _ptr(b))
        return b

    def gmres(self, pc: str = "none", restart=30, rtol=1e-8, atol=1e-12, dtol=1e4, max_it=50000, inner=None,
              want_solution=True, history=0) -> "GmresResult":
        """KSPGMRES(restart), left PC: 'none' | 'jacobi' | 'fieldsplit' (multiplicative) | 'fieldsplit_additive';
        ``inner`` = dict(ksp_type='cg'|'preonly', ksp_rtol, ksp_atol, ksp_max_it) for the Jacobi-CG block solves."""
        code = {"none": 0, "jacobi": 1, "fieldsplit": 2, "fieldsplit_additive": 3}[pc]
        inner = inner or {}
        u = np.empty(self.n_dof) if want_solution else None
        hist = np.zeros(max(history, 1))
        rn, reason, inner_its = C.c_double(), C.c_int(), C.c_int64()
        its = self._lib.orc_gmres(self._h, code, restart, rtol, atol, dtol, max_it,
                                  0 if inner.get("ksp_type", "cg") == "preonly" else 1,
                                  inner.get("ksp_rtol", 1e-5), inner.get("ksp_atol", 1e-50),
                                  inner.get("ksp_max_it", 10000), _ptr(u), C.byref(rn), C.byref(reason),
                                  _ptr(hist) if history else None, history, C.byref(inner_its))
        h = hist[:history] if history else hist[:0]
        nz = np.flatnonzero(h)   # entries written (every recorded norm is > 0 except an exact-zero final one)
        return GmresResult(u, int(its), float(rn.value), int(reason.value), list(h[:nz[-1] + 1] if nz.size else h[:0]),
                           int(inner_its.value))


@dataclass
class GmresResult:
    u: Optional[np.ndarray]
    iteration_number: int
    residual_error: float
    reason: int
    history: List[float]
    inner_iterations: int


def num_threads() -> int:
    return int(load().orc_num_threads())


def set_num_threads(n: int):
    load().orc_set_num_threads(int(n))


def _boundary_nodes(cells: Sequence[int], degree: int):
    cells = tuple(int(c) for c in cells)
    counts = [degree * c + 1 for c in cells]
    on_b = np.zeros(counts, dtype=bool)
    for d, c in enumerate(counts):
        sl = [slice(None)] * len(counts)
        for edge in (0, c - 1):
            sl[d] = edge
            on_b[tuple(sl)] = True
    return np.flatnonzero(on_b.ravel()).astype(np.int32)


def constant_bc_system(cells: Sequence[int], degree: int = 1, k1=1.0, k2=1e-6, beta=1e2, mu=1.0, p1=1.0,
                       p2=0.0) -> COracleSystem:
    """BASELINE config 5 data: constants p1, p2 on the whole boundary (petsc_profiling.py:685-690)."""
    nb = _boundary_nodes(cells, degree)
    return COracleSystem(cells, degree, k1, k2, beta, mu, nb, np.full(nb.size, float(p1)), nb,
                         np.full(nb.size, float(p2)))


def manufactured_system(cells: Sequence[int], degree: int = 1, k1=1.0, k2=1e-2, beta=1.0, mu=1.0) -> COracleSystem:
    """Uniform grid + exact_expressions(_3d) Dirichlet data on the whole boundary (both fields)."""
    from . import dpp_oracle as orc

    cells = tuple(int(c) for c in cells)
    counts = [degree * c + 1 for c in cells]
    on_b = np.zeros(counts, dtype=bool)
    for d, c in enumerate(counts):  # same boundary set as dpp_oracle.structured_mesh, without the full mesh
        sl = [slice(None)] * len(counts)
        for edge in (0, c - 1):
            sl[d] = edge
            on_b[tuple(sl)] = True
    nb = np.flatnonzero(on_b.ravel()).astype(np.int32)
    ijk = np.unravel_index(nb, counts)
    coords = np.stack([np.linspace(0.0, 1.0, c)[i] for i, c in zip(ijk, counts)], axis=1)
    prm = orc.Params(k1=k1, k2=k2, beta=beta, mu=mu)
    p1, p2 = orc.exact_pressures(coords, prm)
    return COracleSystem(cells, degree, k1, k2, beta, mu, nb, p1, nb, p2)
