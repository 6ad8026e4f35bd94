"""Assembled-matrix boundary (perphil.solvers.conditioning, solvers/conditioning.py:51-102):
`get_matrix_data_from_form(a, bcs)` -> CSR of the BC'd 2x2-block matrix, assembled on the GPU; and the
condition numbers of solvers/conditioning.py:105-218 -- `calculate_condition_number` on such a CSR matrix with
the reference's own scipy semantics (host), plus `condition_number_matrix_free`, which gets the extreme singular
values from a Lanczos run on the GPU's matrix-free operator (`dpp_lanczos`) and so is not
capped at the sizes a dense SVD can take."""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional

import numpy as np
from scipy.sparse import csr_matrix

from .forms import DPPForm
from .provider import bc_data, is_mixed
from .solver import handle_for


@dataclass
class MatrixData:
    """solvers/conditioning.py:20-48 (same field names; petsc_matrix is the raw CSR triplet here)."""

    assembled_matrix: object
    petsc_matrix: object
    is_symmetric: bool
    sparse_csr_data: csr_matrix
    number_of_nonzero_entries: int
    number_of_dofs: int
    symmetry_tolerance: float


def assemble_bilinear_form(form: DPPForm, boundary_conditions: List):
    """fd.assemble(form, bcs=..., mat_type="aij") (conditioning.py:51-63): (indptr, indices, data)
    with the full element pattern and sorted column indices."""
    W = form.space
    if form.rank != 2:
        raise ValueError("assemble_bilinear_form expects a rank-2 form")
    prm = form.params
    h = handle_for(W)
    h.set_params(float(prm.k1), float(prm.k2), float(prm.beta), float(prm.mu))
    if not is_mixed(W):
        # per-scale form of dpp_delayed_form on a scalar space V (forms/dpp.py:135-205; assembled with its own
        # scalar BC at notebooks/conforming-galerkin-fem-operator-splitting-2D-perphil.py:463-480): the diagonal
        # block (f, f) of the V x V matrix, assembled and extracted on the GPU
        if len(form.blocks) != 1 or form.blocks[0][0] != form.blocks[0][1]:
            raise ValueError("a form on a scalar space must denote one diagonal block (dpp_delayed_form)")
        f = form.blocks[0][0]
        got = {fl: (n, v) for fl, n, v in bc_data(W, boundary_conditions, scalar_field=f)}
        n, v = got.get(f, (np.zeros(0, np.int32), np.zeros(0)))
        h.set_dirichlet(f, n, v)
        h.set_dirichlet(1 - f, np.zeros(0, np.int32), np.zeros(0))
        return h.assemble_csr_block(f, f)
    got = {f: (n, v) for f, n, v in bc_data(W, boundary_conditions)}
    for f in (0, 1):
        n, v = got.get(f, (np.zeros(0, np.int32), np.zeros(0)))
        h.set_dirichlet(f, n, v)
    return h.assemble_csr()


def get_matrix_data_from_form(form: DPPForm, boundary_conditions: List, symmetry_tolerance: float = 1e-8) -> MatrixData:
    """conditioning.py:66-102: CSR from getValuesCSR(), then eliminate_zeros() (:86)."""
    indptr, indices, data = assemble_bilinear_form(form, boundary_conditions)
    ndofs = indptr.size - 1
    csr = csr_matrix((data, indices, indptr), shape=(ndofs, ndofs))
    csr.eliminate_zeros()
    asym = abs(csr - csr.T)
    is_symmetric = bool(asym.nnz == 0 or asym.max() <= symmetry_tolerance)
    return MatrixData((indptr, indices, data), (indptr, indices, data), is_symmetric, csr, int(csr.nnz), int(ndofs),
                      symmetry_tolerance)


DEFAULT_CONDITION_NUMBER_TOLERANCE = 1e-7   # solvers/conditioning.py:9


def calculate_condition_number(scipy_csr_sparse_matrix: csr_matrix, num_singular_values: Optional[int] = None,
                               use_sparse: bool = False, zero_tol: float = DEFAULT_CONDITION_NUMBER_TOLERANCE) -> float:
    """solvers/conditioning.py:105-218 on a host CSR matrix: ratio of the largest to the smallest singular value
    above `zero_tol`; dense SVD unless `use_sparse` with a small `num_singular_values` (then ARPACK svds for the
    two ends).  Host-side by definition (the reference calls scipy here); the GPU route for large operators is
    `condition_number_matrix_free`."""
    from scipy.linalg import svd
    from scipy.sparse.linalg import svds

    nmin = min(scipy_csr_sparse_matrix.shape)
    if nmin == 0:
        return float("nan")
    dense = (not use_sparse) or num_singular_values is None or num_singular_values <= 0 or int(num_singular_values) >= nmin - 1
    if dense:
        s = np.asarray(svd(scipy_csr_sparse_matrix.toarray(), compute_uv=False, check_finite=False))
        s = s[s > zero_tol]
        return float("inf") if s.size == 0 else float(s.max() / s.min())
    smax = float(np.max(svds(scipy_csr_sparse_matrix, k=1, which="LM", maxiter=10000, return_singular_vectors=False,
                             solver="arpack")))
    smin = float(np.min(svds(scipy_csr_sparse_matrix, k=1, which="SM", maxiter=20000, return_singular_vectors=False,
                             solver="arpack", tol=1e-8)))
    return float("inf") if smin <= zero_tol else smax / smin


@dataclass
class SpectrumEstimate:
    condition_number: float
    sigma_max: float
    sigma_min: float
    lanczos_steps: int
    converged: bool


def condition_number_matrix_free(form: DPPForm, boundary_conditions: List, block: Optional[int] = None,
                                 rtol: float = 1e-8, max_steps: int = 20000, first_steps: int = 64, seed: int = 0,
                                 zero_tol: float = DEFAULT_CONDITION_NUMBER_TOLERANCE) -> SpectrumEstimate:
    """kappa_2 of the BC'd DPP matrix (`block=None`) or of its diagonal block A00 / A11 (`block=0|1`, the slices
    iterative_bench.py:323-324 takes) without assembling it.  The matrix is symmetric, so its singular values are
    the moduli of its eigenvalues; the extreme Ritz values of a Lanczos run on the GPU operator converge to them.
    The run is repeated with twice the steps until kappa changes by less than `rtol` (relative)."""
    from scipy.linalg import eigvalsh_tridiagonal

    W = form.space
    if form.rank != 2 or not (hasattr(W, "num_sub_spaces") and W.num_sub_spaces() == 2):
        raise ValueError("condition_number_matrix_free expects the monolithic rank-2 DPP form")
    prm = form.params
    h = handle_for(W)
    h.set_params(float(prm.k1), float(prm.k2), float(prm.beta), float(prm.mu))
    got = {f: (n, v) for f, n, v in bc_data(W, boundary_conditions)}
    for f in (0, 1):
        n, v = got.get(f, (np.zeros(0, np.int32), np.zeros(0)))
        h.set_dirichlet(f, n, v)
    which = 0 if block is None else 1 + int(block)
    ndof = (2 if block is None else 1) * h.n_nodes
    steps, prev, est = min(first_steps, ndof), None, None
    while True:
        a, b = h.lanczos(steps, which=which, seed=seed)
        m = a.size
        theta = np.abs(eigvalsh_tridiagonal(a, b[: m - 1])) if m > 1 else np.abs(a)
        theta = theta[theta > zero_tol]
        if theta.size == 0:
            return SpectrumEstimate(float("inf"), 0.0, 0.0, m, False)
        kappa = float(theta.max() / theta.min())
        exhausted = m < steps or steps >= ndof
        done = exhausted or (prev is not None and abs(kappa - prev) <= rtol * kappa)
        est = SpectrumEstimate(kappa, float(theta.max()), float(theta.min()), m, bool(done))
        if done or steps >= max_steps:
            return est
        prev = kappa
        steps = min(2 * steps, max_steps, ndof)
