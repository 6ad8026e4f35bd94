"""Assembled-matrix boundary (perphil.solvers.conditioning, solvers/conditioning.py:51-102):
`get_matrix_data_from_form(a, bcs)` -> CSR of the BC'd 2x2-block matrix, assembled on the GPU."""
from __future__ import annotations

from dataclasses import dataclass
from typing import List

import numpy as np
from scipy.sparse import csr_matrix

from .forms import DPPForm
from .provider import bc_data
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
    monolithic = hasattr(W, "num_sub_spaces") and W.num_sub_spaces() == 2
    if not monolithic:
        raise NotImplementedError("assemble per-scale blocks by slicing the monolithic CSR (iterative_bench.py:323-324)")
    prm = form.params
    h = handle_for(W)
    h.set_params(float(prm.k1), float(prm.k2), float(prm.beta), float(prm.mu))
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
