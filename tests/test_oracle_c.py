"""Pins the C/OpenMP restatement (oracle/dpp_oracle_c.c, the timed CPU baseline) to the Python
oracle (oracle/dpp_oracle.py), which tests/test_oracle_golden.py pins to the reference's stored
numbers.  CPU only."""
import numpy as np
import pytest
import scipy.sparse as sp

from oracle import c_oracle as co
from oracle import dpp_oracle as orc


def _pair(cells, degree):
    osys = orc.build_system(orc.structured_mesh(cells, degree), orc.Params(k1=1.0, k2=1e-2, beta=1.0, mu=1.0),
                            "manufactured")
    csys = co.manufactured_system(cells, degree)
    return osys, csys


@pytest.mark.parametrize("cells,degree", [((8, 8, 8), 1), ((5, 7, 9), 1), ((16, 16), 1), ((10, 10), 1),
                                          ((3, 4, 5), 2), ((6, 5), 2)])
def test_c_matrix_and_rhs_match_python_oracle(cells, degree):
    osys, csys = _pair(cells, degree)
    indptr, indices, data, b, u0 = csys.export()
    n = csys.n_dof
    assert n == osys.n_dof
    A = sp.csr_matrix((data, indices, indptr), shape=(n, n))
    # full element pattern: nnz = 4 * prod(3N+1) for Q1
    if degree == 1:
        assert csys.nnz == 4 * int(np.prod([3 * c + 1 for c in cells]))
    for r in range(0, n, max(1, n // 50)):
        assert np.all(np.diff(indices[indptr[r]:indptr[r + 1]]) > 0)
    A.eliminate_zeros()  # conditioning.py:86
    ref = osys.A_bc
    assert np.array_equal(A.indptr, ref.indptr)
    assert np.array_equal(A.indices, ref.indices)
    assert np.max(np.abs(A.data - ref.data)) <= 1e-13 * np.max(np.abs(ref.data))
    ou0, ob = orc.lifted_rhs(osys)
    assert np.array_equal(u0, ou0)
    assert np.linalg.norm(b - ob) <= 1e-13 * np.linalg.norm(ob)
    x = np.random.default_rng(0).standard_normal(n)
    assert np.linalg.norm(csys.spmv(x) - ref @ x) <= 1e-13 * np.linalg.norm(ref @ x)


def test_c_initial_residual_matches_notebook():
    """'0 SNES Function norm 8.485690809593e+04' (operator-splitting notebook, 10x10 quads)."""
    csys = co.manufactured_system((10, 10), 1)
    b = csys.export()[3]
    assert abs(np.linalg.norm(b) - 8.485690809593e04) < 1e-7


@pytest.mark.parametrize("cells", [(8, 8, 8), (16, 16, 16), (16, 16), (12, 9, 7)])
@pytest.mark.parametrize("pc", ["jacobi", "none"])
def test_c_cg_matches_python_oracle(cells, pc):
    osys, csys = _pair(cells, 1)
    ref = orc.solve_dpp_oracle(osys, "cg", pc)
    got = csys.cg(pc, history=ref.iteration_number + 1)
    assert got.reason == ref.reason
    if pc == "jacobi":
        assert got.iteration_number == ref.iteration_number
        assert np.allclose(got.history, ref.history, rtol=1e-9)
    else:
        # unpreconditioned CG on this system (p2 ~ 1e6 boundary data) loses orthogonality: summation-order
        # rounding differences grow along the recurrence, so only the first third is compared tightly
        third = max(2, len(ref.history) // 3)
        assert np.allclose(got.history[:third], ref.history[:third], rtol=1e-8)
        assert abs(got.iteration_number - ref.iteration_number) <= max(2, ref.iteration_number // 30)
    assert np.linalg.norm(got.u - ref.u) <= 1e-8 * np.linalg.norm(ref.u)


def test_c_cg_thread_count_does_not_change_iterations():
    csys = co.manufactured_system((12, 12, 12), 1)
    n0 = co.num_threads()
    co.set_num_threads(1)
    a = csys.cg("jacobi")
    co.set_num_threads(max(2, n0))
    b = csys.cg("jacobi")
    co.set_num_threads(n0)
    assert a.iteration_number == b.iteration_number
    assert np.linalg.norm(a.u - b.u) <= 1e-10 * np.linalg.norm(a.u)
