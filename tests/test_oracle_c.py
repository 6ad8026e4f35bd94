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


INNER = {"ksp_type": "cg", "pc_type": "jacobi", "ksp_rtol": 1e-10, "ksp_atol": 1e-50, "ksp_max_it": 10000}


@pytest.mark.parametrize("cells", [(8, 8, 8), (10, 10), (16, 16)])
@pytest.mark.parametrize("pc", ["none", "jacobi", "fieldsplit", "fieldsplit_additive"])
def test_c_gmres_matches_python_oracle(cells, pc):
    """KSPGMRES(30) / PCFIELDSPLIT of the C oracle (used to pin BASELINE config 5 at 128^3) against the Python
    oracle, which reproduces the reference's stored GMRES histories and counts 10/40/292 (test_oracle_golden)."""
    osys, csys = _pair(cells, 1)
    okw = {"none": dict(pc_type="none"), "jacobi": dict(pc_type="jacobi"),
           "fieldsplit": dict(pc_type="fieldsplit", inner=INNER),
           "fieldsplit_additive": dict(pc_type="fieldsplit", fieldsplit_type="additive", inner=INNER)}[pc]
    ref = orc.solve_dpp_oracle(osys, "gmres", **okw)
    got = csys.gmres(pc, inner=INNER, history=len(ref.history) + 8)
    assert got.reason == ref.reason
    assert abs(got.iteration_number - ref.iteration_number) <= max(2, ref.iteration_number // 30)
    k = min(len(ref.history), len(got.history), 12)
    assert np.allclose(got.history[:k], ref.history[:k], rtol=1e-8)
    assert np.linalg.norm(got.u - ref.u) <= 1e-8 * np.linalg.norm(ref.u)
    if cells == (16, 16) and pc == "none":
        assert got.iteration_number == 292      # convergence.csv (README example size)


def test_c_gmres_config5_matches_python_oracle():
    kw = dict(k1=1.0, k2=1e-6, beta=1e2, mu=1.0)
    osys = orc.build_system(orc.structured_mesh((8, 8, 8), 1), orc.Params(**kw), ("const", 1.0, 0.0))
    csys = co.constant_bc_system((8, 8, 8), 1, **kw)
    for pc, okw in [("none", dict(pc_type="none")), ("jacobi", dict(pc_type="jacobi")),
                    ("fieldsplit", dict(pc_type="fieldsplit", inner=INNER))]:
        ref = orc.solve_dpp_oracle(osys, "gmres", **okw)
        got = csys.gmres(pc, inner=INNER)
        assert abs(got.iteration_number - ref.iteration_number) <= 2
        assert np.linalg.norm(got.u - ref.u) <= 1e-7 * np.linalg.norm(ref.u)
    ref = orc.solve_dpp_oracle(osys, "cg", "jacobi")
    assert csys.cg("jacobi").iteration_number == ref.iteration_number


def test_large_size_pins_belong_to_the_systems_they_name(golden_large):
    """tests/golden/large_sizes.json was produced by tests/golden/make_golden_large.py; the cheap part of each
    record (system size, ||b||, the iteration-0 norm ||D^-1 b||) is recomputed here from a fresh build."""
    for key, build, run in [("cfg3_128", lambda N: co.manufactured_system((N, N, N), 1), None),
                            ("cfg5_128", lambda N: co.constant_bc_system((N, N, N), 1), "cg_jacobi")]:
        pin = golden_large[key]
        csys = build(pin["cells"])
        assert (csys.n_dof, csys.nnz) == (pin["n_dof"], pin["nnz"])
        b = csys.rhs()
        assert np.linalg.norm(b) == pytest.approx(pin["rhs_norm2"], rel=1e-12)
        rec = pin if run is None else pin["runs"][run]
        first = csys.cg("jacobi", max_it=1, history=2)
        assert first.history[0] == pytest.approx(rec["history"][0], rel=1e-12)
        csys.close()
    pin = golden_large["cfg3_256"]
    assert pin["iterations"] == 387 and pin["n_dof"] == 2 * 257 ** 3 and pin["nnz"] == 4 * (3 * 256 + 1) ** 3
