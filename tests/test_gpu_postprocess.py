"""GPU parity of the SURVEY 8(f) items 3 and 4: Darcy-velocity projection (csrc/darcy.cu) against the oracle's
Galerkin projection, and Lanczos condition numbers (krylov.cu: krylov_lanczos) against the numbers the
reference stores in notebooks/results-conforming-*/conditioning/*.csv (dense SVD there)."""
import numpy as np
import pytest

import perphil_b200 as pb
from oracle import dpp_oracle as orc
from perphil_b200 import _lib as L
from tests.test_gpu_parity import _shuffled_distorted
from tests.util import configured_handle, make_problem, rel_err

pytestmark = pytest.mark.gpu


def _pressure(coords, kind):
    x = coords[:, 0]
    y = coords[:, 1]
    z = coords[:, 2] if coords.shape[1] == 3 else 0.0 * x
    if kind == "linear":
        return 2.0 * x - 3.0 * y + 0.5 * z
    return np.exp(0.7 * x) * np.sin(2.0 * y + 0.3) + x * z * z - 0.4 * y * z


@pytest.mark.parametrize("cells,degree", [((6, 5), 1), ((4, 3), 2), ((5, 6, 4), 1), ((3, 2, 4), 2), ((16, 16, 16), 1)])
def test_darcy_velocity_vs_oracle(cells, degree):
    """-k grad(p_h) projected into V^dim: 1e-10 relative per component against the sparse-direct projection
    (the GPU solves the mass systems with Jacobi-CG to rtol 1e-13 here; mass matrices are well conditioned)."""
    W, prm, bcs, osys = make_problem(cells, degree)
    h = configured_handle(W, prm, bcs)
    X = osys.mesh.coords
    k = 0.37
    for kind in ("linear", "smooth"):
        p = _pressure(X, kind)
        ref = orc.darcy_velocity(osys.mesh, p, k)
        vel, its = h.darcy_velocity(k, p=p, rtol=1e-13)
        assert vel.shape == ref.shape and its.size == len(cells) and its.max() < 200
        scale = np.abs(ref).max()
        for c in range(len(cells)):
            assert np.abs(vel[c] - ref[c]).max() <= 1e-10 * scale
        if kind == "linear":   # the gradient of a linear field is reproduced exactly
            g = np.array([2.0, -3.0, 0.5])[: len(cells)]
            for c in range(len(cells)):
                assert np.abs(vel[c] + k * g[c]).max() < 1e-10
    pb.release_handles()


def test_darcy_velocity_public_api_and_last_solution():
    """calculate_darcy_velocity_from_pressure(p1_h, k1) on the split solution of solve_dpp
    (notebooks/conforming-galerkin-fem-operator-splitting-2D-perphil.py:109-113) and the field-of-the-last-solve
    shortcut of the C ABI agree with the oracle projection of the same nodal pressure."""
    W, prm, bcs, osys = make_problem((10, 10), 1)
    sol = pb.solve_dpp(W, prm, bcs, solver_parameters=pb.B200_CG_JACOBI_PARAMS)
    p1, p2 = pb.split_dpp_solution(sol.solution)
    u1 = pb.calculate_darcy_velocity_from_pressure(p1, prm.k1, rtol=1e-12)
    u2 = pb.calculate_darcy_velocity_from_pressure(p2, float(prm.k2), degree=1, rtol=1e-12)
    assert u1.dat.data.shape == (osys.mesh.n_nodes, 2)
    for u, p, k in ((u1, p1, prm.k1), (u2, p2, prm.k2)):
        ref = orc.darcy_velocity(osys.mesh, np.asarray(p.dat.data), float(k))
        assert np.abs(u.dat.data.T - ref).max() <= 1e-9 * np.abs(ref).max()
    h = pb.handle_for(W)
    v_last, _ = h.darcy_velocity(float(prm.k2), p=None, field=1, rtol=1e-12)
    assert np.abs(v_last - u2.dat.data.T).max() <= 1e-9 * np.abs(v_last).max()
    with pytest.raises(NotImplementedError):
        pb.calculate_darcy_velocity_from_pressure(p1, prm.k1, degree=2)
    pb.release_handles()


@pytest.mark.parametrize("cells,degree,distort", [((5, 6, 4), 1, 0.0), ((5, 6, 4), 1, 0.3), ((7, 9), 1, 0.3),
                                                  ((3, 3, 4), 2, 0.25)])
def test_darcy_velocity_unstructured_and_renumbered(cells, degree, distort):
    """Shuffled numbering (numbering map + structured kernels) and distorted cells (general kernels)."""
    from perphil_b200.backend import DppHandle

    m2 = _shuffled_distorted(cells, degree, distort, seed=31)
    h = DppHandle.from_mesh_arrays(m2.dim, degree, m2.cell_node_map, m2.coords, m2.vertex_coords, m2.cell_vertex_map,
                                   n_nodes=m2.n_nodes)
    assert h.info().kernel_family == (L.KERNEL_GENERAL if distort else L.KERNEL_STRUCTURED)
    h.set_params(1.0, 1e-2, 1.0, 1.0)
    p = _pressure(m2.coords, "smooth")
    ref = orc.darcy_velocity(m2, p, 1.3)
    vel, _ = h.darcy_velocity(1.3, p=p, rtol=1e-13)
    assert np.abs(vel - ref).max() <= 1e-10 * np.abs(ref).max()
    h.close()


def test_darcy_velocity_converges_to_the_manufactured_velocities():
    """Reference-derived check of the Darcy row: the projected velocity of the interpolated exact pressure converges
    to u_i = -(k_i/mu) grad p_i of utils/manufactured_solutions.py:21-37 (2-D) and :72-81 (3-D); the GPU projection is
    bitwise repeatable (ordered gather, no atomics)."""
    prm = pb.DPPParameters(k1=1.0, k2=1e-2, beta=1.0, mu=1.0)
    for dim, sizes in [(2, (16, 32, 64)), (3, (8, 16, 32))]:
        errs = {0: [], 1: []}
        for N in sizes:
            mesh = pb.UnitSquareMesh(N, N) if dim == 2 else pb.UnitCubeMesh(N, N, N)
            _, V = pb.create_function_spaces(mesh)
            W = V * V
            u1, p1, u2, p2 = pb.exact_expressions(mesh, prm)
            u1e, p1e, u2e, p2e = pb.interpolate_exact(mesh, None, W.sub(0), prm)
            assert u1e.dat.data.shape == (V.dim(), dim)
            inner = np.ones(V.dim(), bool); inner[V.boundary_nodes] = False
            for f, (pe, ue, k) in enumerate([(p1e, u1e, prm.k1), (p2e, u2e, prm.k2)]):
                vel = pb.calculate_darcy_velocity_from_pressure(pe, k, rtol=1e-12)
                again = pb.calculate_darcy_velocity_from_pressure(pe, k, rtol=1e-12)
                assert np.array_equal(vel.dat.data, again.dat.data)
                d = (vel.dat.data - ue.dat.data)[inner]
                errs[f].append(np.sqrt(np.mean(d ** 2)) / np.sqrt(np.mean(ue.dat.data ** 2)))
            pb.release_handles()
        for f in (0, 1):
            e = errs[f]
            assert e[1] < e[0] / 2.0 and e[2] < e[1] / 2.0, (dim, f, e)
            assert e[2] < (1e-2 if dim == 2 else 4e-2), (dim, f, e)


@pytest.mark.parametrize("N", [4, 8, 12])
def test_lanczos_condition_numbers_match_conditioning_3d_csv(golden, N):
    """kappa(A), kappa(A00), kappa(A11) of the 3-D hex Q1 system with manufactured BCs, as stored by the reference
    (dense SVD of the assembled matrix incl. the identity rows), from Lanczos on the matrix-free GPU operator."""
    row = next(r for r in golden["conditioning_3d_hex_q1"] if r["N"] == N)
    W, prm, bcs, osys = make_problem((N, N, N), 1)
    form, _ = pb.dpp_form(W, prm)
    mono = pb.condition_number_matrix_free(form, bcs, rtol=1e-10)
    assert mono.converged and abs(mono.condition_number - row["cond_monolithic"]) <= 1e-7 * row["cond_monolithic"]
    macro = pb.condition_number_matrix_free(form, bcs, block=0, rtol=1e-10)
    micro = pb.condition_number_matrix_free(form, bcs, block=1, rtol=1e-10)
    assert abs(macro.condition_number - row["cond_macro"]) <= 1e-7 * row["cond_macro"]
    assert abs(micro.condition_number - row["cond_micro"]) <= 1e-7 * row["cond_micro"]
    pb.release_handles()


@pytest.mark.parametrize("N", [8, 32])
def test_lanczos_condition_numbers_match_conditioning_2d_csv(golden, N):
    row = next(r for r in golden["conditioning_2d_quad_q1"] if r["N"] == N)
    W, prm, bcs, osys = make_problem((N, N), 1, bc="homogeneous")
    est = pb.condition_number_matrix_free(pb.dpp_form(W, prm)[0], bcs, rtol=1e-10)
    assert est.converged and abs(est.condition_number - row["cond_monolithic"]) <= 1e-7 * row["cond_monolithic"]
    pb.release_handles()


def test_lanczos_tridiagonal_is_a_projection_of_the_operator():
    """T = V^T A V for the first steps (before orthogonality degrades): alpha_0 and beta_0 follow from one
    operator application to the normalised start vector, reproduced on the host through dpp_apply."""
    W, prm, bcs, osys = make_problem((5, 4, 3), 1)
    h = configured_handle(W, prm, bcs)
    a, b = h.lanczos(40, which=0, seed=3)
    A = osys.A_bc.toarray()
    ev = np.linalg.eigvalsh(0.5 * (A + A.T))
    from scipy.linalg import eigvalsh_tridiagonal

    th = eigvalsh_tridiagonal(a, b[:-1])
    assert abs(th.max() - ev.max()) <= 1e-9 * ev.max() and th.min() >= ev.min() * (1 - 1e-9)
    # every Ritz value lies inside the spectrum of the symmetric operator
    assert th.max() <= ev.max() * (1 + 1e-12) and th.min() >= ev.min() * (1 - 1e-12)
    # large-N scaling run finishes and is monotone (kappa grows with the steps until it converges)
    W2, prm2, bcs2, _ = make_problem((32, 32, 32), 1)
    e1 = pb.condition_number_matrix_free(pb.dpp_form(W2, prm2)[0], bcs2, rtol=1e-6)
    assert e1.converged and e1.condition_number > 3305.0   # kappa(N=32) > kappa(N=16) of the stored table
    pb.release_handles()


def test_slices_at_x_half_match_the_notebook(golden):
    """`slice_along_x(p_h, 0.5)` of the B200 solution on the 10x10 quad mesh reproduces the values the
    operator-splitting notebook prints for the monolithic LU solve (ipynb cell 15: p1_h, p2_h at 11 y-values),
    through the reference's own call sequence solve_dpp -> split_dpp_solution -> slice_along_x."""
    g = golden["operator_splitting_notebook_10x10"]
    W, prm, bcs, _ = make_problem((10, 10), 1)
    sol = pb.solve_dpp(W, prm, bcs, solver_parameters={**pb.B200_CG_JACOBI_PARAMS, "ksp_rtol": 1e-13})
    p1, p2 = pb.split_dpp_solution(sol.solution)
    y_ref, p1_ref, p2_ref = g["slice_x0.5_monolithic_lu"]
    y, v1 = pb.slice_along_x(p1, 0.5)
    _, v2 = pb.slice_along_x(p2, 0.5)
    assert np.allclose(y, y_ref)
    assert np.allclose(v1, p1_ref, rtol=2e-8) and np.allclose(v2, p2_ref, rtol=2e-8)
    pb.release_handles()
