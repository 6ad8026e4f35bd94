"""Host-side logic that needs no GPU: presets, option translation, mesh/space surface, forms,
parameter container (mirrors the reference's own CPU-only tests)."""
import numpy as np
import pytest

import perphil_b200 as pb
from perphil_b200 import _lib as L
from perphil_b200 import parameters as prm
from oracle import dpp_oracle as orc


def test_parameters_defaults():
    # models/dpp/_tests/test_parameters.py:10-23
    p = pb.DPPParameters()
    assert isinstance(p.k1, pb.Constant) and float(p.k1) == 1.0
    assert float(p.k2) == pytest.approx(1.0 / 1e2)
    p = pb.DPPParameters(k1=2.0, k2=None, scale_contrast=10.0)
    assert float(p.k2) == pytest.approx(0.2)
    p = pb.DPPParameters(k1=1.0, k2=1e-2, beta=1.0, mu=1.0)
    assert p.eta == pytest.approx(np.sqrt(1.01 / 1e-2))
    assert p.eta == pytest.approx(orc.Params(k2=1e-2).eta)


def test_presets_follow_reference_style():
    # solvers/_tests/test_solver_parameters.py:1-26 style checks on the new presets
    for name in dir(prm):
        if name.startswith("B200_") and name.endswith("_PARAMS"):
            d = getattr(prm, name)
            assert d[prm.B200_BACKEND_KEY] == prm.B200_BACKEND
    assert prm.B200_CG_JACOBI_PARAMS["ksp_type"] == "cg" and prm.B200_CG_JACOBI_PARAMS["pc_type"] == "jacobi"
    assert prm.B200_CG_JACOBI_PARAMS["ksp_rtol"] == 1e-8 and prm.B200_CG_JACOBI_PARAMS["ksp_atol"] == 1e-12
    assert prm.B200_CG_JACOBI_PARAMS["ksp_max_it"] == 50000
    assert prm.B200_GMRES_PARAMS["pc_type"] == "none"
    fs = prm.B200_GMRES_FIELDSPLIT_PARAMS
    assert fs["pc_type"] == "fieldsplit" and fs["pc_fieldsplit_type"] == "multiplicative"
    assert fs["pc_fieldsplit_0_fields"] == "0" and fs["pc_fieldsplit_1_fields"] == "1"
    assert prm.B200_CG_FIELDSPLIT_PARAMS["pc_fieldsplit_type"] == "additive"
    assert prm.B200_PICARD_SPLIT_PARAMS["snes_rtol"] == 1e-8


class _FakeHandle:
    def default_options(self):
        import ctypes

        o = L.DppOptions()
        L.load().dpp_default_options(ctypes.byref(o))
        return o


def test_option_translation():
    from perphil_b200.solver import options_from_petsc

    h = _FakeHandle()
    o = options_from_petsc(h, prm.B200_CG_JACOBI_PARAMS)
    assert (o.ksp_type, o.pc_type, o.operator_mode) == (L.KSP_CG, L.PC_JACOBI, L.OP_MATRIX_FREE)
    assert (o.rtol, o.atol, o.max_it) == (1e-8, 1e-12, 50000)
    o = options_from_petsc(h, prm.B200_CG_JACOBI_AIJ_PARAMS)
    assert o.operator_mode == L.OP_ASSEMBLED
    o = options_from_petsc(h, prm.B200_GMRES_FIELDSPLIT_PARAMS)
    assert (o.ksp_type, o.pc_type, o.fieldsplit_type) == (L.KSP_GMRES, L.PC_FIELDSPLIT, L.FS_MULTIPLICATIVE)
    assert (o.inner_ksp_type, o.inner_pc_type, o.inner_rtol) == (L.INNER_CG, L.PC_JACOBI, 1e-10)
    o = options_from_petsc(h, prm.B200_PICARD_SPLIT_PARAMS, nonlinear=True)
    assert o.ksp_type == L.KSP_PICARD and o.rtol == 1e-8
    # PETSc defaults when keys are absent; ksp_rtol: Firedrake's injected default 1e-7
    o = options_from_petsc(h, {"ksp_type": "gmres"})
    assert (o.rtol, o.atol, o.max_it, o.gmres_restart, o.pc_type) == (1e-7, 1e-50, 10000, 30, L.PC_NONE)
    # a restart length the device-resident GMRES cannot hold is refused, not clamped
    with pytest.raises(NotImplementedError):
        options_from_petsc(h, {"ksp_type": "gmres", "ksp_gmres_restart": 50})
    # the reference's SNES NGS presets are refused, not silently replaced by block Picard
    with pytest.raises(NotImplementedError):
        options_from_petsc(h, {"snes_type": "ngs"}, nonlinear=True)
    # reference presets that need MUMPS/ILU (K8) are refused, not silently replaced
    with pytest.raises(NotImplementedError):
        options_from_petsc(h, {"ksp_type": "preonly", "pc_type": "lu"})
    with pytest.raises(NotImplementedError):
        options_from_petsc(h, {"ksp_type": "gmres", "pc_type": "ilu"})
    lu = {"ksp_type": "preonly", "pc_type": "lu"}
    with pytest.raises(NotImplementedError):
        options_from_petsc(h, {"ksp_type": "gmres", "pc_type": "fieldsplit", "fieldsplit_0": lu, "fieldsplit_1": lu})


@pytest.mark.parametrize("cells,degree", [((4, 5), 1), ((3, 4, 5), 1), ((3, 2), 2), ((2, 3, 2), 2)])
def test_mesh_matches_oracle_mesh(cells, degree):
    mesh = pb.UnitSquareMesh(*cells) if len(cells) == 2 else pb.UnitCubeMesh(*cells)
    _, V = pb.create_function_spaces(mesh, pressure_deg=degree)
    om = orc.structured_mesh(cells, degree)
    assert np.array_equal(V.cell_node_map().values, om.cell_node_map)
    assert np.allclose(V.node_coordinates, om.coords)
    assert np.array_equal(V.boundary_nodes, om.boundary_nodes)
    assert np.array_equal(mesh.coordinates.cell_node_map().values, om.cell_vertex_map)
    assert np.allclose(mesh.coordinates.dat.data_ro, om.vertex_coords)
    W = V * V
    assert W.num_sub_spaces() == 2 and W.dim() == 2 * om.n_nodes and W.sub(1).index == 1


def test_mesh_contract():
    # mesh/_tests/test_mesh.py:10-20, forms/_tests/test_spaces.py:11-18
    mesh = pb.create_mesh(2, 2, quadrilateral=True)
    assert mesh.geometric_dimension() == 2 and mesh.num_cells() == 4
    _, V = pb.create_function_spaces(mesh)
    assert V.dim() == 9
    with pytest.raises(ValueError):
        pb.create_mesh(2, 2, quadrilateral=False)


def test_forms_structure():
    # forms/_tests/test_dpp_regressions/test_dpp_form_structure_regression.yml: integrals 4, rank 2
    mesh = pb.create_mesh(2, 2)
    _, V = pb.create_function_spaces(mesh)
    W = V * V
    p = pb.DPPParameters()
    a, Lf = pb.dpp_form(W, p)
    assert len(a.integrals()) == 4 and len(a.arguments()) == 2 and len(Lf.arguments()) == 1
    with pytest.raises(ValueError):
        pb.dpp_form(V, p)
    F, fields = pb.dpp_splitted_form(W, p)
    assert isinstance(fields, pb.Function) and len(F.arguments()) == 1
    with pytest.raises(ValueError):
        pb.dpp_splitted_form(V, p)
    (a0, L0), (a1, L1) = pb.dpp_delayed_form(V, V, p, pb.Function(V), pb.Function(V))
    assert a0.blocks == ((0, 0),) and a1.blocks == ((1, 1),)
    assert [i.coefficient for i in a1.integrals()] == [1e-2, 1.0]


def test_manufactured_matches_oracle():
    p = pb.DPPParameters(k1=1.0, k2=1e-2, beta=1.0, mu=1.0)
    for cells in [(4, 4), (3, 3, 3)]:
        mesh = pb.UnitSquareMesh(*cells) if len(cells) == 2 else pb.UnitCubeMesh(*cells)
        _, V = pb.create_function_spaces(mesh)
        _, e1, _, e2 = pb.exact_expressions(mesh, p)
        o1, o2 = orc.exact_pressures(V.node_coordinates, orc.Params(k2=1e-2))
        assert np.array_equal(e1(V.node_coordinates), o1) and np.array_equal(e2(V.node_coordinates), o2)


def test_manufactured_velocities_are_the_darcy_fluxes_of_the_pressures():
    """utils/manufactured_solutions.py:21-37 (2-D, written out) and :72-81 (3-D, -(k/mu) grad p): checked against
    central differences of the pressures, and the 2-D formula against the literal reference expression."""
    p = pb.DPPParameters(k1=1.3, k2=2e-2, beta=0.7, mu=1.9)
    k1, k2, beta, mu, eta = float(p.k1), float(p.k2), float(p.beta), float(p.mu), p.eta
    rng = np.random.default_rng(0)
    for dim in (2, 3):
        mesh = pb.UnitSquareMesh(2, 2) if dim == 2 else pb.UnitCubeMesh(2, 2, 2)
        u1, p1, u2, p2 = pb.exact_expressions(mesh, p) if dim == 2 else pb.exact_expressions_3d(mesh, p)
        X = rng.random((40, dim))
        eps = 1e-6
        for u, pr, k in [(u1, p1, k1), (u2, p2, k2)]:
            g = np.stack([(pr(X + eps * np.eye(dim)[d]) - pr(X - eps * np.eye(dim)[d])) / (2 * eps) for d in range(dim)], axis=1)
            assert u(X).shape == (40, dim)
            assert np.allclose(u(X), -(k / mu) * g, rtol=1e-6, atol=1e-6 * np.abs(g).max())
        if dim == 2:
            x, y = X[:, 0], X[:, 1]
            lit1 = np.stack([-k1 * (np.exp(np.pi * x) * np.sin(np.pi * y)),
                             -k1 * (np.exp(np.pi * x) * np.cos(np.pi * y) - (eta / (beta * k1)) * np.exp(eta * y))], axis=1)
            lit2 = np.stack([-k2 * (np.exp(np.pi * x) * np.sin(np.pi * y)),
                             -k2 * (np.exp(np.pi * x) * np.cos(np.pi * y) + (eta / (beta * k2)) * np.exp(eta * y))], axis=1)
            assert np.allclose(u1(X), lit1, rtol=1e-14) and np.allclose(u2(X), lit2, rtol=1e-14)


def test_oracle_darcy_velocity_converges_to_the_manufactured_velocity():
    """Pins oracle.darcy_velocity (the checker of dpp_darcy_velocity) to the reference's own exact velocities:
    the L2 projection of -k grad(I_h p) converges to u = -(k/mu) grad p (mu = 1) at the nodes (rate between 1 and 2:
    the boundary rows of the projection pollute at O(h^1.5); measured ratios per halving 2.5, 2.65)."""
    prm = pb.DPPParameters(k1=1.0, k2=1e-2, beta=1.0, mu=1.0)
    errs = []
    for N in (16, 32, 64):
        mesh = pb.UnitSquareMesh(N, N)
        _, V = pb.create_function_spaces(mesh)
        u1, p1, _, _ = pb.exact_expressions(mesh, prm)
        om = orc.structured_mesh((N, N), 1)
        v = orc.darcy_velocity(om, p1(V.node_coordinates), float(prm.k1))      # [dim, n]
        ex = u1(V.node_coordinates).T
        inner = np.ones(V.dim(), bool); inner[V.boundary_nodes] = False
        errs.append(np.sqrt(np.mean((v - ex)[:, inner] ** 2)) / np.sqrt(np.mean(ex ** 2)))
    assert errs[1] < errs[0] / 2.2 and errs[2] < errs[1] / 2.4 and errs[2] < 1e-2


def test_bc_data_extraction():
    from perphil_b200.provider import bc_data, space_data

    mesh = pb.UnitCubeMesh(3, 3, 3)
    _, V = pb.create_function_spaces(mesh)
    W = V * V
    bcs = [pb.DirichletBC(W.sub(0), pb.Constant(1.0), "on_boundary"), pb.DirichletBC(W.sub(1), lambda X: X[:, 0], "on_boundary")]
    out = bc_data(W, bcs)
    assert [f for f, _, _ in out] == [0, 1]
    assert np.all(out[0][2] == 1.0) and np.allclose(out[1][2], V.node_coordinates[V.boundary_nodes, 0])
    sd = space_data(W)
    assert (sd.dim, sd.degree, sd.n_nodes) == (3, 1, 64)


def test_local_order_normalisation():
    from perphil_b200.provider import normalise_local_order

    om = orc.structured_mesh((3, 2, 2), 1)
    rng = np.random.default_rng(0)
    perm = rng.permutation(8)
    scrambled = om.cell_node_map[:, perm]
    assert np.array_equal(normalise_local_order(scrambled, om.coords), om.cell_node_map)
    # per-cell different orders
    s2 = om.cell_node_map.copy()
    for c in range(s2.shape[0]):
        s2[c] = s2[c, rng.permutation(8)]
    assert np.array_equal(normalise_local_order(s2, om.coords), om.cell_node_map)


@pytest.mark.parametrize("cells,degree", [((5, 6, 4), 1), ((3, 3, 4), 2), ((7, 9), 1), ((5, 4), 2)])
def test_lattice_detection_on_scrambled_numbering(cells, degree):
    """A DMPlex-style arbitrary numbering of a tensor grid is recognised and mapped to lexicographic order."""
    from perphil_b200.lattice import detect_lattice
    from tests.test_gpu_parity import _shuffled_distorted

    m = _shuffled_distorted(cells, degree, 0.0, seed=3)
    lat = detect_lattice(m.dim, degree, m.cell_node_map, m.coords, m.vertex_coords, m.cell_vertex_map)
    assert lat is not None and not lat.is_identity and lat.cells == tuple(cells)
    ref = orc.structured_mesh(cells, degree)
    # node u sits at lexicographic position perm[u]
    assert np.allclose(ref.coords[lat.perm], m.coords, atol=1e-12)
    assert np.array_equal(lat.cell_node_map, ref.cell_node_map)
    assert np.allclose(lat.vertex_coords, ref.vertex_coords)
    # a lexicographic mesh maps to itself; a distorted one is not a lattice
    assert detect_lattice(ref.dim, degree, ref.cell_node_map, ref.coords).is_identity
    d = _shuffled_distorted(cells, degree, 0.3, seed=3)
    assert detect_lattice(d.dim, degree, d.cell_node_map, d.coords, d.vertex_coords, d.cell_vertex_map) is None


def test_approach_routing_follows_the_reference_harness():
    """experiments/iterative_bench.py:31-48,157-188: same enum values where an approach exists in the reference;
    MUMPS/ILU approaches are refused on the B200 path."""
    from perphil_b200 import experiments as ex

    assert ex.Approach.PLAIN_GMRES.value == "GMRES"
    assert ex.Approach.SS_GMRES.value == "Scale-Splitting GMRES"
    assert ex.Approach.MONOLITHIC_MUMPS.value == "Monolithic LU with MUMPS"
    p = ex.params_for(ex.Approach.SS_GMRES)
    assert p["pc_type"] == "fieldsplit" and p["pc_fieldsplit_type"] == "multiplicative" and p[prm.B200_BACKEND_KEY] == "b200"
    assert ex.params_for("GMRES")["pc_type"] == "none"
    p["ksp_rtol"] = 0.5
    assert ex.params_for(ex.Approach.SS_GMRES)["ksp_rtol"] == 1e-8   # a copy is returned
    for a in (ex.Approach.GMRES_ILU, ex.Approach.SS_GMRES_ILU, ex.Approach.MONOLITHIC_MUMPS):
        with pytest.raises(NotImplementedError):
            ex.params_for(a)
    mp = ex.default_model_params()
    assert float(mp.k2) == 1e-2 and float(mp.k1) == 1.0
    mesh = pb.UnitSquareMesh(3, 3)
    _, V = pb.create_function_spaces(mesh)
    bcs = ex.default_bcs(V * V)
    assert len(bcs) == 2 and np.all(bcs[0].values() == 0.0)
    assert "time_KSPSolve" in ex.CSV_COLUMNS and "time_MatMult" in ex.CSV_COLUMNS


def test_calculate_condition_number_matches_the_reference_table(golden):
    """solvers/conditioning.py:105-218 semantics (dense SVD, zero_tol filter) on the oracle's assembled matrices:
    conditioning_3d.csv N=4 and conditioning.csv N=8, including the macro/micro blocks."""
    from oracle import dpp_oracle as orc
    from perphil_b200.conditioning import calculate_condition_number

    row = next(r for r in golden["conditioning_3d_hex_q1"] if r["N"] == 4)
    osys = orc.build_system(orc.structured_mesh((4, 4, 4), 1), orc.Params(k1=1.0, k2=1e-2), "manufactured")
    A = osys.A_bc.tocsr()
    n = osys.n_nodes
    assert abs(calculate_condition_number(A, None) - row["cond_monolithic"]) < 1e-9 * row["cond_monolithic"]
    assert abs(calculate_condition_number(A[:n, :n], 0) - row["cond_macro"]) < 1e-9 * row["cond_macro"]
    assert abs(calculate_condition_number(A[n:, n:], n) - row["cond_micro"]) < 1e-9 * row["cond_micro"]
    # sparse route (ARPACK at both ends) agrees to its tolerance
    ks = calculate_condition_number(A, 6, use_sparse=True)
    assert abs(ks - row["cond_monolithic"]) < 1e-5 * row["cond_monolithic"]
    import scipy.sparse as sp

    assert np.isnan(calculate_condition_number(sp.csr_matrix((0, 0)), None))


def test_oracle_darcy_projection_reproduces_polynomial_gradients():
    """Galerkin projection of -k grad(p_h): exact whenever grad(p_h) lies in the space (Q1: linear p; Q2:
    p = x^2 + xy)."""
    from oracle import dpp_oracle as orc

    m = orc.structured_mesh((4, 5, 3), 1)
    X = m.coords
    v = orc.darcy_velocity(m, 2 * X[:, 0] - 3 * X[:, 1] + 0.5 * X[:, 2], 2.0)
    assert np.abs(v - np.array([-4.0, 6.0, -1.0])[:, None]).max() < 1e-12
    m = orc.structured_mesh((3, 4), 2)
    X = m.coords
    v = orc.darcy_velocity(m, X[:, 0] ** 2 + X[:, 0] * X[:, 1], 1.0)
    assert np.abs(v[0] + 2 * X[:, 0] + X[:, 1]).max() < 1e-12 and np.abs(v[1] + X[:, 0]).max() < 1e-12


def test_slice_along_x_is_point_evaluation_of_the_fe_function():
    """utils/postprocessing.py:66-86: values at (x, y_j) for the distinct node ordinates y_j; exact for fields
    the space represents (Q1: bilinear, Q2: biquadratic), also off the node lines."""
    for deg in (1, 2):
        mesh = pb.UnitSquareMesh(5, 4)
        _, V = pb.create_function_spaces(mesh, pressure_deg=deg)
        X = V.node_coordinates
        f = pb.Function(V, val=2.0 * X[:, 0] ** deg + 3.0 * X[:, 1] * X[:, 0])
        for x in (0.0, 0.37, 0.6, 1.0):
            y, v = pb.slice_along_x(f, x)
            assert y.size == deg * 4 + 1
            assert np.abs(v - (2.0 * x ** deg + 3.0 * y * x)).max() < 1e-13
    with pytest.raises(ValueError):
        pb.slice_along_x(f, 1.5)


def test_morton_permutation_is_a_locality_preserving_permutation():
    """lattice.morton_permutation (numbering map of non-lattice meshes for the element-based kernels)."""
    from perphil_b200.lattice import detect_lattice, morton_permutation
    import sys
    sys.path.insert(0, ".")
    from tools.general_mesh import shuffled_distorted_hex

    cnm, X, bn = shuffled_distorted_hex(8, 0.25, seed=2)
    assert cnm.shape == (512, 8) and X.shape == (729, 3) and bn.size == 729 - 343
    assert detect_lattice(3, 1, cnm, X, X, cnm) is None          # distorted: not a tensor grid
    perm = morton_permutation(X)
    assert perm.dtype == np.int32 and np.array_equal(np.sort(perm), np.arange(729))
    # locality: the 8 nodes of a cell end up close in the new numbering (random numbering: spread ~ n / 2)
    spread = np.ptp(perm[cnm], axis=1)
    assert np.median(spread) < 729 / 6 and np.median(np.ptp(cnm, axis=1)) > 729 / 3
    # an undistorted shuffled lattice IS detected and gets the lexicographic map instead
    cnm2, X2, _ = shuffled_distorted_hex(5, 0.0, seed=3)
    lat = detect_lattice(3, 1, cnm2, X2, X2, cnm2)
    assert lat is not None and lat.cells == (5, 5, 5)


def test_bench_cpu_arm_sets_its_thread_count_itself(monkeypatch):
    """under torch.distributed.run OMP_NUM_THREADS=1 silently changed the CPU baseline."""
    import bench
    from oracle import c_oracle as co

    monkeypatch.setenv("OMP_NUM_THREADS", "1")
    n = bench.host_threads()
    assert n >= 1 and isinstance(bench.cpu_model(), str)
    out = bench.cpu_baseline(12, single_thread_size=8)
    assert out["cores"] == n == co.num_threads() and out["kind"] == "port"
    assert out["single_thread"]["cores"] == 1 and out["single_thread"]["value"] > 0
    assert out["iterations"] == co.manufactured_system((12, 12, 12), 1).cg("jacobi").iteration_number
