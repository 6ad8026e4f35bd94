"""Small solves through every kernel family, meant to run under compute-sanitizer (memcheck / racecheck /
synccheck): fused Jacobi-CG (TMA kernels), GMRES, fieldsplit, block Picard, the general (unstructured) kernels on
a shuffled + distorted mesh, Q2, CSR assembly + SpMV, error norms, Darcy velocity, Lanczos.  Prints one OK line
per family; any sanitizer finding is reported by the tool itself."""
import sys
sys.path.insert(0, '.')
import numpy as np
import perphil_b200 as pb
from perphil_b200.backend import DppHandle
from tests.util import make_problem, configured_handle

N = int(sys.argv[1]) if len(sys.argv) > 1 else 10
which = sys.argv[2] if len(sys.argv) > 2 else "all"

def on(name):
    return which in ("all", name)

W, prm, bcs, osys = make_problem((N, N, N), 1)
if on("cg"):
    for preset in (pb.B200_CG_JACOBI_PARAMS, pb.B200_CG_PARAMS, pb.B200_CG_PBJACOBI_PARAMS):
        s = pb.solve_dpp(W, prm, bcs, solver_parameters=preset)
        print("OK cg", preset["pc_type"], s.iteration_number, flush=True)
if on("gmres"):
    for preset in (pb.B200_GMRES_PARAMS, pb.B200_GMRES_JACOBI_PARAMS, pb.B200_GMRES_FIELDSPLIT_PARAMS, pb.B200_CG_FIELDSPLIT_PARAMS):
        s = pb.solve_dpp(W, prm, bcs, solver_parameters=preset)
        print("OK", preset["ksp_type"], preset["pc_type"], s.iteration_number, flush=True)
if on("picard"):
    s = pb.solve_dpp_nonlinear(W, prm, bcs, solver_parameters=pb.B200_PICARD_SPLIT_PARAMS)
    print("OK picard", s.iteration_number, flush=True)
if on("csr"):
    a, _ = pb.dpp_form(W, prm)
    md = pb.get_matrix_data_from_form(a, bcs)
    s = pb.solve_dpp(W, prm, bcs, solver_parameters=pb.B200_CG_JACOBI_AIJ_PARAMS)
    print("OK csr", md.number_of_nonzero_entries, s.iteration_number, flush=True)
if on("post"):
    s = pb.solve_dpp(W, prm, bcs, solver_parameters=pb.B200_CG_JACOBI_PARAMS)
    p1, p2 = pb.split_dpp_solution(s.solution)
    _, e1, _, e2 = pb.exact_expressions_3d(W.mesh(), prm)
    print("OK norms", pb.l2_error(p1, e1), pb.h1_seminorm_error(p2, e2), flush=True)
    v = pb.calculate_darcy_velocity_from_pressure(p1, prm.k1)
    print("OK darcy", v.cg_iterations, flush=True)
    est = pb.condition_number_matrix_free(pb.dpp_form(W, prm)[0], bcs, rtol=1e-6)
    print("OK lanczos", est.condition_number, flush=True)
pb.release_handles()
if on("q2"):
    W2, prm2, bcs2, _ = make_problem((max(N // 2, 3),) * 3, 2)
    s = pb.solve_dpp(W2, prm2, bcs2, solver_parameters=pb.B200_CG_JACOBI_PARAMS)
    print("OK q2 cg", s.iteration_number, flush=True)
    s = pb.solve_dpp_nonlinear(W2, prm2, bcs2, solver_parameters=pb.B200_PICARD_SPLIT_PARAMS)
    print("OK q2 picard", s.iteration_number, flush=True)
    pb.release_handles()
if on("general"):
    from tests.test_gpu_parity import _shuffled_distorted
    for degree in (1, 2):
        m2 = _shuffled_distorted((5, 6, 4), degree, 0.3, seed=4)
        h = DppHandle(m2.dim, degree, m2.cell_node_map, m2.vertex_coords, m2.cell_vertex_map, n_nodes=m2.n_nodes)
        h.set_params(1.0, 1e-2, 1.0, 1.0)
        nb = m2.boundary_nodes
        rng = np.random.default_rng(1)
        h.set_dirichlet(0, nb, rng.standard_normal(nb.size)); h.set_dirichlet(1, nb, rng.standard_normal(nb.size))
        y = h.apply(rng.standard_normal(2 * m2.n_nodes))
        u, info = h.solve()
        print("OK general deg", degree, info.iterations, float(np.linalg.norm(y)), flush=True)
        h.close()
print("SANITIZE-SCRIPT DONE")
