import sys, time
sys.path.insert(0, '.')
import numpy as np
import perphil_b200 as pb
from tests.util import configured_handle
for N in (32, 64, 96):
    mesh = pb.UnitCubeMesh(N, N, N)
    _, V = pb.create_function_spaces(mesh, pressure_deg=2); W = V * V
    prm = pb.DPPParameters(k1=1.0, k2=1e-2, beta=1.0, mu=1.0)
    _, p1, _, p2 = pb.exact_expressions_3d(mesh, prm)
    bcs = [pb.DirichletBC(W.sub(0), p1, "on_boundary"), pb.DirichletBC(W.sub(1), p2, "on_boundary")]
    t0 = time.perf_counter(); h = configured_handle(W, prm, bcs); t1 = time.perf_counter()
    ms = h.time_apply(reps=5, warmup=2)
    ndof = 2 * h.n_nodes
    print(f"Q2 {N}^3: {ndof} DoF, family {h.info().kernel_family}, setup {t1-t0:.1f} s, apply {ms:.3f} ms -> {ndof/ms/1e6:.1f} GDoF/s", flush=True)
    if N <= 64:
        t0 = time.perf_counter()
        sol = pb.solve_dpp(W, prm, bcs, solver_parameters=pb.B200_CG_JACOBI_PARAMS)
        info = pb.last_solve_info()
        print(f"   Jacobi-CG its {sol.iteration_number} solve {info.solve_ms:.1f} ms", flush=True)
    pb.release_handles()
