"""Device timings of the SURVEY 8(f) rows 3 and 4 on one B200: Darcy-velocity projection and the Lanczos
condition-number estimate on 3-D hex Q1 meshes (wall clock around the C-ABI calls, host buffers in and out)."""
import json, sys, time
sys.path.insert(0, '.')
import numpy as np
import perphil_b200 as pb

out = []
for N in (int(a) for a in (sys.argv[1:] or ["64", "128", "256"])):
    mesh = pb.UnitCubeMesh(N, N, N)
    _, V = pb.create_function_spaces(mesh); W = V * V
    prm = pb.DPPParameters(k1=1.0, k2=1e-2, beta=1.0, mu=1.0)
    _, p1, _, p2 = pb.exact_expressions_3d(mesh, prm)
    bcs = [pb.DirichletBC(W.sub(0), p1, "on_boundary"), pb.DirichletBC(W.sub(1), p2, "on_boundary")]
    sol = pb.solve_dpp(W, prm, bcs, solver_parameters=pb.B200_CG_JACOBI_PARAMS)
    p1h, _ = pb.split_dpp_solution(sol.solution)
    h = pb.handle_for(W)
    n = h.n_nodes
    row = {"N": N, "n_nodes": n}
    best = None
    for _ in range(3):
        t0 = time.perf_counter()
        vel, its = h.darcy_velocity(float(prm.k1), p=None, field=0, rtol=1e-8)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    # accuracy of the recovered velocity against -k1 grad(p1_exact) at the nodes (O(h) for Q1)
    row.update(darcy_ms=best * 1e3, darcy_cg_its=[int(i) for i in its], darcy_mnodes_per_s=n / best / 1e6)
    if N <= 128:
        t0 = time.perf_counter()
        est = pb.condition_number_matrix_free(pb.dpp_form(W, prm)[0], bcs, rtol=1e-6)
        row.update(lanczos_s=time.perf_counter() - t0, lanczos_steps=est.lanczos_steps, kappa=est.condition_number,
                   sigma_max=est.sigma_max, sigma_min=est.sigma_min, converged=est.converged)
    else:   # fixed number of steps: time per Lanczos step at the headline size
        t0 = time.perf_counter()
        a, b = h.lanczos(200, which=0, seed=0)
        dt = time.perf_counter() - t0
        row.update(lanczos_us_per_step=dt / a.size * 1e6, lanczos_steps=int(a.size))
    print(json.dumps(row), flush=True)
    out.append(row)
    pb.release_handles()
