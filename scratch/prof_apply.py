import sys; sys.path.insert(0, '.')
import numpy as np, time
import perphil_b200 as pb
from tests.util import configured_handle
N = int(sys.argv[1]) if len(sys.argv) > 1 else 128
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
mesh = pb.UnitCubeMesh(N, N, N); _, V = pb.create_function_spaces(mesh); W = V * V
prm = pb.DPPParameters(k1=1.0, k2=1e-2, beta=1.0, mu=1.0)
_, p1, _, p2 = pb.exact_expressions_3d(mesh, prm)
bcs = [pb.DirichletBC(W.sub(0), p1, "on_boundary"), pb.DirichletBC(W.sub(1), p2, "on_boundary")]
t0 = time.time(); h = configured_handle(W, prm, bcs); print("handle", round(time.time() - t0, 2), "s")
ms = h.time_apply(reps=reps, warmup=2, with_dot=True)
n = (N + 1) ** 3
print(f"N={N} apply {ms:.4f} ms  {34*n/ms/1e6:.1f} GB/s (34 B/node)  {2*n/ms/1e6:.2f} GDoF/s")
if len(sys.argv) > 3:
    sol = pb.solve_dpp(W, prm, bcs, solver_parameters=pb.B200_CG_JACOBI_PARAMS)
    i = pb.last_solve_info(); print("solve its", sol.iteration_number, "ms", i.solve_ms, "per-it", i.solve_ms / sol.iteration_number)
