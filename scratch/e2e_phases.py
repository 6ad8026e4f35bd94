import sys, time
sys.path.insert(0, '.')
import numpy as np
import bench, perphil_b200 as pb
from perphil_b200 import solver as S
from perphil_b200.provider import bc_data
from perphil_b200.backend import PINNED
N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
W, V, prm, bcs = bench.build_problem(N)
preset = pb.B200_CG_JACOBI_PARAMS
for i in range(3):
    t0 = time.perf_counter(); sol = pb.solve_dpp(W, prm, bcs, solver_parameters=preset); print("solve_dpp %.1f ms" % ((time.perf_counter() - t0) * 1e3), pb.last_solve_info().solve_ms)
h = pb.handle_for(W)
for i in range(3):
    T = [time.perf_counter()]
    h.set_params(1.0, 1e-2, 1.0, 1.0); T.append(time.perf_counter())
    got = {f: (n, v) for f, n, v in bc_data(W, bcs)}; T.append(time.perf_counter())
    for f in (0, 1):
        h.set_dirichlet(f, *got[f])
    T.append(time.perf_counter())
    opt = S.options_from_petsc(h, preset); T.append(time.perf_counter())
    buf = PINNED.take(2 * h.n_nodes); T.append(time.perf_counter())
    _, info = h.solve(opt, want_solution=True, out=buf); T.append(time.perf_counter())
    print("phases ms: params %.2f bc_data %.2f set_dirichlet %.2f opts %.2f pinned %.2f solve %.2f (device %.2f + %.2f)" % (
        *[(T[k + 1] - T[k]) * 1e3 for k in range(6)], info.setup_ms, info.solve_ms))
    _, info = h.solve(opt, want_solution=False)
    t0 = time.perf_counter(); _, info = h.solve(opt, want_solution=False); print("  solve no D2H %.2f ms wall, device %.2f+%.2f" % ((time.perf_counter() - t0) * 1e3, info.setup_ms, info.solve_ms))
