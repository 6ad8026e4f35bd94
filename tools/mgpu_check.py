"""Multi-rank GPU check (torchrun): slab-partitioned Jacobi-CG equals the single-GPU solve."""
import os, sys
sys.path.insert(0, '.')
import numpy as np, torch
import perphil_b200 as pb
from perphil_b200.distributed import SlabComm
comm = SlabComm.from_env()
torch.cuda.set_device(comm.device)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 32
DEG = int(sys.argv[2]) if len(sys.argv) > 2 else 1
def problem(c):
    mesh = pb.UnitCubeMesh(N, N, N, comm=c); _, V = pb.create_function_spaces(mesh, pressure_deg=DEG); W = V * V
    prm = pb.DPPParameters(k1=1.0, k2=1e-2, beta=1.0, mu=1.0)
    _, p1, _, p2 = pb.exact_expressions_3d(mesh, prm)
    return mesh, V, W, prm, [pb.DirichletBC(W.sub(0), p1, "on_boundary"), pb.DirichletBC(W.sub(1), p2, "on_boundary")]
mesh, V, W, prm, bcs = problem(comm)
for preset in ("B200_CG_JACOBI_PARAMS", "B200_GMRES_JACOBI_PARAMS", "B200_GMRES_FIELDSPLIT_PARAMS", "B200_PICARD_SPLIT_PARAMS"):
    fn = pb.solve_dpp_nonlinear if "PICARD" in preset else pb.solve_dpp
    sol = fn(W, prm, bcs, solver_parameters=getattr(pb, preset))
    first_ms = pb.last_solve_info().solve_ms          # carries the one-time NCCL channel set-up of the first halo
    sol = fn(W, prm, bcs, solver_parameters=getattr(pb, preset))
    info = pb.last_solve_info()
    # single-GPU reference on every rank's own device
    m1, V1, W1, prm1, bcs1 = problem(None)
    ref = fn(W1, prm1, bcs1, solver_parameters=getattr(pb, preset))
    slab = mesh.slab
    pn = (DEG * N + 1) ** 2
    lo, hi = DEG * slab.local_plane_lo * pn, (DEG * (slab.local_plane_hi - 1) + 1) * pn
    err = 0.0
    for f in range(2):
        a = sol.solution.sub(f).dat.data; b = ref.solution.sub(f).dat.data[lo:hi]
        err = max(err, float(np.linalg.norm(a - b) / np.linalg.norm(b)))
    print(f"rank {comm.rank}/{comm.size} ipc={pb.handle_for(W).info().peer_memory} {preset}: its {sol.iteration_number} vs {ref.iteration_number}, rel err {err:.2e}, solve {info.solve_ms:.2f} ms (first call {first_ms:.2f} ms)", flush=True)
    assert abs(sol.iteration_number - ref.iteration_number) <= (0 if "CG_JACOBI" in preset else 2) and err < 1e-7
    pb.release_handles()
comm.barrier()
if comm.rank == 0: print("MGPU OK")
comm.destroy()
