"""Per-rank problem sizes of the 8-GPU run: stand-alone kernel times and ms per CG iteration.
  python tools/per_rank_probe.py 32 64        one GPU emulating one slab (no exchange)
  torchrun --nproc-per-node N tools/per_rank_probe.py 32      N slabs of 32 cell layers each (exchange included)
The gap between the two is the exchange cost."""
import os, sys
sys.path.insert(0, '.')
import perphil_b200 as pb

world = int(os.environ.get("WORLD_SIZE", "1"))
comm = None
if world > 1:
    import torch
    from perphil_b200.distributed import SlabComm
    comm = SlabComm.from_env()
    torch.cuda.set_device(comm.device)
tag = " ".join(f"{k}={v}" for k, v in os.environ.items() if k.startswith("DPP_"))
for nx in (int(a) for a in (sys.argv[1:] or ["32", "64", "128", "256"])):
    mesh = pb.UnitCubeMesh(nx * world, 256, 256, comm=comm)
    _, V = pb.create_function_spaces(mesh); W = V * V
    prm = pb.DPPParameters(k1=1.0, k2=1e-2, beta=1.0, mu=1.0)
    _, p1, _, p2 = pb.exact_expressions_3d(mesh, prm)
    bcs = [pb.DirichletBC(W.sub(0), p1, "on_boundary"), pb.DirichletBC(W.sub(1), p2, "on_boundary")]
    best = None
    for _ in range(4):
        try:
            sol = pb.solve_dpp(W, prm, bcs, solver_parameters=pb.B200_CG_JACOBI_PARAMS)
        except Exception as ex:   # DPP_DEBUG_* timing experiments break the numerics on purpose
            print("solve failed:", ex); break
        info = pb.last_solve_info()
        per = info.solve_ms / max(1, sol.iteration_number)
        if comm is not None:
            per = comm.max_float(per)
        best = per if best is None else min(best, per)
    a, u, m = pb.handle_for(W).time_cg_kernels(reps=50, warmup=5)
    if comm is not None:
        a, u, m = comm.max_float(a), comm.max_float(u), comm.max_float(m)
    if comm is None or comm.rank == 0:
        print(f"world={world} nx/rank={nx:4d} [{tag}] fused apply {a*1e3:7.1f} us  r_update {u*1e3:6.1f} us  sum {1e3*(a+u):6.1f} us  "
              f"matvec {m*1e3:6.1f} us  solve: {sol.iteration_number} its, {best*1e3:6.1f} us/iteration", flush=True)
    pb.release_handles()
if comm is not None:
    comm.barrier()
    comm.destroy()
