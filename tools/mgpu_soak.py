"""Soak test of the cross-GPU protocol (peer-memory halo push + mailbox all-reduce): the same slab-partitioned
Jacobi-CG solve repeated many times; every solve on every rank must report the SAME iteration count and
bit-identical residual norm (all ranks hold bit-identical reduction results by construction: every rank adds the
per-rank partial sums in rank order), and a sequence-counter slip, a lost halo plane or a torn mailbox word would
show up as a different count, a different residual, a NaN or the 30 s mailbox timeout.
  torchrun --nproc-per-node N tools/mgpu_soak.py [cells per direction = 32] [solves = 1000]"""
import struct, sys, time
sys.path.insert(0, '.')
import numpy as np, torch
import perphil_b200 as pb
from perphil_b200.distributed import SlabComm

comm = SlabComm.from_env()
torch.cuda.set_device(comm.device)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 32
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
mesh = pb.UnitCubeMesh(N, N, N, comm=comm)
_, V = pb.create_function_spaces(mesh)
W = V * V
prm = pb.DPPParameters(k1=1.0, k2=1e-2, beta=1.0, mu=1.0)
_, p1, _, p2 = pb.exact_expressions_3d(mesh, prm)
bcs = [pb.DirichletBC(W.sub(0), p1, "on_boundary"), pb.DirichletBC(W.sub(1), p2, "on_boundary")]
first = None
t0 = time.perf_counter()
for k in range(reps):
    sol = pb.solve_dpp(W, prm, bcs, solver_parameters=pb.B200_CG_JACOBI_PARAMS)
    sig = (sol.iteration_number, struct.pack("d", sol.residual_error), struct.pack("d", float(np.sum(sol.solution.sub(0).dat.data))))
    if first is None:
        first = sig
    assert sig == first, f"rank {comm.rank}: solve {k} differs: {sig} vs {first}"
dt = time.perf_counter() - t0
sigs = comm.all_gather_bytes(struct.pack("i", first[0]) + first[1])
assert len(set(sigs)) == 1, f"ranks disagree: {sigs}"
info = pb.handle_for(W).info()
if comm.rank == 0:
    print(f"SOAK OK: {reps} solves of {N}^3 on {comm.size} ranks (peer_memory={info.peer_memory}), every solve {first[0]} iterations, "
          f"residual bits {first[1].hex()} on every rank, {1e3 * dt / reps:.2f} ms per solve_dpp call", flush=True)
comm.barrier()
comm.destroy()
