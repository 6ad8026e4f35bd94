"""BASELINE.json configs other than the bench line: iterations / device time / throughput on one B200
(under torchrun: slab-partitioned).  usage: run_configs.py [cfg ...] with cfg in 1 2 4 5; sizes via env."""
import json, os, sys, time
sys.path.insert(0, '.')
import numpy as np
import perphil_b200 as pb

world = int(os.environ.get("WORLD_SIZE", "1"))
comm = None
if world > 1:
    import torch
    from perphil_b200.distributed import SlabComm
    comm = SlabComm.from_env()
    torch.cuda.set_device(comm.device)
rank = comm.rank if comm else 0
want = [a if a in ("4w", "2s") else int(a) for a in sys.argv[1:]] or [1, 2, 4, 5]
out = []


def problem(cells, degree, k2, beta, bc, lengths=None):
    if lengths is not None:
        from perphil_b200.mesh import Mesh
        mesh = Mesh(cells, lengths=lengths, comm=comm)
    else:
        mesh = pb.UnitSquareMesh(*cells, comm=comm) if len(cells) == 2 else pb.UnitCubeMesh(*cells, comm=comm)
    _, V = pb.create_function_spaces(mesh, pressure_deg=degree)
    W = V * V
    prm = pb.DPPParameters(k1=1.0, k2=k2, beta=beta, mu=1.0)
    if bc == "manufactured":
        ex = pb.exact_expressions(mesh, prm)
        g1, g2 = ex[1], ex[3]
    else:  # petsc_profiling.py:685-690 constant mode (config 5: the manufactured data overflows)
        g1, g2 = pb.Constant(1.0), pb.Constant(0.0)
    return W, V, prm, [pb.DirichletBC(W.sub(0), g1, "on_boundary"), pb.DirichletBC(W.sub(1), g2, "on_boundary")]


def run(tag, cells, degree, preset_name, k2=1e-2, beta=1.0, bc="manufactured", nonlinear=False, repeats=2, lengths=None):
    W, V, prm, bcs = problem(cells, degree, k2, beta, bc, lengths)
    fn = pb.solve_dpp_nonlinear if nonlinear else pb.solve_dpp
    preset = getattr(pb, preset_name)
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        sol = fn(W, prm, bcs, solver_parameters=preset)
        wall = time.perf_counter() - t0
        info = pb.last_solve_info()
        dev = info.setup_ms + info.solve_ms
        if best is None or dev < best[0]:
            best = (dev, wall, info)
    dev, wall, info = best
    nglob = int(np.prod([degree * c + 1 for c in cells])) * 2
    rec = {"config": tag, "cells": list(cells), "degree": degree, "preset": preset_name, "n_dof": nglob, "n_gpus": world,
           "iterations": int(sol.iteration_number), "inner_iterations": int(info.inner_iterations),
           "applies": int(info.apply_count), "residual": float(sol.residual_error), "reason": int(info.converged_reason),
           "device_ms": dev, "wall_ms": wall * 1e3, "tts_mdofs": nglob / dev / 1e3,
           "apply_gdofs_equiv": nglob * max(int(info.apply_count), 1) / dev / 1e6}
    if rank == 0:
        print(json.dumps(rec), flush=True)
    out.append(rec)
    pb.release_handles()


N2 = int(os.environ.get("CFG2_N", "128"))
N4 = int(os.environ.get("CFG4_N", "192"))
N5 = int(os.environ.get("CFG5_N", "128"))
if 1 in want and world == 1:
    run("1: 2D 16x16 quad Q1, GMRES(30) (reference: 292 its)", (16, 16), 1, "B200_GMRES_PARAMS")
    run("1: 2D 16x16 quad Q1, Jacobi-CG", (16, 16), 1, "B200_CG_JACOBI_PARAMS")
if 2 in want:
    run(f"2: 3D hex Q1 {N2}^3, CG + block-Jacobi (additive) fieldsplit", (N2,) * 3, 1, "B200_CG_FIELDSPLIT_PARAMS")
    run(f"2: 3D hex Q1 {N2}^3, Jacobi-CG", (N2,) * 3, 1, "B200_CG_JACOBI_PARAMS")
if "2s" in want:   # configs[1]: "scaled 8^3 -> 128^3, CG + block-Jacobi fieldsplit on 1 B200"
    for n in (8, 16, 32, 64, 128):
        run(f"2 sweep: 3D hex Q1 {n}^3, CG + block-Jacobi (additive) fieldsplit", (n,) * 3, 1, "B200_CG_FIELDSPLIT_PARAMS")
        run(f"2 sweep: 3D hex Q1 {n}^3, Jacobi-CG", (n,) * 3, 1, "B200_CG_JACOBI_PARAMS")
if 4 in want:
    run(f"4: 3D hex Q2 {N4}^3, block Picard (scale splitting), Jacobi-CG blocks", (N4,) * 3, 2, "B200_PICARD_SPLIT_PARAMS",
        nonlinear=True, repeats=2 if world > 1 else 1)
    run(f"4: 3D hex Q2 {N4}^3, Jacobi-CG monolithic", (N4,) * 3, 2, "B200_CG_JACOBI_PARAMS", repeats=2 if world > 1 else 1)
if "4w" in want:
    # weak scaling of config 4: a fixed 24 x 192 x 192-cell Q2 slab per GPU (192^3 at 8 GPUs), cubic cells kept by
    # growing the domain length along x with the rank count
    cw = (24 * world, 192, 192)
    run(f"4 weak: 3D hex Q2 {cw[0]}x192x192 (24 cell layers per GPU), block Picard, Jacobi-CG blocks", cw, 2,
        "B200_PICARD_SPLIT_PARAMS", nonlinear=True, repeats=2, lengths=(world / 8.0, 1.0, 1.0))
    run(f"4 weak: 3D hex Q2 {cw[0]}x192x192 (24 cell layers per GPU), Jacobi-CG monolithic", cw, 2,
        "B200_CG_JACOBI_PARAMS", repeats=2, lengths=(world / 8.0, 1.0, 1.0))
if 5 in want:
    run(f"5: 3D hex Q1 {N5}^3 k2=1e-6 beta=1e2, GMRES + multiplicative fieldsplit", (N5,) * 3, 1,
        "B200_GMRES_FIELDSPLIT_PARAMS", k2=1e-6, beta=1e2, bc="const")
    run(f"5: 3D hex Q1 {N5}^3 k2=1e-6 beta=1e2, Jacobi-GMRES(30)", (N5,) * 3, 1, "B200_GMRES_JACOBI_PARAMS", k2=1e-6,
        beta=1e2, bc="const")
if comm is not None:
    comm.barrier()
    comm.destroy()
