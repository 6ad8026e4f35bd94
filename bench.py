"""bench.py -- BASELINE.json metric: DPP solve GDoF/s + matvec HBM GB/s (3D hex Q1).

Step = one full solve of the linear DPP pressure system (lifting + Jacobi-CG to rtol 1e-8) on the
3-D hex Q1 unit cube with manufactured Dirichlet data (BASELINE configs[2], 256^3, matrix-free).
value  = N_dof * iterations / device time  (iteration-normalised solve throughput),
         inputs resident in HBM, timed with CUDA events on the library's stream.
e2e    = same metric through perphil_b200.solve_dpp(W, params, bcs, preset) with host buffers
         (Dirichlet data H2D + solution D2H inside the timed region, wall clock around the call).
roofline = the matrix-free apply kernel (dominant kernel of every Krylov iteration's operator part).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.gpu)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); smax.append(float(r[2]))
                for nm, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def build_problem(N, comm=None):
    import perphil_b200 as pb

    mesh = pb.UnitCubeMesh(N, N, N, comm=comm)
    _, V = pb.create_function_spaces(mesh)
    W = V * V
    prm = pb.DPPParameters(k1=1.0, k2=1e-2, beta=1.0, mu=1.0)   # iterative_bench.py:131
    _, p1, _, p2 = pb.exact_expressions_3d(mesh, prm)
    bcs = [pb.DirichletBC(W.sub(0), p1, "on_boundary"), pb.DirichletBC(W.sub(1), p2, "on_boundary")]
    return W, V, prm, bcs


def cpu_baseline(N, repeats=1):
    """Oracle (numpy/scipy port of the PETSc path) Jacobi-CG on a bounded sample of the workload."""
    from oracle import dpp_oracle as orc

    t0 = time.perf_counter()
    osys = orc.build_system(orc.structured_mesh((N, N, N), 1), orc.Params(k1=1.0, k2=1e-2, beta=1.0, mu=1.0),
                            "manufactured", route="kron")
    t_asm = time.perf_counter() - t0
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        sol = orc.solve_dpp_oracle(osys, "cg", "jacobi")
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    ndof = osys.n_dof
    return {"value": ndof * sol.iteration_number / best / 1e9, "unit": "GDoF/s", "cores": 1, "kind": "port",
            "sample": f"{N}^3 hex Q1 ({ndof} DoF), scipy CSR SpMV + numpy BLAS-1 Jacobi-CG, {sol.iteration_number} its "
                      f"in {best:.2f} s (assembly {t_asm:.1f} s not included)",
            "iterations": int(sol.iteration_number), "seconds": best}


def run_reference(args):
    """--impl reference: the CPU restatement of the PETSc path, timed on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import dpp_oracle as orc

    N = args.ref_size
    osys = orc.build_system(orc.structured_mesh((N, N, N), 1), orc.Params(k1=1.0, k2=1e-2, beta=1.0, mu=1.0),
                            "manufactured", route="kron")
    for _ in range(args.warmup):
        sol = orc.solve_dpp_oracle(osys, "cg", "jacobi")
    t0 = time.perf_counter()
    for _ in range(args.steps):
        sol = orc.solve_dpp_oracle(osys, "cg", "jacobi")
    dt = (time.perf_counter() - t0) / max(args.steps, 1)
    val = osys.n_dof * sol.iteration_number / dt / 1e9
    line = {
        "impl": "reference", "metric": "dpp_solve_gdofs", "value": val, "unit": "GDoF/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"3D hex Q1 {args.size}^3 monolithic DPP Jacobi-CG rtol 1e-8 (bounded sample {N}^3)"},
        "cpu_baseline": {"value": val, "unit": "GDoF/s", "cores": 1, "kind": "port",
                         "sample": f"{N}^3 hex Q1, {sol.iteration_number} its per step"},
        "e2e": {"value": val, "unit": "GDoF/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--size", type=int, default=256, help="cells per direction (BASELINE configs[2]: 256)")
    ap.add_argument("--ref-size", type=int, default=64)
    ap.add_argument("--cpu-size", type=int, default=64)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch

    import perphil_b200 as pb
    from perphil_b200.solver import options_from_petsc

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    comm = None
    if world > 1:
        from perphil_b200.distributed import SlabComm

        comm = SlabComm.from_env()
    torch.cuda.set_device(local_rank)
    N = args.size
    W, V, prm, bcs = build_problem(N, comm)
    h = pb.handle_for(W)
    info = h.info()
    n_nodes_global = (N + 1) ** 3
    ndof = 2 * n_nodes_global
    preset = pb.B200_CG_JACOBI_PARAMS

    # ---- e2e warm-up through the public API (also uploads BCs / params)
    for _ in range(max(args.warmup, 3)):
        sol = pb.solve_dpp(W, prm, bcs, solver_parameters=preset)
    its = sol.iteration_number
    opt = options_from_petsc(h, preset)

    def barrier():
        if comm is not None:
            comm.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timed region
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = h.launch_count()
    barrier()
    dev_ms = []
    t0 = time.perf_counter()
    for _ in range(args.steps):
        _, si = h.solve(opt, want_solution=False)
        dev_ms.append(si.setup_ms + si.solve_ms)
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3 / args.steps
    launches = h.launch_count() - l0
    ms = float(np.mean(dev_ms))
    if comm is not None:
        ms = comm.max_float(ms)
        wall_ms = comm.max_float(wall_ms)

    # ---- end-to-end through solve_dpp with host buffers
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        sol = pb.solve_dpp(W, prm, bcs, solver_parameters=preset)
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / args.steps
    if comm is not None:
        e2e_ms = comm.max_float(e2e_ms)
    clocks = sampler.stop() if rank == 0 else None
    nb = int(V.boundary_nodes.size)
    h2d = 2 * nb * (4 + 8) * world if comm is None else comm.sum_int(2 * nb * (4 + 8))
    d2h = 2 * info.n_nodes * 8 if comm is None else comm.sum_int(2 * info.n_nodes * 8)

    # ---- apply roofline (dominant operator kernel), CUDA events inside the library
    apply_ms = h.time_apply(reps=20, warmup=3, with_dot=True)
    if comm is not None:
        apply_ms = comm.max_float(apply_ms)
    peak, peak_kind = measured_peaks()
    structured = info.kernel_family == 1
    alg_bytes = 34 * n_nodes_global if structured else 58 * n_nodes_global + 32 * N ** 3
    achieved = alg_bytes / (apply_ms * 1e-3) / 1e9 / world
    iter_bytes = alg_bytes + 88 * ndof          # fused Jacobi-PCG minimum
    solve_gbs = iter_bytes * its / (np.mean([m for m in dev_ms]) * 1e-3) / 1e9 / world

    if rank != 0:
        return
    line = {
        "metric": "dpp_solve_gdofs", "value": ndof * its / (ms * 1e-3) / 1e9, "unit": "GDoF/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"3D hex Q1 {N}^3 monolithic DPP, matrix-free Jacobi-CG rtol 1e-8, manufactured BCs "
                               f"(BASELINE configs[2]); {ndof} DoF; inputs 10x L2, no flush needed",
                   "preset": "B200_CG_JACOBI_PARAMS", "iterations": its, "parallelism": f"slab x{world}",
                   "kernel_family": "structured" if structured else "general"},
        "iterations": its, "residual_error": sol.residual_error, "wall_ms_per_step": wall_ms,
        "tts_mdofs": ndof / (ms * 1e-3) / 1e6,
        "matvec_gdofs": ndof / (apply_ms * 1e-3) / 1e9, "matvec_ms": apply_ms,
        "e2e": {"value": ndof * its / (e2e_ms * 1e-3) / 1e9, "unit": "GDoF/s", "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": "k_apply_uniform<2> + k_fix_rows (matrix-free apply, fused <p,Ap>)" if structured else "k_general",
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                     "peak_kind": peak_kind, "bytes_model": "34 B/node structured" if structured else "58 B/node + 32 B/cell",
                     "algorithmic_bytes": alg_bytes},
        "solve_roofline": {"achieved": solve_gbs, "peak": peak, "unit": "GB/s", "frac": solve_gbs / peak,
                           "bytes_per_iteration": iter_bytes},
    }
    if not args.no_cpu:
        line["cpu_baseline"] = cpu_baseline(args.cpu_size)
    print(json.dumps(line))


if __name__ == "__main__":
    main()
