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


# dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full captures
# (profiles/): filled in after each profiling pass, None until a capture of the current kernel exists.
TRAFFIC_NCU = {"fused_apply": 1648143872, "r_update": 819625216,   # profiles/r01_cg_kernels_final2.md (256^3)
               "fused_apply_deferred": 1100252880,                 # profiles/r02_cg_kernels.md (256^3)
               "fused_apply_no_w": 850131968}                      # profiles/r02_cg_kernels_b.md (256^3)


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


def host_threads():
    """Threads the CPU legs use: every core this process may run on -- set explicitly, because launchers
    (torch.distributed.run) export OMP_NUM_THREADS=1 and would otherwise silently change the baseline."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_model():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def cpu_baseline(N, repeats=1, single_thread_size=64):
    """C/OpenMP restatement of the reference's CPU path (assembled AIJ + SeqAIJ MatMult + KSPCG/PCJACOBI,
    oracle/dpp_oracle_c.c) on a bounded sample of the workload: all host threads, plus the reference's own
    protocol -- one thread (notebooks/petsc-profiling-time-benchmarks-3d.py:11 runs OMP_NUM_THREADS=1) -- on a
    smaller sample so the default bench run stays within minutes."""
    from oracle import c_oracle as co

    nthr = host_threads()
    co.set_num_threads(nthr)
    t0 = time.perf_counter()
    csys = co.manufactured_system((N, N, N), 1)
    t_asm = time.perf_counter() - t0
    best, res = None, None
    for _ in range(repeats):
        t0 = time.perf_counter()
        res = csys.cg("jacobi", want_solution=False)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    ndof, nnz = csys.n_dof, csys.nnz
    csys.close()
    one = None
    if single_thread_size:
        M = single_thread_size
        c1 = co.manufactured_system((M, M, M), 1)
        co.set_num_threads(1)
        t0 = time.perf_counter()
        r1 = c1.cg("jacobi", want_solution=False)
        d1 = time.perf_counter() - t0
        co.set_num_threads(nthr)
        one = {"value": c1.n_dof * r1.iteration_number / d1 / 1e9, "unit": "GDoF/s", "cores": 1,
               "sample": f"{M}^3 hex Q1 ({c1.n_dof} DoF), {r1.iteration_number} its in {d1:.2f} s, OMP threads 1 "
                         "(the reference's profiling protocol)"}
        c1.close()
    return {"value": ndof * res.iteration_number / best / 1e9, "unit": "GDoF/s", "cores": nthr,
            "kind": "port",
            "sample": f"{N}^3 hex Q1 ({ndof} DoF, nnz {nnz}), C/OpenMP CSR SpMV + Jacobi-CG (PETSc KSPCG semantics), "
                      f"{res.iteration_number} its in {best:.2f} s (assembly {t_asm:.1f} s not included; "
                      f"SpMV {100 * res.spmv_seconds / best:.0f}% of the solve)",
            "iterations": int(res.iteration_number), "seconds": best, "host_cpus": os.cpu_count(),
            "cpu_model": cpu_model(), "single_thread": one}


def general_kernel_entry(N, peak, peak_kind):
    """Second roofline entry: the element-based matrix-free apply (csrc/apply_cells.cu) on a GENERAL mesh -- the
    N^3 hex Q1 lattice with perturbed interior vertices (non-affine cells), randomly permuted node numbering and
    cell order -- against the unstructured bytes model of SURVEY 8(d): 58 B/node + 32 B/cell.  The kernel is bound
    by the fp64 FMA pipe, not by HBM (~1260 fp64 instructions per distorted cell), so the fraction
    of the fp64 issue rate (64 lanes/clk/SM at the sampled SM clock) is reported next to the HBM fraction."""
    from perphil_b200.backend import DppHandle
    from tools.general_mesh import shuffled_distorted_hex

    t0 = time.perf_counter()
    cnm, X, bn = shuffled_distorted_hex(N, 0.25, seed=1)
    h = DppHandle.from_mesh_arrays(3, 1, cnm, X, X, cnm, n_nodes=X.shape[0])
    h.set_params(1.0, 1e-2, 1.0, 1.0)
    g = np.random.default_rng(2).standard_normal(bn.size)
    h.set_dirichlet(0, bn, g)
    h.set_dirichlet(1, bn, -g)
    h.time_apply(reps=1, warmup=0, with_dot=True)          # one-time cell-block setup (host) + first launch
    setup_s = time.perf_counter() - t0
    ms = h.time_apply(reps=20, warmup=3, with_dot=True)
    u, info = h.solve()                                    # Jacobi-CG (unfused kernel sequence) on the same mesh
    nn, ncell = X.shape[0], cnm.shape[0]
    bytes_model = 58 * nn + 32 * ncell
    fam = h.info().kernel_family
    sm = h.info().sm_count
    h.close()
    fp64_rate = sm * 64 * 1.9e9
    return {"bound": "fp64", "kernel": "k_cells_stage + k_cells_q1<2> + k_cells_gather<2> (general hex Q1, distorted + shuffled)",
            "workload": f"{N}^3 hex Q1, vertices perturbed by 0.25 h, random node/cell numbering ({2 * nn} DoF)",
            "kernel_family": "general" if fam == 0 else "structured",
            "ms_per_apply": ms, "matvec_gdofs": 2 * nn / ms / 1e6,
            "achieved": bytes_model / ms / 1e6, "peak": peak, "unit": "GB/s", "frac": bytes_model / ms / 1e6 / peak,
            "peak_kind": peak_kind, "bytes_model": "58 B/node + 32 B/cell (SURVEY 8d, unstructured data model)",
            "algorithmic_bytes": bytes_model, "traffic": None,
            "fp64_instructions_model": 1260 * ncell, "fp64_pipe_frac_at_1.9GHz": 1260 * ncell / (ms * 1e-3) / fp64_rate,
            "solve": {"ksp": "cg", "pc": "jacobi", "iterations": int(info.iterations), "solve_ms": info.solve_ms,
                      "gdofs": 2 * nn * info.iterations / info.solve_ms / 1e6},
            "first_call_s": setup_s}


def assembly_entry(N, peak):
    """CSR assembly of the 2x2-block matrix (BASELINE configs[1] top size by default): symbolic and numeric phase
    separately (CUDA events), numeric against the nnz * 12 B model (8 B value written + 4 B column index)."""
    import perphil_b200 as pb

    mesh = pb.UnitCubeMesh(N, N, N)
    _, V = pb.create_function_spaces(mesh)
    W = V * V
    prm = pb.DPPParameters(k1=1.0, k2=1e-2, beta=1.0, mu=1.0)
    _, p1, _, p2 = pb.exact_expressions_3d(mesh, prm)
    bcs = [pb.DirichletBC(W.sub(0), p1, "on_boundary"), pb.DirichletBC(W.sub(1), p2, "on_boundary")]
    from perphil_b200.solver import configure_handle

    h = pb.handle_for(W)
    configure_handle(h, W, prm, bcs)
    sym_ms, num_ms, nnz = h.time_assembly(reps=5)
    spmv_ms = h.time_apply(reps=10, warmup=2, assembled=True)
    ndof = 2 * h.n_nodes
    out = {"workload": f"{N}^3 hex Q1, nnz {nnz}", "symbolic_ms": sym_ms, "numeric_ms": num_ms,
           "numeric_gbs": 12 * nnz / num_ms / 1e6, "numeric_frac": 12 * nnz / num_ms / 1e6 / peak,
           "bytes_model": "nnz * 12 B", "spmv_ms": spmv_ms, "spmv_gbs": (12 * nnz + 24 * ndof) / spmv_ms / 1e6,
           "spmv_frac": (12 * nnz + 24 * ndof) / spmv_ms / 1e6 / peak}
    pb.release_handles()
    return out


def run_reference(args):
    """--impl reference: the reference's CPU path for this metric.  perphil itself is Python glue over
    Firedrake/PETSc, which cannot be installed here or on the GPU box, so the arm times the
    C/OpenMP restatement of that path (oracle/dpp_oracle_c.c) with all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import c_oracle as co

    N = args.cpu_size
    nthr = host_threads()
    co.set_num_threads(nthr)   # identical at every N: the launcher's OMP_NUM_THREADS does not decide the baseline
    csys = co.manufactured_system((N, N, N), 1)
    res = None
    for _ in range(args.warmup):
        res = csys.cg("jacobi", want_solution=False)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        res = csys.cg("jacobi", want_solution=False)
    dt = (time.perf_counter() - t0) / max(args.steps, 1)
    val = csys.n_dof * res.iteration_number / dt / 1e9
    sample = (f"{N}^3 hex Q1 ({csys.n_dof} DoF), assembled CSR SpMV + Jacobi-CG, {res.iteration_number} its per step, "
              f"{co.num_threads()} OpenMP threads on {cpu_model()}")
    line = {
        "impl": "reference", "metric": "dpp_solve_gdofs", "value": val, "unit": "GDoF/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"3D hex Q1 {args.size}^3 monolithic DPP Jacobi-CG rtol 1e-8 (bounded sample {N}^3)"},
        "iterations": int(res.iteration_number),
        "cpu_baseline": {"value": val, "unit": "GDoF/s", "cores": co.num_threads(), "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "GDoF/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def run_other_config(args):
    """BASELINE configs[3] (Q2 192^3 block Picard, 8-GPU weak scaling) and configs[4] (high-contrast 128^3 GMRES +
    fieldsplit) as bench lines of the same schema.  value = N_dof x Krylov iterations (inner Jacobi-CG iterations of
    the block solves) / device time: the iteration-normalised throughput of SURVEY 8(d)."""
    import torch

    import perphil_b200 as pb
    from perphil_b200.mesh import Mesh

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    comm = None
    if world > 1:
        from perphil_b200.distributed import SlabComm

        comm = SlabComm.from_env()
    torch.cuda.set_device(local_rank)
    if args.config == 4:
        cells, degree = (24 * world, 192, 192), 2
        mesh = Mesh(cells, lengths=(world / 8.0, 1.0, 1.0), comm=comm)     # cubic cells at every N
        prm = pb.DPPParameters(k1=1.0, k2=1e-2, beta=1.0, mu=1.0)
        _, g1, _, g2 = pb.exact_expressions_3d(mesh, prm)
        preset, name, fn, scaling = pb.B200_PICARD_SPLIT_PARAMS, "B200_PICARD_SPLIT_PARAMS", pb.solve_dpp_nonlinear, "weak"
        workload = (f"3D hex Q2 {cells[0]}x192x192 cells (24 cell layers per GPU; 192^3 at 8 GPUs), block Picard "
                    "(scale splitting) with Jacobi-CG block solves rtol 1e-10, manufactured BCs (BASELINE configs[3])")
    else:
        if world > 1:
            raise SystemExit("--config 5 is a single-GPU configuration")
        cells, degree = (128, 128, 128), 1
        mesh = pb.UnitCubeMesh(*cells)
        prm = pb.DPPParameters(k1=1.0, k2=1e-6, beta=1e2, mu=1.0)
        g1, g2 = pb.Constant(1.0), pb.Constant(0.0)
        preset, name, fn, scaling = pb.B200_GMRES_FIELDSPLIT_PARAMS, "B200_GMRES_FIELDSPLIT_PARAMS", pb.solve_dpp, "strong"
        workload = ("3D hex Q1 128^3, k2 = 1e-6, beta = 1e2, constant BCs p1 = 1, p2 = 0, GMRES(30) + multiplicative "
                    "fieldsplit with Jacobi-CG block solves rtol 1e-10 (BASELINE configs[4])")
    _, V = pb.create_function_spaces(mesh, pressure_deg=degree)
    W = V * V
    bcs = [pb.DirichletBC(W.sub(0), g1, "on_boundary"), pb.DirichletBC(W.sub(1), g2, "on_boundary")]
    ndof = 2 * int(np.prod([degree * c + 1 for c in cells]))

    def barrier():
        if comm is not None:
            comm.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 1) if args.config == 4 else max(args.warmup, 3)):
        sol = fn(W, prm, bcs, solver_parameters=preset)
    h = pb.handle_for(W)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = h.launch_count()
    barrier()
    t0 = time.perf_counter()
    dev = []
    for _ in range(args.steps):
        sol = fn(W, prm, bcs, solver_parameters=preset)     # host buffers in, host buffers out: this IS the e2e call
        info = pb.last_solve_info()
        dev.append(info.setup_ms + info.solve_ms)
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / args.steps
    launches = h.launch_count() - l0
    ms = float(np.mean(dev))
    if comm is not None:
        ms, e2e_ms = comm.max_float(ms), comm.max_float(e2e_ms)
    clocks = sampler.stop() if rank == 0 else None
    its = max(int(info.inner_iterations), int(sol.iteration_number))
    peak, peak_kind = measured_peaks()
    nodes_global = ndof // 2
    fused_q2 = args.config == 4 and h.fused_cg_supported()
    if fused_q2:
        # the block solves of the Picard iteration run the two-kernel Jacobi-CG iteration on one field; its apply
        # kernel (r, p_old read; p, w written: 32 B per node) is the dominant launch
        apply_ms, _ = h.time_cg_block_kernels(0, reps=10, warmup=3)
        apply_bytes, bytes_model = 32 * nodes_global, "32 B/node: r, p_old read, p, w written (one field)"
    else:
        apply_ms = h.time_apply(reps=10, warmup=3, with_dot=True)
        apply_bytes, bytes_model = 34 * nodes_global, "34 B/node structured (x read, y written, 1 B Dirichlet)"
    if comm is not None:
        apply_ms = comm.max_float(apply_ms)
    nb = int(V.boundary_nodes.size)
    h2d = 2 * nb * 12 if comm is None else comm.sum_int(2 * nb * 12)
    d2h = 2 * h.n_nodes * 8 if comm is None else comm.sum_int(2 * h.n_nodes * 8)
    if rank != 0:
        return
    kern = ("k_cg_fused_apply_q2<1> (Q2 block iteration kernel)" if fused_q2 else
            "k_apply_q2u<2> (uniform-grid Q2 apply)" if degree == 2 else "k_apply_uniform<2> (Q1 apply, caller's layout)")
    line = {
        "metric": "dpp_solve_gdofs", "value": ndof * its / (ms * 1e-3) / 1e9, "unit": "GDoF/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": scaling,
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload + f"; {ndof} DoF; inputs larger than L2", "preset": name,
                   "outer_iterations": int(sol.iteration_number), "inner_iterations": int(info.inner_iterations),
                   "parallelism": f"slab x{world}"},
        "iterations": its, "residual_error": sol.residual_error, "tts_mdofs": ndof / (ms * 1e-3) / 1e6,
        "e2e": {"value": ndof * its / (e2e_ms * 1e-3) / 1e9, "unit": "GDoF/s", "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
        "gpu_launches": int(launches), "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": kern, "achieved": apply_bytes / (apply_ms * 1e-3) / 1e9 / world, "peak": peak,
                     "unit": "GB/s", "frac": apply_bytes / (apply_ms * 1e-3) / 1e9 / world / peak, "traffic": None,
                     "peak_kind": peak_kind, "bytes_model": bytes_model,
                     "algorithmic_bytes": apply_bytes // world, "algorithmic_bytes_scope": "per rank, per launch",
                     "ms_per_launch": apply_ms},
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--size", type=int, default=256, help="cells per direction (BASELINE configs[2]: 256)")
    ap.add_argument("--cpu-size", type=int, default=96,
                    help="cells per direction of the bounded CPU sample (cpu_baseline AND --impl reference)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--general-size", type=int, default=128,
                    help="cells per direction of the shuffled + distorted mesh of the general-kernel entry (0: skip)")
    ap.add_argument("--assembly-size", type=int, default=128, help="CSR assembly timing entry (0: skip)")
    ap.add_argument("--config", type=int, default=3, choices=[3, 4, 5],
                    help="BASELINE configs index + 1: 3 = 256^3 Q1 Jacobi-CG (the metric, default); 4 = Q2 block Picard "
                         "weak scaling (24 x 192 x 192 cells per GPU: 192^3 at 8 GPUs); 5 = high-contrast 128^3 GMRES "
                         "fieldsplit (1 GPU)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.config != 3:
        return run_other_config(args)

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
    t0 = time.perf_counter()
    h = pb.handle_for(W)          # dpp_create: mesh upload, structured detection, (N > 1) NCCL + IPC wiring
    create_ms = (time.perf_counter() - t0) * 1e3
    info = h.info()
    n_nodes_global = (N + 1) ** 3
    ndof = 2 * n_nodes_global
    preset = pb.B200_CG_JACOBI_PARAMS

    # ---- e2e warm-up through the public API (also uploads BCs / params); the very first call carries the
    # one-time costs (BC upload + classification, work-vector allocation, tensor maps, CUDA-graph capture)
    t0 = time.perf_counter()
    sol = pb.solve_dpp(W, prm, bcs, solver_parameters=preset)
    first_solve_ms = (time.perf_counter() - t0) * 1e3
    for _ in range(max(args.warmup, 3) - 1):
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
    dev_ms, setup_ms = [], []
    t0 = time.perf_counter()
    for _ in range(args.steps):
        _, si = h.solve(opt, want_solution=False)
        dev_ms.append(si.setup_ms + si.solve_ms)
        setup_ms.append(si.setup_ms)
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

    # ---- roofline of the dominant kernel, CUDA events on the library's stream (each kernel timed alone)
    apply_ms = h.time_apply(reps=20, warmup=3, with_dot=True)      # plain matrix-free apply (MatMult)
    peak, peak_kind = measured_peaks()
    structured = info.kernel_family == 1
    apply_bytes = 34 * n_nodes_global if structured else 58 * n_nodes_global + 32 * N ** 3
    fused = structured
    try:
        fa_ms, fu_ms, mv_ms = h.time_cg_kernels(reps=20, warmup=3)
    except Exception:
        fused, fa_ms, fu_ms, mv_ms = False, None, None, None
    if comm is not None:
        apply_ms = comm.max_float(apply_ms)
        if fused:
            fa_ms, fu_ms, mv_ms = comm.max_float(fa_ms), comm.max_float(fu_ms), comm.max_float(mv_ms)
    # deferred x update (default wherever the reduction epilogues run inside the kernels): the apply kernel moves
    # r, p_old in and p, A p out; x is brought up to date by the r-update kernel every 15th iteration from a ring of
    # 16 direction buffers ((15 + 2) / 15 passes per iteration instead of 2)
    deferred = fused and os.environ.get("DPP_NO_DEFER_X") is None and (world == 1 or (info.peer_memory & 1))
    # residual update that recomputes A p (k_cg_fused_apply<2, 2>; default for full-boundary Dirichlet sets wherever x
    # is deferred and, on slabs, the residual halo goes through peer memory): the iteration kernel stores no w
    stencil_rupd = (deferred and os.environ.get("DPP_NO_STENCIL_RUPD") is None
                    and (world == 1 or (info.peer_memory & 2)))
    traffic_key = None
    if fused and stencil_rupd:
        dom_name = "k_cg_fused_apply<2, 1> (deferred x, no w store): p update + matrix-free apply + <p,Ap>"
        dom_bytes = (3 * 8 + 1) * ndof
        dom_ms = fa_ms
        iter_bytes = dom_bytes + 3 * 8 * ndof + (17 * 8 * ndof) // 15
        bytes_model = ("25 B/DoF: r, p_old in; p out; 1 B/DoF Dirichlet rows (A p is recomputed by the residual update "
                       "k_cg_fused_apply<2, 2>: p, r in; r out = 24 B/DoF; x: 17/15 passes per iteration there)")
        traffic_key = "fused_apply_no_w"
    elif fused and deferred:
        dom_name = "k_cg_fused_apply<2> (deferred x): p update + matrix-free apply + <p,Ap>"
        dom_bytes = (4 * 8 + 1) * ndof
        dom_ms = fa_ms
        iter_bytes = dom_bytes + 3 * 8 * ndof + (17 * 8 * ndof) // 15
        bytes_model = "33 B/DoF: r, p_old in; p, Ap out; 1 B/DoF Dirichlet rows (x: 17/15 passes per iteration in k_cg_r_update)"
        traffic_key = "fused_apply_deferred"
    elif fused:
        # k_cg_fused_apply: reads r, p_old, x and writes p, A p, x (6 passes of 8 B per DoF) + the 1 B/DoF
        # Dirichlet information (row fix-up list); k_cg_r_update: reads r, A p, writes r (3 passes)
        dom_name = "k_cg_fused_apply<2> (+k_fix_rows): p/x update + matrix-free apply + <p,Ap>"
        dom_bytes = (6 * 8 + 1) * ndof
        dom_ms = fa_ms
        iter_bytes = dom_bytes + 3 * 8 * ndof
        bytes_model = "49 B/DoF: r, p_old, x in; p, Ap, x out; 1 B/DoF Dirichlet rows"
        traffic_key = "fused_apply"
    else:
        dom_name = "k_apply (matrix-free apply, fused <p,Ap>)"
        dom_bytes, dom_ms = apply_bytes, apply_ms
        iter_bytes = apply_bytes + 88 * ndof      # unfused-kernel minimum
        bytes_model = "58 B/node + 32 B/cell"
    # per-rank figures: at N > 1 every rank launches the kernel on its slab (1/N of the DoF); the time is the max
    # over ranks, so bytes / world / time is what ONE GPU achieves against ITS HBM peak
    achieved = dom_bytes / (dom_ms * 1e-3) / 1e9 / world
    solve_gbs = iter_bytes * its / (np.mean([m for m in dev_ms]) * 1e-3) / 1e9 / world

    if rank != 0:
        return
    line = {
        "metric": "dpp_solve_gdofs", "value": ndof * its / (ms * 1e-3) / 1e9, "unit": "GDoF/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"3D hex Q1 {N}^3 monolithic DPP, matrix-free Jacobi-CG rtol 1e-8, manufactured BCs "
                               f"(BASELINE configs[2]); {ndof} DoF; inputs 10x L2, no flush needed",
                   "preset": "B200_CG_JACOBI_PARAMS", "iterations": its, "x_update": "deferred (ring of 16)" if deferred else "per iteration", "residual_update": "recomputes A p (no w vector)" if (fused and stencil_rupd) else "reads the stored w", "parallelism": f"slab x{world}" + (" peer-memory halo + mailbox allreduce" if h.info().peer_memory == 3 else (" NCCL" if world > 1 else "")),
                   "kernel_family": "structured" if structured else "general"},
        "iterations": its, "residual_error": sol.residual_error, "wall_ms_per_step": wall_ms,
        "lifting_setup_ms": float(np.mean(setup_ms)), "krylov_us_per_iteration": (ms - float(np.mean(setup_ms))) / its * 1e3,
        "tts_mdofs": ndof / (ms * 1e-3) / 1e6,
        "matvec_gdofs": ndof / ((mv_ms if fused else apply_ms) * 1e-3) / 1e9, "matvec_ms": mv_ms if fused else apply_ms,
        "matvec_gbs": apply_bytes / ((mv_ms if fused else apply_ms) * 1e-3) / 1e9 / world,
        "matvec_roofline_frac": apply_bytes / ((mv_ms if fused else apply_ms) * 1e-3) / 1e9 / world / peak,
        "e2e": {"value": ndof * its / (e2e_ms * 1e-3) / 1e9, "unit": "GDoF/s", "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
        "gpu_launches": int(launches),
        "first_call_ms": {"dpp_create": create_ms, "first_solve_dpp": first_solve_ms,
                          "note": "one-time costs outside every other number: mesh upload + structured detection; "
                                  "first solve = BC upload/classification, work vectors, tensor maps, graph capture"},
        "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": dom_name, "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak,
                     # ncu dram bytes of ONE launch: captured on one GPU at 256^3 only (profiles/); no capture
                     # exists for the per-rank slab kernels, so N > 1 and other sizes report null
                     "traffic": TRAFFIC_NCU.get(traffic_key) if (fused and world == 1 and N == 256) else None,
                     "peak_kind": peak_kind, "bytes_model": bytes_model,
                     "algorithmic_bytes": dom_bytes // world, "algorithmic_bytes_scope": "per rank, per launch",
                     "ms_per_launch": dom_ms},
        "kernels": {"cg_fused_apply_ms": fa_ms, "cg_r_update_ms": fu_ms, "tma_matvec_ms": mv_ms, "plain_apply_ms": apply_ms,
                    "plain_apply_gbs": apply_bytes / (apply_ms * 1e-3) / 1e9 / world,
                    "plain_apply_frac": apply_bytes / (apply_ms * 1e-3) / 1e9 / world / peak,
                    "plain_apply_bytes_model": "34 B/node structured" if structured else "58 B/node + 32 B/cell",
                    # the residual-update mode of the same kernel (the other half of the iteration, 52 % of the kernel
                    # time in profiles/r02_launches_b.md): 24 B/DoF per launch + the x accumulation (17 passes) that
                    # one launch of the 20-launch timing loop of dpp_time_cg_kernels carries
                    "cg_r_update_gbs": ((24 * ndof + (17 * 8 * ndof) // 20) / (fu_ms * 1e-3) / 1e9 / world
                                        if (fused and stencil_rupd) else None),
                    "cg_r_update_frac": ((24 * ndof + (17 * 8 * ndof) // 20) / (fu_ms * 1e-3) / 1e9 / world / peak
                                         if (fused and stencil_rupd) else None)},
        "solve_roofline": {"achieved": solve_gbs, "peak": peak, "unit": "GB/s", "frac": solve_gbs / peak,
                           "bytes_per_iteration": iter_bytes},
    }
    if world == 1 and args.general_size > 0:
        line["roofline_general"] = general_kernel_entry(args.general_size, peak, peak_kind)
    if world == 1 and args.assembly_size > 0:
        line["assembly"] = assembly_entry(args.assembly_size, peak)
    if not args.no_cpu and world == 1:   # reported on rank 0 at N=1 only
        line["cpu_baseline"] = cpu_baseline(args.cpu_size)
    print(json.dumps(line))


if __name__ == "__main__":
    main()
