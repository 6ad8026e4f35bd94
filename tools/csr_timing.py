"""Assembled-matrix path at BASELINE sizes: CSR assembly time and SpMV bandwidth (one B200)."""
import sys, time
sys.path.insert(0, '.')
import numpy as np
import perphil_b200 as pb
from perphil_b200 import _lib as L
from perphil_b200.solver import configure_handle


def configured_handle(W, prm, bcs):
    """Handle of W with the parameters and Dirichlet data uploaded (package API only: no test / oracle imports)."""
    h = pb.handle_for(W)
    configure_handle(h, W, prm, bcs)
    return h
import ctypes as C
for N in (64, 128):
    mesh = pb.UnitCubeMesh(N, N, N)
    _, V = pb.create_function_spaces(mesh); W = V * V
    prm = pb.DPPParameters(k1=1.0, k2=1e-2, beta=1.0, mu=1.0)
    _, p1, _, p2 = pb.exact_expressions_3d(mesh, prm)
    bcs = [pb.DirichletBC(W.sub(0), p1, "on_boundary"), pb.DirichletBC(W.sub(1), p2, "on_boundary")]
    h = configured_handle(W, prm, bcs)
    nnz = C.c_int64()
    t0 = time.perf_counter(); h._check(h._lib.dpp_assemble_csr(h._h, C.byref(nnz)), "asm"); t1 = time.perf_counter()
    h.set_params(1.0, 2e-2, 1.0, 1.0)   # new values, same pattern: numeric fill only
    t2 = time.perf_counter(); h._check(h._lib.dpp_assemble_csr(h._h, C.byref(nnz)), "asm"); t3 = time.perf_counter()
    sym_ms, num_ms, _ = h.time_assembly(reps=5)
    print(f"Q1 {N}^3: device time symbolic {sym_ms:.2f} ms, numeric {num_ms:.3f} ms = {12*nnz.value/num_ms/1e6:.0f} GB/s of the "
          f"nnz*12 B model ({8*nnz.value/num_ms/1e6:.0f} GB/s of value bytes written)", flush=True)
    ms = h.time_apply(reps=10, warmup=2, assembled=True)
    ndof = 2 * h.n_nodes
    bytes_spmv = 12 * nnz.value + 24 * ndof
    print(f"Q1 {N}^3: nnz {nnz.value}, first assembly (symbolic+numeric) {1e3*(t1-t0):.1f} ms, re-assembly (numeric) {1e3*(t3-t2):.1f} ms "
          f"({12*nnz.value/(t3-t2)/1e9:.0f} GB/s of nnz*12 B), SpMV {ms:.3f} ms = {bytes_spmv/ms/1e6:.0f} GB/s, {ndof/ms/1e6:.1f} GDoF/s", flush=True)
    sol = pb.solve_dpp(W, pb.DPPParameters(k1=1.0, k2=1e-2, beta=1.0, mu=1.0), bcs, solver_parameters=pb.B200_CG_JACOBI_AIJ_PARAMS)
    info = pb.last_solve_info()
    print(f"   Jacobi-CG on the assembled matrix: {sol.iteration_number} its, {info.solve_ms:.1f} ms", flush=True)
    pb.release_handles()
