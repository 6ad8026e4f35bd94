"""Profile driver: structured Q2 kernels at N^3 cells (default 96): the stand-alone apply (time_apply), the two kernels
of the fused Jacobi-CG iteration (time_cg_kernels) and, with `solve`, 48 iterations of a real solve (ncu target)."""
import sys
sys.path.insert(0, '.')
import perphil_b200 as pb
from perphil_b200.solver import configure_handle


def configured_handle(W, prm, bcs):
    """Handle of W with the parameters and Dirichlet data uploaded (package API only: no test / oracle imports)."""
    h = pb.handle_for(W)
    configure_handle(h, W, prm, bcs)
    return h
N = int(sys.argv[1]) if len(sys.argv) > 1 else 96
mesh = pb.UnitCubeMesh(N, N, N)
_, V = pb.create_function_spaces(mesh, pressure_deg=2); W = V * V
prm = pb.DPPParameters(k1=1.0, k2=1e-2, beta=1.0, mu=1.0)
bcs = [pb.DirichletBC(W.sub(0), pb.Constant(1.0), "on_boundary"), pb.DirichletBC(W.sub(1), pb.Constant(0.0), "on_boundary")]
h = configured_handle(W, prm, bcs)
ms = h.time_apply(reps=5, warmup=2, with_dot=True)
print(f"Q2 {N}^3 apply {ms:.4f} ms  {2*h.n_nodes/ms/1e6:.1f} GDoF/s")
if h.fused_cg_supported():
    fa, fu, mv = h.time_cg_kernels(reps=10, warmup=3)
    nb = 2 * h.n_nodes * 8
    print(f"Q2 {N}^3 fused apply {fa:.4f} ms ({4*nb/fa/1e6:.0f} GB/s of 4 passes), r-update {fu:.4f} ms "
          f"({3*nb/fu/1e6:.0f} GB/s of 3 passes), plain apply {mv:.4f} ms")
    ab, ub = h.time_cg_block_kernels(0, reps=10, warmup=3)
    print(f"Q2 {N}^3 one-field block: fused apply {ab:.4f} ms ({2*nb/ab/1e6:.0f} GB/s of 4 passes), r-update {ub:.4f} ms "
          f"({1.5*nb/ub/1e6:.0f} GB/s of 3 passes)")
if "solve" in sys.argv:
    opt = h.default_options(); opt.max_it = 48
    _, info = h.solve(opt, want_solution=False)
    print(f"48 iterations: {info.solve_ms:.2f} ms  ({info.solve_ms/48:.4f} ms / iteration)")
