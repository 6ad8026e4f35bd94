"""Profile driver: structured Q2 apply at 96^3 (time_apply launches)."""
import sys
sys.path.insert(0, '.')
import perphil_b200 as pb
from tests.util import configured_handle
N = int(sys.argv[1]) if len(sys.argv) > 1 else 96
mesh = pb.UnitCubeMesh(N, N, N)
_, V = pb.create_function_spaces(mesh, pressure_deg=2); W = V * V
prm = pb.DPPParameters(k1=1.0, k2=1e-2, beta=1.0, mu=1.0)
bcs = [pb.DirichletBC(W.sub(0), pb.Constant(1.0), "on_boundary"), pb.DirichletBC(W.sub(1), pb.Constant(0.0), "on_boundary")]
h = configured_handle(W, prm, bcs)
ms = h.time_apply(reps=5, warmup=2, with_dot=True)
print(f"Q2 {N}^3 apply {ms:.4f} ms  {2*h.n_nodes/ms/1e6:.1f} GDoF/s")
