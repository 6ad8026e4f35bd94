"""Profile driver: a few fused CG kernel launches at 256^3 (no CPU baseline, no e2e)."""
import sys
sys.path.insert(0, '.')
import bench, perphil_b200 as pb
N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
W, V, prm, bcs = bench.build_problem(N)
from perphil_b200.solver import configure_handle


def configured_handle(W, prm, bcs):
    """Handle of W with the parameters and Dirichlet data uploaded (package API only: no test / oracle imports)."""
    h = pb.handle_for(W)
    configure_handle(h, W, prm, bcs)
    return h
h = configured_handle(W, prm, bcs)
a, u, m = h.time_cg_kernels(reps=4, warmup=2)
print("fused apply %.4f ms, r update %.4f ms, padded TMA matvec %.4f ms" % (a, u, m))
print("plain apply %.4f ms" % h.time_apply(reps=4, warmup=2, with_dot=True))
