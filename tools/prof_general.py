"""Profile driver: a few applies of the general (cell-block) kernel on a shuffled + distorted N^3 Q1 hex mesh."""
import sys
sys.path.insert(0, '.')
import numpy as np
from perphil_b200.backend import DppHandle
from tools.general_mesh import shuffled_distorted_hex
N = int(sys.argv[1]) if len(sys.argv) > 1 else 128
distort = float(sys.argv[2]) if len(sys.argv) > 2 else 0.25
cnm, X, bn = shuffled_distorted_hex(N, distort, seed=1)
h = DppHandle.from_mesh_arrays(3, 1, cnm, X, X, cnm, n_nodes=X.shape[0])
if not distort:
    h.force_kernel_family(0)
h.set_params(1.0, 1e-2, 1.0, 1.0)
g = np.random.default_rng(2).standard_normal(bn.size)
h.set_dirichlet(0, bn, g); h.set_dirichlet(1, bn, -g)
print("general apply %.4f ms" % h.time_apply(reps=4, warmup=2, with_dot=True))
