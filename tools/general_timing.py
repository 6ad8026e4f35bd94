"""General-geometry (element-based) matrix-free apply on shuffled + distorted Q1 hex meshes: time per apply,
GDoF/s, fraction of the HBM roofline for the unstructured bytes model (58 B/node + 32 B/cell) and of the fp64
FMA pipe for the instruction-count model; parity of the cell-block kernel against the row-owner
kernel on the same mesh."""
import json, os, sys, time
sys.path.insert(0, '.')
import numpy as np
from perphil_b200.backend import DppHandle
from tools.general_mesh import shuffled_distorted_hex

sizes = [int(a) for a in sys.argv[1:]] or [64, 128]
peak = 6558.4
if os.path.exists("MEASURED_PEAKS.json"):
    peak = float(json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"])
for N in sizes:
    for distort, shuffle in [(0.25, True), (0.0, True)]:
        t0 = time.perf_counter()
        cnm, X, bn = shuffled_distorted_hex(N, distort, seed=1, shuffle=shuffle)
        t1 = time.perf_counter()
        h = DppHandle.from_mesh_arrays(3, 1, cnm, X, X, cnm, n_nodes=X.shape[0], renumber=bool(distort) or True)
        if not distort:
            h.force_kernel_family(0)      # undistorted lattice: detected as structured; force the general family
        h.set_params(1.0, 1e-2, 1.0, 1.0)
        g = np.random.default_rng(2).standard_normal(bn.size)
        h.set_dirichlet(0, bn, g); h.set_dirichlet(1, bn, -g)
        t2 = time.perf_counter()
        ms0 = h.time_apply(reps=1, warmup=0, with_dot=True)   # includes the one-time cell-block setup
        t3 = time.perf_counter()
        ms = h.time_apply(reps=10, warmup=2, with_dot=True)
        nn, ncell = X.shape[0], cnm.shape[0]
        bytes_model = 58 * nn + 32 * ncell
        inst = (1260 if distort else 720) * ncell           # fp64 instructions
        fp64_peak = 148 * 64 * 1.90e9                       # lanes/clk/SM x SMs x clock (nominal)
        line = dict(N=N, distort=distort, shuffled=shuffle, family=h.info().kernel_family, n_nodes=nn, n_cells=ncell,
                    apply_ms=ms, gdofs=2 * nn / ms / 1e6, hbm_model_gbs=bytes_model / ms / 1e6,
                    hbm_frac=bytes_model / ms / 1e6 / peak, fp64_inst_frac=inst / (ms * 1e-3) / fp64_peak,
                    mesh_build_s=t1 - t0, create_s=t2 - t1, first_apply_s=t3 - t2)
        if N <= 64:
            x = np.random.default_rng(3).standard_normal(2 * nn)
            y = h.apply(x)
            os.environ["DPP_GENERAL_ROW_OWNER"] = "1"
            y_ref = h.apply(x)
            ms_old = h.time_apply(reps=3, warmup=1, with_dot=True)
            del os.environ["DPP_GENERAL_ROW_OWNER"]
            line["rel_err_vs_row_owner"] = float(np.linalg.norm(y - y_ref) / np.linalg.norm(y_ref))
            line["row_owner_ms"] = ms_old
        print(json.dumps(line), flush=True)
        h.close()
