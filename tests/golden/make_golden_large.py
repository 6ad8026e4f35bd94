"""Generates tests/golden/large_sizes.json: oracle pins at the BASELINE sizes the Python oracle cannot reach.

Runs the C/OpenMP oracle (oracle/dpp_oracle_c.c, pinned to oracle/dpp_oracle.py and through it to the
reference's stored numbers) ONCE on the CPU at
  cfg3  3-D hex Q1 256^3, manufactured BCs, KSPCG + PCJACOBI, rtol 1e-8      (BASELINE configs[2])
  cfg5  3-D hex Q1 128^3, k2=1e-6, beta=1e2, constant BCs p1=1, p2=0:       (BASELINE configs[4])
        GMRES(30) + multiplicative fieldsplit (Jacobi-CG blocks, rtol 1e-10), GMRES(30) + Jacobi, CG + Jacobi
and stores what solvers/solver.py:73-75 reports (iteration count, final residual norm) together with the
initial norms, every 10th residual of the history and one line of the solution (x = 0.5, z = 0.5, all y)
per field, plus per-field 2-norms.  The GPU parity tests compare against these numbers
(tests/test_gpu_parity.py); nothing here runs on the GPU box.

usage:  python tests/golden/make_golden_large.py [cfg3] [cfg5] [--n3 256] [--n5 128]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import c_oracle as co  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "large_sizes.json")
INNER = {"ksp_type": "cg", "pc_type": "jacobi", "ksp_rtol": 1e-10, "ksp_atol": 1e-50, "ksp_max_it": 10000}


def line_and_norms(u, N):
    n = N + 1
    nn = n ** 3
    out = {}
    for f, name in enumerate(("p1", "p2")):
        v = u[f * nn:(f + 1) * nn].reshape(n, n, n)
        out[name + "_line_x0.5_z0.5"] = [float(t) for t in v[N // 2, :, N // 2]]
        out[name + "_norm2"] = float(np.linalg.norm(v))
        out[name + "_sum"] = float(np.sum(v))
    return out


def record(res, hist_stride=10):
    h = res.history
    return {"iterations": res.iteration_number, "reason": res.reason, "residual_error": res.residual_error,
            "history_stride": hist_stride, "history": [float(t) for t in h[::hist_stride]],
            "history_last": float(h[-1]) if h else None, "history_len": len(h)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("which", nargs="*", default=["cfg3", "cfg5"])
    ap.add_argument("--n3", type=int, default=256)
    ap.add_argument("--n5", type=int, default=128)
    args = ap.parse_args()
    data = {}
    if os.path.exists(OUT):
        with open(OUT) as f:
            data = json.load(f)
    if "cfg3" in args.which:
        N = args.n3
        t0 = time.time()
        csys = co.manufactured_system((N, N, N), 1)
        b = csys.rhs()
        print(f"cfg3 {N}^3 built in {time.time() - t0:.0f} s, nnz {csys.nnz}", flush=True)
        t0 = time.time()
        res = csys.cg("jacobi", history=2000)
        print(f"cfg3 CG-Jacobi its {res.iteration_number} rnorm {res.residual_error:.15e} in {time.time() - t0:.0f} s",
              flush=True)
        rec = record(res)
        rec.update(line_and_norms(res.u, N))
        rec.update({"cells": N, "n_dof": csys.n_dof, "nnz": csys.nnz, "rhs_norm2": float(np.linalg.norm(b)),
                    "params": {"k1": 1.0, "k2": 1e-2, "beta": 1.0, "mu": 1.0}, "bc": "manufactured",
                    "ksp": "cg", "pc": "jacobi", "rtol": 1e-8, "atol": 1e-12, "threads": co.num_threads()})
        data[f"cfg3_{N}"] = rec
        csys.close()
        with open(OUT, "w") as f:
            json.dump(data, f, indent=0)
    if "cfg5" in args.which:
        N = args.n5
        csys = co.constant_bc_system((N, N, N), 1, k1=1.0, k2=1e-6, beta=1e2, mu=1.0, p1=1.0, p2=0.0)
        b = csys.rhs()
        base = {"cells": N, "n_dof": csys.n_dof, "nnz": csys.nnz, "rhs_norm2": float(np.linalg.norm(b)),
                "params": {"k1": 1.0, "k2": 1e-6, "beta": 1e2, "mu": 1.0}, "bc": ["const", 1.0, 0.0],
                "rtol": 1e-8, "atol": 1e-12, "threads": co.num_threads()}
        runs = {}
        t0 = time.time()
        res = csys.gmres("fieldsplit", inner=INNER, history=4000)
        print(f"cfg5 GMRES+fieldsplit its {res.iteration_number} inner {res.inner_iterations} "
              f"rnorm {res.residual_error:.15e} in {time.time() - t0:.0f} s", flush=True)
        rec = record(res, 1)
        rec["inner_iterations"] = res.inner_iterations
        rec.update(line_and_norms(res.u, N))
        runs["gmres_fieldsplit_multiplicative_cg_jacobi_1e-10"] = rec
        data[f"cfg5_{N}"] = {**base, "runs": runs}
        with open(OUT, "w") as f:
            json.dump(data, f, indent=0)
        t0 = time.time()
        res = csys.cg("jacobi", history=60000)
        print(f"cfg5 CG-Jacobi its {res.iteration_number} rnorm {res.residual_error:.15e} in {time.time() - t0:.0f} s",
              flush=True)
        rec = record(res, 50)
        rec.update(line_and_norms(res.u, N))
        runs["cg_jacobi"] = rec
        data[f"cfg5_{N}"] = {**base, "runs": runs}
        with open(OUT, "w") as f:
            json.dump(data, f, indent=0)
        t0 = time.time()
        res = csys.gmres("jacobi", history=60000, max_it=50000)
        print(f"cfg5 GMRES+Jacobi its {res.iteration_number} rnorm {res.residual_error:.15e} in {time.time() - t0:.0f} s",
              flush=True)
        rec = record(res, 50)
        rec.update(line_and_norms(res.u, N))
        runs["gmres_jacobi"] = rec
        data[f"cfg5_{N}"] = {**base, "runs": runs}
        csys.close()
        with open(OUT, "w") as f:
            json.dump(data, f, indent=0)


if __name__ == "__main__":
    main()
