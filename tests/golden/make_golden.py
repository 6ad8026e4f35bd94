"""Extract the reference's own stored numbers into small JSON fixtures.

Run ONCE in the build container (where /root/reference exists):

    python tests/golden/make_golden.py

It copies *numbers only* (no code) out of artefacts that the reference
repository itself stores:

  notebooks/results-conforming-3d/conditioning/conditioning_3d.csv      (3D hex Q1 kappa)
  notebooks/results-conforming-2d/conditioning/conditioning.csv         (2D quad Q1 kappa)
  notebooks/results-conforming-2d/convergence.csv                       (its / residual / errors)
  notebooks/conforming-galerkin-fem-operator-splitting-2D-perphil.ipynb (cell outputs: KSP/SNES
      monitors, x=0.5 slices, condition numbers)

The fixtures written next to this script are what `tests/` (and nothing else) reads; the GPU box
has no /root/reference.
"""
import csv
import json
import os
import re

REF = "/root/reference/notebooks"
HERE = os.path.dirname(os.path.abspath(__file__))


def read_csv(path):
    with open(path) as f:
        return list(csv.DictReader(f))


def floats_in_arrays(text):
    arrays = re.findall(r"array\(\[(.*?)\]\)", text, flags=re.S)
    return [[float(v) for v in a.replace("\n", " ").split(",") if v.strip()] for a in arrays]


def monitor(text, key):
    out = []
    for line in text.splitlines():
        m = re.match(r"\s*(\d+) %s\s+([0-9.eE+-]+)" % key, line)
        if m:
            out.append([int(m.group(1)), float(m.group(2))])
    return out


def main():
    gold = {}
    rows = read_csv(f"{REF}/results-conforming-3d/conditioning/conditioning_3d.csv")
    gold["conditioning_3d_hex_q1"] = [
        {k: (int(r[k]) if k in ("N", "n_dofs", "n0", "n1") else float(r[k])) for k in r} for r in rows
    ]
    rows = read_csv(f"{REF}/results-conforming-2d/conditioning/conditioning.csv")
    gold["conditioning_2d_quad_q1"] = [
        {k: (int(r[k]) if k == "N" else float(r[k])) for k in r} for r in rows
    ]
    rows = read_csv(f"{REF}/results-conforming-2d/convergence.csv")
    conv = []
    for r in rows:
        conv.append(
            {
                "N": int(r["N"]),
                "degree": int(r["degree"]),
                "solver": r["solver"],
                "it": int(r["it"]),
                "res": float(r["res"]),
                "e1_L2": float(r["e1_L2"]),
                "e2_L2": float(r["e2_L2"]),
                "e1_H1s": float(r["e1_H1s"]),
                "e2_H1s": float(r["e2_H1s"]),
            }
        )
    gold["convergence_2d"] = conv

    nb = json.load(open(f"{REF}/conforming-galerkin-fem-operator-splitting-2D-perphil.ipynb"))
    outs = {}
    for i, c in enumerate(nb["cells"]):
        if c["cell_type"] != "code":
            continue
        for o in c.get("outputs", []):
            t = "".join(o.get("text", [])) if "text" in o else "".join(
                o.get("data", {}).get("text/plain", [])
            )
            if t:
                outs.setdefault(i, "")
                outs[i] += t
    nb_gold = {
        "setup": "10x10 quad Q1 UnitSquare, k1=1 k2=1e-2 beta=1 mu=1, manufactured Dirichlet BCs, "
        "executed with ksp_rtol=1e-12 (per-iteration norms are tolerance independent)",
        "slice_x0.5_monolithic_lu": floats_in_arrays(outs[15]),
        "plain_gmres_ksp": monitor(outs[18], "KSP Residual norm"),
        "plain_gmres_snes": monitor(outs[18], "SNES Function norm"),
        "slice_x0.5_plain_gmres": floats_in_arrays(outs[19]),
        "fieldsplit_mult_lu_gmres_ksp": monitor(outs[27], "KSP Residual norm"),
        "fieldsplit_mult_lu_gmres_snes": monitor(outs[27], "SNES Function norm"),
        "ngs_snes": monitor(outs[32], "SNES Function norm"),
        "cond_monolithic": float(re.search(r"Number: ([0-9.eE+-]+)", outs[43]).group(1)),
        "cond_macro": float(re.search(r"Macro system Condition Number: ([0-9.eE+-]+)", outs[45]).group(1)),
        "cond_micro": float(re.search(r"Micro system Condition Number: ([0-9.eE+-]+)", outs[45]).group(1)),
    }
    gold["operator_splitting_notebook_10x10"] = nb_gold
    with open(os.path.join(HERE, "reference_stored.json"), "w") as f:
        json.dump(gold, f, indent=1)
    print("wrote", os.path.join(HERE, "reference_stored.json"))
    print({k: (len(v) if hasattr(v, "__len__") else v) for k, v in gold.items()})
    print("plain gmres its", len(nb_gold["plain_gmres_ksp"]), "ngs its", len(nb_gold["ngs_snes"]))


if __name__ == "__main__":
    main()
