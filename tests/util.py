"""Shared helpers for the parity tests: build the same problem for the oracle and for libdppb200."""
import numpy as np

import perphil_b200 as pb
from oracle import dpp_oracle as orc


def make_problem(cells, degree=1, k1=1.0, k2=1e-2, beta=1.0, mu=1.0, bc="manufactured"):
    """Returns (W, params, bcs, oracle System)."""
    cells = tuple(cells)
    mesh = pb.UnitSquareMesh(*cells) if len(cells) == 2 else pb.UnitCubeMesh(*cells)
    _, V = pb.create_function_spaces(mesh, pressure_deg=degree)
    W = V * V
    prm = pb.DPPParameters(k1=k1, k2=k2, beta=beta, mu=mu)
    oprm = orc.Params(k1=k1, k2=k2, beta=beta, mu=mu)
    omesh = orc.structured_mesh(cells, degree)
    assert np.array_equal(omesh.cell_node_map, V.cell_node_map().values)
    assert np.allclose(omesh.coords, V.node_coordinates)
    if bc == "manufactured":
        _, p1, _, p2 = pb.exact_expressions(mesh, prm)
        bcs = [pb.DirichletBC(W.sub(0), p1, "on_boundary"), pb.DirichletBC(W.sub(1), p2, "on_boundary")]
        osys = orc.build_system(omesh, oprm, "manufactured")
    elif bc == "homogeneous":
        bcs = [pb.DirichletBC(W.sub(0), pb.Constant(0.0), "on_boundary"),
               pb.DirichletBC(W.sub(1), pb.Constant(0.0), "on_boundary")]
        osys = orc.build_system(omesh, oprm, "homogeneous")
    elif isinstance(bc, tuple) and bc[0] == "const":
        bcs = [pb.DirichletBC(W.sub(0), pb.Constant(bc[1]), "on_boundary"),
               pb.DirichletBC(W.sub(1), pb.Constant(bc[2]), "on_boundary")]
        osys = orc.build_system(omesh, oprm, bc)
    elif bc == "none":
        bcs = []
        e = np.zeros(0, dtype=np.int64)
        osys = orc.build_system(omesh, oprm, (e, np.zeros(0), e, np.zeros(0)))
    else:
        raise ValueError(bc)
    return W, prm, bcs, osys


def configured_handle(W, prm, bcs):
    from perphil_b200.provider import bc_data

    h = pb.handle_for(W)
    h.set_params(float(prm.k1), float(prm.k2), float(prm.beta), float(prm.mu))
    got = {f: (n, v) for f, n, v in bc_data(W, bcs)}
    for f in (0, 1):
        n, v = got.get(f, (np.zeros(0, np.int32), np.zeros(0)))
        h.set_dirichlet(f, n, v)
    return h


def rel_err(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(b), 1e-300))
