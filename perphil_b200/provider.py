"""Pull the arrays libdppb200 needs out of a (Firedrake or synthetic) mixed space and its BCs.

Reads only documented attributes: `W.sub(i)`, `V.cell_node_map().values`,
`mesh.coordinates.dat.data_ro`, `mesh.coordinates.cell_node_map().values`, `bc.nodes`,
`bc.function_arg`, `bc.function_space().index`.  For real Firedrake spaces the cell-local node
order is normalised to tensor-lexicographic from the node coordinates (never assumed).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Tuple

import numpy as np

from .mesh import Mesh as _SynthMesh, evaluate


@dataclass
class SpaceData:
    dim: int
    degree: int
    n_nodes: int
    cell_node_map: np.ndarray
    coords: np.ndarray
    coord_cell_node_map: np.ndarray
    slab: object = None
    node_coords: Optional[np.ndarray] = None   # pressure-node coordinates (real Firedrake meshes): lattice detection


def _degree_of(V) -> int:
    el = V.ufl_element()
    if isinstance(el, tuple):
        return int(el[1])
    deg = el.degree() if callable(getattr(el, "degree", None)) else getattr(el, "degree")
    if isinstance(deg, (tuple, list)):
        deg = max(deg)
    return int(deg)


def _lexicographic_local_order(cell_coords: np.ndarray) -> np.ndarray:
    """Permutation that sorts the nodes of one cell lexicographically by (x, y[, z])."""
    keys = tuple(np.round(cell_coords[:, d], 12) for d in reversed(range(cell_coords.shape[1])))
    return np.lexsort(keys)


def normalise_local_order(cnm: np.ndarray, node_coords: np.ndarray) -> np.ndarray:
    """Reorder every cell's nodes to tensor-lexicographic order using their coordinates.  Valid for
    axis-aligned (tensor) cells; a cell whose order cannot be inferred raises."""
    out = np.empty_like(cnm)
    first = _lexicographic_local_order(node_coords[cnm[0]])
    trial = cnm[:, first]
    # fast path: the same permutation works for every cell (true for Firedrake's FInAT elements);
    # verified on a sample of cells, otherwise fall back to a per-cell sort
    ident = np.arange(cnm.shape[1])
    sample = np.linspace(0, cnm.shape[0] - 1, min(256, cnm.shape[0])).astype(int)
    if all(np.array_equal(_lexicographic_local_order(node_coords[trial[c]]), ident) for c in sample):
        return np.ascontiguousarray(trial)
    for c in range(cnm.shape[0]):
        out[c] = cnm[c, _lexicographic_local_order(node_coords[cnm[c]])]
    return out


def is_mixed(W) -> bool:
    return hasattr(W, "num_sub_spaces") and W.num_sub_spaces() == 2


def space_data(W) -> SpaceData:
    """W: the 2-field mixed space, or one scalar pressure space V (the per-scale systems of dpp_delayed_form,
    forms/dpp.py:135-205, live on V; the handle then holds V x V and serves its diagonal blocks)."""
    V = W.sub(0) if is_mixed(W) else W
    mesh = W.mesh()
    degree = _degree_of(V)
    dim = int(mesh.geometric_dimension())
    coords = np.asarray(mesh.coordinates.dat.data_ro, dtype=np.float64)
    ccnm = np.asarray(mesh.coordinates.cell_node_map().values, dtype=np.int32)
    cnm = np.asarray(V.cell_node_map().values, dtype=np.int32)
    n_nodes = int(getattr(V, "node_count", None) or V.dim())
    node_coords = None
    if not isinstance(mesh, _SynthMesh):  # real Firedrake: infer local orders from coordinates
        ccnm = normalise_local_order(ccnm, coords)
        if degree == 1:
            cnm = normalise_local_order(cnm, coords) if cnm is not ccnm else ccnm
            node_coords = coords
        else:
            import firedrake as fd  # noqa: only reachable with a real Firedrake mesh

            Vc = fd.VectorFunctionSpace(mesh, V.ufl_element())
            xn = fd.Function(Vc).interpolate(fd.SpatialCoordinate(mesh)).dat.data_ro
            node_coords = np.asarray(xn)
            cnm = normalise_local_order(cnm, node_coords)
    return SpaceData(dim, degree, n_nodes, cnm, coords, ccnm, getattr(mesh, "slab", None), node_coords)


def bc_data(W, bcs, scalar_field: Optional[int] = None) -> List[Tuple[int, np.ndarray, np.ndarray]]:
    """[(field, nodes, values)] for each DirichletBC (later BCs on the same field override).  `scalar_field`:
    the BCs live on a scalar space V (fd.DirichletBC(V, g, ...)) and constrain that field of V x V."""
    out = []
    for bc in bcs or []:
        Vb = bc.function_space()
        field = getattr(Vb, "index", None)
        if field is None:
            if scalar_field is None:
                raise ValueError("DirichletBC must be built on W.sub(i)")
            field = scalar_field
        nodes = np.asarray(bc.nodes, dtype=np.int32)
        if hasattr(bc, "values"):
            vals = np.asarray(bc.values(), dtype=np.float64)
        else:  # Firedrake: function_arg is a Function (expressions are interpolated at construction) or Constant
            g = bc.function_arg
            if hasattr(g, "dat"):
                vals = np.asarray(g.dat.data_ro, dtype=np.float64)[nodes]
            else:
                vals = np.full(nodes.size, float(g))
        out.append((int(field), nodes, vals))
    merged = {}
    for field, nodes, vals in out:
        if field in merged:
            n0, v0 = merged[field]
            lut = dict(zip(n0.tolist(), v0.tolist()))
            lut.update(zip(nodes.tolist(), vals.tolist()))
            keys = np.fromiter(sorted(lut), dtype=np.int32)
            merged[field] = (keys, np.array([lut[k] for k in keys.tolist()]))
        else:
            merged[field] = (nodes, vals)
    return [(f, n, v) for f, (n, v) in sorted(merged.items())]
