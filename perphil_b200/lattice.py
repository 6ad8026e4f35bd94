"""Lattice detection: is an arbitrarily numbered quad/hex mesh geometrically a rectilinear tensor grid?

Firedrake numbers `UnitSquareMesh(quadrilateral=True)` / `UnitCubeMesh(hexahedral=True)` in DMPlex order
(SURVEY Appendix C), so the arrays pulled from a real mesh never look lexicographic although the mesh
is a tensor grid.  The structured kernels of libdppb200 need lexicographic numbering; this module finds
the permutation (user node -> lexicographic position) so that the handle can be created on the
re-numbered mesh and `dpp_set_numbering` can translate at the host boundary.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Tuple

import numpy as np


@dataclass
class Lattice:
    perm: np.ndarray            # int32 [n_nodes]: user node id -> lexicographic node id (x slowest)
    cells: Tuple[int, ...]      # cells per axis
    axes: Tuple[np.ndarray, ...]  # 1-D vertex coordinates per axis
    cell_node_map: np.ndarray   # lexicographic mesh the library is created on
    vertex_coords: np.ndarray
    cell_vertex_map: np.ndarray

    @property
    def is_identity(self) -> bool:
        return bool(np.array_equal(self.perm, np.arange(self.perm.size, dtype=self.perm.dtype)))


def _cell_map(cells, p):
    from .mesh import _cell_map as cm

    return cm(cells, p)


def detect_lattice(dim: int, degree: int, cell_node_map: np.ndarray, node_coords: np.ndarray,
                   vertex_coords: Optional[np.ndarray] = None, cell_vertex_map: Optional[np.ndarray] = None,
                   rel_tol: float = 1e-9) -> Optional[Lattice]:
    """Returns the lattice description, or None when the nodes do not form a tensor grid whose cells
    are the mesh cells (distorted / genuinely unstructured meshes keep the general kernels)."""
    cnm = np.asarray(cell_node_map)
    X = np.asarray(node_coords, dtype=np.float64)
    n = X.shape[0]
    p = int(degree)
    if X.shape[1] != dim or cnm.shape[1] != (p + 1) ** dim or n == 0:
        return None
    ext = float(np.max(X.max(axis=0) - X.min(axis=0)))
    if not ext > 0.0:
        return None
    tol = rel_tol * ext
    idx, naxes = [], []
    for d in range(dim):
        c = X[:, d]
        order = np.argsort(c, kind="stable")
        cs = c[order]
        new = np.concatenate([[True], np.diff(cs) > tol])
        group = np.cumsum(new) - 1
        i_d = np.empty(n, dtype=np.int64)
        i_d[order] = group
        naxes.append(cs[new])
        idx.append(i_d)
        # every other axis of a lattice has at least p + 1 node planes: more distinct coordinates than
        # n / (p + 1)^(dim - 1) on one axis (a distorted mesh has ~n) cannot be one -- skip the remaining sorts
        if naxes[-1].size * (p + 1) ** (dim - 1) > n:
            return None
    counts = [a.size for a in naxes]
    if int(np.prod(counts)) != n or any((c - 1) % p for c in counts) or any(c < p + 1 for c in counts):
        return None
    lex = np.ravel_multi_index(tuple(idx), counts)
    if np.unique(lex).size != n:
        return None
    cells = tuple((c - 1) // p for c in counts)
    if int(np.prod(cells)) != cnm.shape[0]:
        return None
    # every mesh cell must be a lattice cell (same node set), each lattice cell exactly once
    cell_idx = [idx[d][cnm].min(axis=1) for d in range(dim)]
    if any(np.any(ci % p) for ci in cell_idx):
        return None
    cell_lat = tuple(ci // p for ci in cell_idx)
    if any(np.any(cl >= nc) for cl, nc in zip(cell_lat, cells)):
        return None
    cell_lex = np.ravel_multi_index(cell_lat, cells)
    if np.unique(cell_lex).size != cnm.shape[0]:
        return None
    ref_map = _cell_map(cells, p)  # [n_cells, npc], rows in lexicographic cell order
    if not np.array_equal(np.sort(lex[cnm], axis=1), np.sort(ref_map[cell_lex].astype(np.int64), axis=1)):
        return None
    # the geometry (vertex coordinate field) must be that lattice too: corner l of every cell sits where
    # the corner node of the pressure space sits (cell-local orders are tensor-lexicographic)
    if vertex_coords is not None and cell_vertex_map is not None:
        VX, ccnm = np.asarray(vertex_coords, dtype=np.float64), np.asarray(cell_vertex_map)
        corner = np.array([int(sum(p * b * (p + 1) ** (dim - 1 - d) for d, b in enumerate(bits)))
                           for bits in np.ndindex(*(2,) * dim)])
        if ccnm.shape != (cnm.shape[0], 2 ** dim) or np.abs(VX[ccnm] - X[cnm[:, corner]]).max() > tol:
            return None
    # vertex axes: every p-th node plane
    vaxes = tuple(a[::p].copy() for a in naxes)
    grid = np.meshgrid(*vaxes, indexing="ij")
    vcoords = np.stack([g.ravel() for g in grid], axis=1)
    return Lattice(lex.astype(np.int32), cells, vaxes, ref_map, vcoords, _cell_map(cells, 1))


def _spread21(v: np.ndarray) -> np.ndarray:
    v = v.astype(np.uint64) & np.uint64(0x1FFFFF)
    v = (v | (v << np.uint64(32))) & np.uint64(0x1F00000000FFFF)
    v = (v | (v << np.uint64(16))) & np.uint64(0x1F0000FF0000FF)
    v = (v | (v << np.uint64(8))) & np.uint64(0x100F00F00F00F00F)
    v = (v | (v << np.uint64(4))) & np.uint64(0x10C30C30C30C30C3)
    v = (v | (v << np.uint64(2))) & np.uint64(0x1249249249249249)
    return v


def morton_permutation(coords: np.ndarray) -> np.ndarray:
    """int32 [n]: node id -> position along a Morton (Z-order) curve through the node coordinates.

    Used for meshes that are NOT tensor grids: the element-based kernels stage each block of cells in shared
    memory through that block's sorted node list, so nodes that are close in space should be close in memory
    (Firedrake's own DMPlex/RCM numbering has some locality; a renumbering along the curve gives every cell block a
    few contiguous index ranges whatever numbering the mesh came with).  The map is registered with
    `dpp_set_numbering`, so callers keep their own numbering."""
    X = np.asarray(coords, dtype=np.float64)
    lo, hi = X.min(axis=0), X.max(axis=0)
    scale = np.where(hi > lo, 2097151.0 / np.where(hi > lo, hi - lo, 1.0), 0.0)
    q = ((X - lo) * scale).astype(np.uint64)
    dim = X.shape[1]
    key = np.zeros(X.shape[0], dtype=np.uint64)
    for d in range(dim):
        key |= _spread21(q[:, d]) << np.uint64(dim - 1 - d)
    order = np.argsort(key, kind="stable")          # position -> node
    perm = np.empty(X.shape[0], dtype=np.int32)
    perm[order] = np.arange(X.shape[0], dtype=np.int32)
    return perm
