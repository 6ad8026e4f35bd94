"""Synthetic meshes and function spaces with the Firedrake attribute surface the DPP path uses.

The reference builds its meshes with Firedrake (`mesh/builtin.py:4-20` -> fd.UnitSquareMesh,
`notebooks/condition-number-study-3d.py:66` -> fd.UnitCubeMesh(..., hexahedral=True)) and its
spaces with `forms/spaces.py:5-36`.  Firedrake is not installable here, so these classes provide
the same *names* (`cell_node_map().values`, `coordinates.dat.data_ro`,
`num_sub_spaces()`, `sub(i)`, `dim()`, `DirichletBC.nodes/function_arg`, `Function.dat.data`)
over plain numpy arrays; `perphil_b200.provider` reads real Firedrake objects and these alike.

Numbering: lexicographic, x slowest; cell-local node order tensor-lexicographic.
A `comm` (perphil_b200.distributed.SlabComm) makes each rank build only its slab of x-planes.
"""
from __future__ import annotations

from typing import Callable, Optional, Sequence, Tuple, Union

import numpy as np


class _Dat:
    def __init__(self, data):
        self.data = data

    @property
    def data_ro(self):
        return self.data


class _Map:
    def __init__(self, values: np.ndarray):
        self.values = values


class Constant:
    def __init__(self, value: float):
        self.value = float(value)

    def __float__(self):
        return self.value

    def __repr__(self):
        return f"Constant({self.value})"


def _lattice(counts: Sequence[int], axes: Sequence[np.ndarray]) -> np.ndarray:
    grid = np.meshgrid(*axes, indexing="ij")
    return np.stack([g.ravel() for g in grid], axis=1)


def _cell_map(cells: Sequence[int], p: int, plane_offset: int = 0) -> np.ndarray:
    dim = len(cells)
    counts = [p * c + 1 for c in cells]
    strides = [int(np.prod(counts[d + 1:])) for d in range(dim)]
    cell_idx = np.meshgrid(*[np.arange(c, dtype=np.int64) for c in cells], indexing="ij")
    base = sum(p * ci.ravel() * s for ci, s in zip(cell_idx, strides))
    loc = np.meshgrid(*[np.arange(p + 1, dtype=np.int64)] * dim, indexing="ij")
    off = sum(li.ravel() * s for li, s in zip(loc, strides))
    return (base[:, None] + off[None, :]).astype(np.int32)


class Mesh:
    """Tensor-product quad/hex mesh of [0,L]^d with optional slab decomposition along x."""

    def __init__(self, cells: Sequence[int], lengths: Optional[Sequence[float]] = None, comm=None):
        self.cells_global = tuple(int(c) for c in cells)
        self.dim = len(self.cells_global)
        self.lengths = tuple(float(v) for v in (lengths or (1.0,) * self.dim))
        self.comm = comm
        nx = self.cells_global[0]
        if comm is not None and comm.size > 1:
            # cell slabs along x: rank r owns vertex planes [lo, hi) ; keeps one ghost plane per side
            self.slab = comm.slab(nx)
        else:
            self.slab = None
        self._spaces = {}
        V1 = FunctionSpace(self, "CG", 1)
        self._coord_space = V1
        self.coordinates = _CoordinateFunction(V1)

    # Firedrake names
    def geometric_dimension(self):
        return self.dim

    def num_cells(self):
        return self._coord_space.cell_node_map().values.shape[0]

    def local_axes(self, degree: int):
        """1-D node coordinates per axis of the local (slab) lattice for a degree-p space."""
        axes = []
        for d, (c, L) in enumerate(zip(self.cells_global, self.lengths)):
            full = np.linspace(0.0, L, degree * c + 1)
            if d == 0 and self.slab is not None:
                lo, hi = self.slab.cell_lo, self.slab.cell_hi
                full = full[degree * lo: degree * hi + 1]
            axes.append(full)
        return axes


class _CoordinateFunction:
    def __init__(self, V):
        self._V = V
        self.dat = _Dat(V.node_coordinates)

    def cell_node_map(self):
        return self._V.cell_node_map()

    def function_space(self):
        return self._V


class FunctionSpace:
    """Scalar CG space of degree 1 or 2 ("Q1"/"Q2" on quads/hexes)."""

    def __init__(self, mesh: Mesh, family: str = "CG", degree: int = 1):
        if family not in ("CG", "Q", "Lagrange"):
            raise ValueError("only continuous Lagrange spaces are supported")
        if degree not in (1, 2):
            raise ValueError("degree must be 1 or 2")
        self._mesh = mesh
        self.degree = degree
        axes = mesh.local_axes(degree)
        self.grid_nodes = tuple(len(a) for a in axes)
        self.node_coordinates = _lattice(self.grid_nodes, axes)
        cells = [(n - 1) // degree for n in self.grid_nodes]
        self._cnm = _Map(_cell_map(cells, degree))
        self.node_count = self.node_coordinates.shape[0]
        self.index = None
        self.parent = None
        # global boundary test (slab interfaces are not boundary)
        idx = np.meshgrid(*[np.arange(n) for n in self.grid_nodes], indexing="ij")
        onb = np.zeros(idx[0].shape, dtype=bool)
        for d in range(mesh.dim):
            lo_is_boundary = hi_is_boundary = True
            if d == 0 and mesh.slab is not None:
                lo_is_boundary = mesh.slab.cell_lo == 0
                hi_is_boundary = mesh.slab.cell_hi == mesh.cells_global[0]
            if lo_is_boundary:
                onb |= idx[d] == 0
            if hi_is_boundary:
                onb |= idx[d] == self.grid_nodes[d] - 1
        self.boundary_nodes = np.flatnonzero(onb.ravel()).astype(np.int32)

    def mesh(self):
        return self._mesh

    def dim(self):
        return self.node_count

    def cell_node_map(self):
        return self._cnm

    def ufl_element(self):
        return ("CG", self.degree)

    def __mul__(self, other):
        return MixedFunctionSpace((self, other))


class _SubSpace:
    """W.sub(i): the i-th component of a mixed space (what fd.DirichletBC receives)."""

    def __init__(self, parent, index: int, V: FunctionSpace):
        self.parent = parent
        self.index = index
        self._V = V

    def __getattr__(self, name):
        return getattr(self._V, name)

    def dim(self):
        return self._V.dim()


class MixedFunctionSpace:
    def __init__(self, spaces: Sequence[FunctionSpace]):
        self._spaces = tuple(spaces)
        self._subs = tuple(_SubSpace(self, i, V) for i, V in enumerate(self._spaces))

    def num_sub_spaces(self):
        return len(self._spaces)

    def sub(self, i: int):
        return self._subs[i]

    def mesh(self):
        return self._spaces[0].mesh()

    def dim(self):
        return sum(V.dim() for V in self._spaces)

    def __iter__(self):
        return iter(self._subs)


class Function:
    """Nodal coefficient vector(s) on a (mixed) space; `.dat.data` like Firedrake."""

    def __init__(self, space, name: Optional[str] = None, val=None, buffer: Optional[np.ndarray] = None):
        self._space = space
        self.name = name
        if isinstance(space, MixedFunctionSpace):
            # one contiguous field-blocked vector [p1; p2]; the sub-functions are views into it
            dims = [space.sub(i).dim() for i in range(space.num_sub_spaces())]
            flat = np.zeros(sum(dims)) if buffer is None else buffer
            offs = np.concatenate([[0], np.cumsum(dims)])
            self.vector = flat
            self._subs = tuple(Function(space.sub(i), val=flat[offs[i]: offs[i + 1]])
                               for i in range(space.num_sub_spaces()))
            self.dat = _Dat([f.dat.data for f in self._subs])
        else:
            self._subs = ()
            self.dat = _Dat(np.zeros(space.dim()) if val is None else np.asarray(val, dtype=float))
            self.vector = self.dat.data

    def function_space(self):
        return self._space

    def sub(self, i: int):
        return self._subs[i]

    @property
    def subfunctions(self):
        return self._subs

    def split(self):
        return self._subs

    def interpolate(self, expr):
        V = self._space
        self.dat.data[:] = evaluate(expr, V.node_coordinates)
        return self


class Expression:
    """A pointwise expression of the coordinates (stand-in for a UFL expression)."""

    def __init__(self, fn: Callable[[np.ndarray], np.ndarray]):
        self.fn = fn

    def __call__(self, coords: np.ndarray) -> np.ndarray:
        return np.asarray(self.fn(coords), dtype=float)


def evaluate(g, coords: np.ndarray) -> np.ndarray:
    if isinstance(g, Expression):
        return g(coords)
    if isinstance(g, Function):
        return g.dat.data
    if callable(g):
        return np.asarray(g(coords), dtype=float)
    return np.full(coords.shape[0], float(g))


class DirichletBC:
    """fd.DirichletBC(W.sub(i), g, "on_boundary") (README.md:79-82 of the reference)."""

    def __init__(self, V, g, sub_domain="on_boundary"):
        if sub_domain != "on_boundary":
            raise ValueError('synthetic meshes support sub_domain="on_boundary" only')
        self._V = V
        self.function_arg = g
        self.sub_domain = sub_domain
        self.nodes = V.boundary_nodes
        # like Firedrake, the boundary expression is interpolated when the BC is built
        self._values = np.ascontiguousarray(evaluate(g, V.node_coordinates[self.nodes]), dtype=np.float64)
        if isinstance(g, Function):
            self._values = np.ascontiguousarray(g.dat.data[self.nodes])

    def function_space(self):
        return self._V

    def values(self) -> np.ndarray:
        return self._values


def UnitSquareMesh(nx: int, ny: int, quadrilateral: bool = True, comm=None) -> Mesh:
    if not quadrilateral:
        raise ValueError("perphil_b200 covers quadrilateral/hexahedral tensor-product cells only")
    return Mesh((nx, ny), comm=comm)


def UnitCubeMesh(nx: int, ny: int, nz: int, hexahedral: bool = True, comm=None) -> Mesh:
    if not hexahedral:
        raise ValueError("perphil_b200 covers quadrilateral/hexahedral tensor-product cells only")
    return Mesh((nx, ny, nz), comm=comm)


def create_mesh(nx: int, ny: int, quadrilateral: bool = True) -> Mesh:
    """perphil.mesh.builtin.create_mesh (mesh/builtin.py:4-20)."""
    return UnitSquareMesh(nx, ny, quadrilateral=quadrilateral)


def create_function_spaces(mesh: Mesh, velocity_deg: int = 1, pressure_deg: int = 1, velocity_family: str = "CG",
                           pressure_family: str = "CG") -> Tuple[None, FunctionSpace]:
    """perphil.forms.spaces.create_function_spaces (forms/spaces.py:5-36).  The velocity space is
    post-processing only (out of scope) and is returned as None."""
    return None, FunctionSpace(mesh, pressure_family, pressure_deg)
