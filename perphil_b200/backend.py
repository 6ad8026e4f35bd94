"""numpy-facing wrapper of one libdppb200 handle (one mesh + function space on one GPU)."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np

from . import _lib as L


class DppError(RuntimeError):
    pass


@dataclass
class SolveInfo:
    iterations: int
    converged_reason: int
    inner_iterations: int
    residual_norm: float
    rhs_norm: float
    solve_ms: float
    setup_ms: float
    apply_count: int
    history: np.ndarray


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class DppHandle:
    """Owns the device copy of a mesh/function space; mirrors include/dpp_b200.h one to one."""

    def __init__(self, dim: int, degree: int, cell_node_map: np.ndarray, coords: np.ndarray,
                 coord_cell_node_map: Optional[np.ndarray] = None, n_nodes: Optional[int] = None, device: int = 0):
        self._lib = L.load()
        self._h = C.c_void_p()
        cnm = np.ascontiguousarray(cell_node_map, dtype=np.int32)
        xyz = np.ascontiguousarray(coords, dtype=np.float64)
        ccnm = cnm if coord_cell_node_map is None else np.ascontiguousarray(coord_cell_node_map, dtype=np.int32)
        if n_nodes is None:
            n_nodes = int(cnm.max()) + 1
        self.n_nodes = int(n_nodes)
        self.dim, self.degree = dim, degree
        rc = self._lib.dpp_create(C.byref(self._h), device, dim, degree, self.n_nodes, cnm.shape[0], cnm.shape[1],
                                  _ptr(cnm), xyz.shape[0], _ptr(xyz), _ptr(ccnm))
        if rc != 0:
            msg = self._lib.dpp_last_error(None).decode()
            self._h = C.c_void_p()
            raise DppError(f"dpp_create failed ({rc}): {msg}")
        self._perm = None   # user -> internal node map (set_numbering)
        self._uploaded = {}  # host copies of what set_params / set_dirichlet last sent (skip identical uploads)

    @classmethod
    def from_mesh_arrays(cls, dim: int, degree: int, cell_node_map: np.ndarray, node_coords: np.ndarray,
                         vertex_coords: np.ndarray, coord_cell_node_map: np.ndarray, n_nodes: Optional[int] = None,
                         device: int = 0, renumber: bool = True) -> "DppHandle":
        """Handle for an arbitrarily numbered mesh.  If the mesh is geometrically a rectilinear tensor grid
        (perphil_b200.lattice) it is created on the lexicographically re-numbered mesh, so that the
        structured kernels serve it, and the numbering map is registered: every method of this class keeps
        speaking the caller's numbering."""
        from .lattice import detect_lattice, morton_permutation

        lat = detect_lattice(dim, degree, cell_node_map, node_coords, vertex_coords, coord_cell_node_map) if renumber else None
        if lat is not None:
            if lat.is_identity:
                return cls(dim, degree, cell_node_map, vertex_coords, coord_cell_node_map, n_nodes=n_nodes, device=device)
            h = cls(dim, degree, lat.cell_node_map, lat.vertex_coords, lat.cell_vertex_map, n_nodes=lat.perm.size,
                    device=device)
            h.set_numbering(lat.perm)
            return h
        if not renumber:
            return cls(dim, degree, cell_node_map, vertex_coords, coord_cell_node_map, n_nodes=n_nodes, device=device)
        # not a tensor grid: the general (element-based) kernels.  Re-number nodes and vertices along a Morton
        # curve so that the node list of every cell block is a few contiguous ranges (coalesced gathers)
        cnm = np.asarray(cell_node_map)
        ccnm = cnm if coord_cell_node_map is None else np.asarray(coord_cell_node_map)
        perm = morton_permutation(node_coords)
        same = ccnm is cnm or (ccnm.shape == cnm.shape and np.asarray(vertex_coords).shape[0] == perm.size
                               and np.array_equal(ccnm, cnm))
        vperm = perm if same else morton_permutation(vertex_coords)
        vx = np.empty_like(np.asarray(vertex_coords, dtype=np.float64))
        vx[vperm] = np.asarray(vertex_coords, dtype=np.float64)
        new_cnm = np.ascontiguousarray(perm[cnm], dtype=np.int32)
        new_ccnm = new_cnm if same else np.ascontiguousarray(vperm[ccnm], dtype=np.int32)
        h = cls(dim, degree, new_cnm, vx, new_ccnm, n_nodes=perm.size, device=device)
        h.set_numbering(perm)
        return h

    def set_numbering(self, user_to_internal: np.ndarray):
        perm = np.ascontiguousarray(user_to_internal, dtype=np.int32)
        if perm.size != self.n_nodes:
            raise ValueError("numbering map must have one entry per node")
        self._check(self._lib.dpp_set_numbering(self._h, _ptr(perm)), "dpp_set_numbering")
        self._perm = perm
        self._uploaded = {k: v for k, v in self._uploaded.items() if k == "params"}   # the library dropped its BCs

    # -- plumbing
    def _check(self, rc: int, what: str):
        if rc != 0:
            raise DppError(f"{what} failed ({rc}): {self._lib.dpp_last_error(self._h).decode()}")

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._lib.dpp_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- setup
    def info(self) -> L.DppInfo:
        info = L.DppInfo()
        self._check(self._lib.dpp_get_info(self._h, C.byref(info)), "dpp_get_info")
        return info

    def force_kernel_family(self, family: int):
        self._check(self._lib.dpp_force_kernel_family(self._h, family), "dpp_force_kernel_family")

    def set_params(self, k1: float, k2: float, beta: float, mu: float):
        """dpp_set_params -- skipped when the handle already holds exactly these values (the call invalidates
        the diagonal, the boundary classification and the captured CUDA graphs of the handle)."""
        prm = (float(k1), float(k2), float(beta), float(mu))
        if self._uploaded.get("params") == prm:
            return
        self._check(self._lib.dpp_set_params(self._h, *prm), "dpp_set_params")
        self._uploaded["params"] = prm

    def set_dirichlet(self, field: int, nodes: Sequence[int], values: Sequence[float]):
        """dpp_set_dirichlet -- skipped when nodes and values are unchanged since the last upload."""
        nodes = np.ascontiguousarray(nodes, dtype=np.int32)
        values = np.ascontiguousarray(values, dtype=np.float64)
        if nodes.shape != values.shape:
            raise ValueError("nodes and values must have the same length")
        old = self._uploaded.get(("bc", int(field)))
        if old is not None and old[0].shape == nodes.shape and np.array_equal(old[0], nodes) and np.array_equal(old[1], values):
            return
        self._check(self._lib.dpp_set_dirichlet(self._h, field, nodes.size, _ptr(nodes), _ptr(values)),
                    "dpp_set_dirichlet")
        self._uploaded[("bc", int(field))] = (nodes.copy(), values.copy())

    def comm_init(self, rank: int, world: int, unique_id: Optional[bytes], owned_begin: int, owned_end: int):
        buf = C.create_string_buffer(unique_id, 128) if unique_id is not None else None
        self._check(self._lib.dpp_comm_init(self._h, rank, world, buf, owned_begin, owned_end), "dpp_comm_init")

    def comm_add_neighbor(self, peer: int, send_nodes, recv_nodes):
        s = np.ascontiguousarray(send_nodes, dtype=np.int32)
        r = np.ascontiguousarray(recv_nodes, dtype=np.int32)
        self._check(self._lib.dpp_comm_add_neighbor(self._h, peer, s.size, _ptr(s), r.size, _ptr(r)),
                    "dpp_comm_add_neighbor")

    def comm_ipc_export(self) -> bytes:
        """This rank's CUDA IPC blob (residual vector + mailbox handles) for the peer-memory fast path."""
        n = self._lib.dpp_comm_ipc_blob_size()
        buf = C.create_string_buffer(n)
        self._check(self._lib.dpp_comm_ipc_export(self._h, buf), "dpp_comm_ipc_export")
        return buf.raw

    def comm_ipc_import(self, blobs) -> None:
        """blobs: the exported blobs of ALL ranks, in rank order."""
        raw = b"".join(blobs)
        buf = C.create_string_buffer(raw, len(raw))
        self._check(self._lib.dpp_comm_ipc_import(self._h, buf), "dpp_comm_ipc_import")

    def comm_ipc_disable(self) -> None:
        self._check(self._lib.dpp_comm_ipc_disable(self._h), "dpp_comm_ipc_disable")

    def fused_cg_supported(self) -> bool:
        return self._lib.dpp_fused_cg_supported(self._h) == 1

    def set_fused_cg(self, enable: bool) -> None:
        self._check(self._lib.dpp_set_fused_cg(self._h, 1 if enable else 0), "dpp_set_fused_cg")

    # -- operator
    def apply(self, x: np.ndarray, assembled: bool = False) -> np.ndarray:
        x = np.ascontiguousarray(x, dtype=np.float64)
        if x.size != 2 * self.n_nodes:
            raise ValueError("x must have 2*n_nodes entries")
        y = np.empty_like(x)
        self._check(self._lib.dpp_apply_host(self._h, _ptr(x), _ptr(y), L.OP_ASSEMBLED if assembled else L.OP_MATRIX_FREE),
                    "dpp_apply_host")
        return y

    def diagonal(self) -> np.ndarray:
        d = np.empty(2 * self.n_nodes)
        self._check(self._lib.dpp_get_diagonal_host(self._h, _ptr(d)), "dpp_get_diagonal_host")
        return d

    def assemble_csr(self):
        nnz = C.c_int64()
        self._check(self._lib.dpp_assemble_csr(self._h, C.byref(nnz)), "dpp_assemble_csr")
        indptr = np.empty(2 * self.n_nodes + 1, dtype=np.int64)
        indices = np.empty(nnz.value, dtype=np.int32)
        data = np.empty(nnz.value, dtype=np.float64)
        self._check(self._lib.dpp_get_csr_host(self._h, _ptr(indptr), _ptr(indices), _ptr(data)), "dpp_get_csr_host")
        if self._perm is not None:
            # the library assembled on the re-numbered mesh: A_user[u, v] = A_int[perm[u], perm[v]]
            import scipy.sparse as sp

            n = self.n_nodes
            pd = np.concatenate([self._perm.astype(np.int64), n + self._perm.astype(np.int64)])
            A = sp.csr_matrix((data, indices, indptr), shape=(2 * n, 2 * n))[pd][:, pd].tocsr()
            A.sort_indices()
            return A.indptr.astype(np.int64), A.indices.astype(np.int32), A.data
        return indptr, indices, data

    def assemble_csr_block(self, row_field: int, col_field: int):
        """Block (row_field, col_field) of the assembled matrix as a scalar-space CSR triplet."""
        nnz = C.c_int64()
        self._check(self._lib.dpp_assemble_csr(self._h, C.byref(nnz)), "dpp_assemble_csr")
        n = self.n_nodes
        indptr = np.empty(n + 1, dtype=np.int64)
        indices = np.empty(nnz.value // 4, dtype=np.int32)
        data = np.empty(nnz.value // 4, dtype=np.float64)
        self._check(self._lib.dpp_get_csr_block_host(self._h, int(row_field), int(col_field), _ptr(indptr), _ptr(indices),
                                                     _ptr(data)), "dpp_get_csr_block_host")
        if self._perm is not None:
            import scipy.sparse as sp

            pm = self._perm.astype(np.int64)
            A = sp.csr_matrix((data, indices, indptr), shape=(n, n))[pm][:, pm].tocsr()
            A.sort_indices()
            return A.indptr.astype(np.int64), A.indices.astype(np.int32), A.data
        return indptr, indices, data

    def time_assembly(self, reps: int = 3):
        """(symbolic_ms, numeric_ms, nnz): device time of the two CSR assembly phases."""
        a, b, nnz = C.c_double(), C.c_double(), C.c_int64()
        self._check(self._lib.dpp_time_assembly(self._h, int(reps), C.byref(a), C.byref(b), C.byref(nnz)),
                    "dpp_time_assembly")
        return a.value, b.value, nnz.value

    # -- solve
    def default_options(self) -> L.DppOptions:
        o = L.DppOptions()
        self._lib.dpp_default_options(C.byref(o))
        return o

    def solve(self, options: Optional[L.DppOptions] = None, want_solution: bool = True, history: int = 0,
              out: Optional[np.ndarray] = None):
        opt = options if options is not None else self.default_options()
        u = None
        if want_solution:
            u = out if out is not None else np.empty(2 * self.n_nodes)
        hist = np.zeros(max(history, 0))
        res = L.DppResult()
        self._check(self._lib.dpp_solve(self._h, C.byref(opt), _ptr(u), C.byref(res), _ptr(hist) if history > 0 else None,
                                        int(history)), "dpp_solve")
        info = SolveInfo(res.iterations, res.converged_reason, res.inner_iterations, res.residual_norm, res.rhs_norm,
                         res.solve_ms, res.setup_ms, res.apply_count, hist[: res.history_len].copy())
        return u, info

    # -- measurement
    def time_apply(self, reps: int = 20, warmup: int = 3, assembled: bool = False, with_dot: bool = False) -> float:
        ms = C.c_double()
        self._check(self._lib.dpp_time_apply(self._h, L.OP_ASSEMBLED if assembled else L.OP_MATRIX_FREE, warmup, reps,
                                             1 if with_dot else 0, C.byref(ms)), "dpp_time_apply")
        return ms.value

    def time_cg_kernels(self, reps: int = 20, warmup: int = 3):
        """(apply_ms, update_ms, matvec_ms): the two kernels of one fused Jacobi-CG iteration and the plain
        matrix-free apply of the same (TMA, padded-layout) kernel, each timed alone."""
        a, u, m = C.c_double(), C.c_double(), C.c_double()
        self._check(self._lib.dpp_time_cg_kernels(self._h, warmup, reps, C.byref(a), C.byref(u), C.byref(m)),
                    "dpp_time_cg_kernels")
        return a.value, u.value, m.value

    def time_cg_block_kernels(self, field: int = 0, reps: int = 20, warmup: int = 3):
        """(apply_ms, update_ms) of the fused Jacobi-CG iteration on the one-field diagonal block of `field`."""
        a, u = C.c_double(), C.c_double()
        self._check(self._lib.dpp_time_cg_block_kernels(self._h, int(field), warmup, reps, C.byref(a), C.byref(u)),
                    "dpp_time_cg_block_kernels")
        return a.value, u.value

    def error_norms(self, u: Optional[np.ndarray] = None, exact: Optional[np.ndarray] = None, nq: int = 6):
        """(L2_1, L2_2, H1semi_1, H1semi_2) of u - exact; u None = the last solve's solution, exact None = the
        manufactured closed form for the handle's parameters."""
        out = np.zeros(4)
        uu = None if u is None else np.ascontiguousarray(u, dtype=np.float64)
        ee = None if exact is None else np.ascontiguousarray(exact, dtype=np.float64)
        self._check(self._lib.dpp_error_norms(self._h, _ptr(uu), _ptr(ee), int(nq), _ptr(out)), "dpp_error_norms")
        return tuple(float(np.sqrt(v)) for v in out)

    def darcy_velocity(self, conductivity: float, p: Optional[np.ndarray] = None, field: int = 0, rtol: float = 1e-8,
                       max_it: int = 10000):
        """L2 projection of -k grad(p_h) into V^dim: (velocity [dim, n_nodes], CG iterations per component).
        p None = field `field` of the last solve's solution."""
        pp = None if p is None else np.ascontiguousarray(p, dtype=np.float64)
        if pp is not None and pp.size != self.n_nodes:
            raise ValueError("p must have n_nodes entries")
        vel = PINNED.take(self.dim * self.n_nodes).reshape(self.dim, self.n_nodes)   # page-locked: full-rate D2H
        its = np.zeros(self.dim, dtype=np.int32)
        self._check(self._lib.dpp_darcy_velocity(self._h, _ptr(pp), int(field), float(conductivity), float(rtol),
                                                 int(max_it), _ptr(vel), _ptr(its)), "dpp_darcy_velocity")
        return vel, its

    def lanczos(self, steps: int, which: int = 0, seed: int = 0):
        """(alpha, beta) of `steps` Lanczos steps on A_bc (which=0) or its diagonal block A00 (1) / A11 (2)."""
        a = np.zeros(int(steps))
        b = np.zeros(int(steps))
        done = C.c_int32()
        self._check(self._lib.dpp_lanczos(self._h, int(which), int(steps), int(seed), _ptr(a), _ptr(b), C.byref(done)),
                    "dpp_lanczos")
        return a[: done.value].copy(), b[: done.value].copy()

    def launch_count(self) -> int:
        n = C.c_int64()
        self._check(self._lib.dpp_kernel_launch_count(self._h, C.byref(n)), "dpp_kernel_launch_count")
        return n.value


class _Lease:
    """Returns its pinned block to the pool when the last numpy view of it dies."""

    def __init__(self, pool, nbytes, ptr):
        self.pool, self.nbytes, self.ptr = pool, nbytes, ptr

    def __del__(self):
        try:
            self.pool._give(self.nbytes, self.ptr)
        except Exception:
            pass


class PinnedPool:
    """Page-locked result buffers (cudaMallocHost through the C ABI), recycled by size: the D2H copy
    of a 272 MB solution runs at PCIe rate instead of the pageable-memory rate."""

    def __init__(self):
        self._free = {}

    def take(self, n_doubles: int) -> np.ndarray:
        nbytes = 8 * int(n_doubles)
        stack = self._free.get(nbytes)
        if stack:
            ptr = stack.pop()
        else:
            p = C.c_void_p()
            rc = L.load().dpp_host_alloc(C.byref(p), nbytes)
            if rc != 0:
                raise DppError(f"dpp_host_alloc({nbytes}) failed ({rc})")
            ptr = p.value
        buf = (C.c_double * int(n_doubles)).from_address(ptr)
        buf._lease = _Lease(self, nbytes, ptr)   # numpy keeps `buf` alive through .base
        return np.frombuffer(buf, dtype=np.float64)

    def _give(self, nbytes, ptr):
        stack = self._free.setdefault(nbytes, [])
        if len(stack) < 2:
            stack.append(ptr)
        else:
            L.load().dpp_host_free(C.c_void_p(ptr))


PINNED = PinnedPool()


def nccl_unique_id() -> bytes:
    buf = C.create_string_buffer(128)
    rc = L.load().dpp_nccl_unique_id(buf)
    if rc != 0:
        raise DppError(f"dpp_nccl_unique_id failed ({rc})")
    return buf.raw
