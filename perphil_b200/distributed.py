"""Slab partition of the structured hex/quad meshes across the GPUs of one box.

One process per GPU (torchrun); torch.distributed carries only the plumbing -- broadcasting the
NCCL unique id, barriers and max/sum of scalars for reporting.  The data path (halo planes before
each apply, allreduce of the Krylov scalars) is NCCL inside libdppb200 (csrc/comm.cu).

Partition: the nx+1 node planes x = const are split into `size` contiguous chunks; rank r owns
planes [lo, hi) and additionally stores one ghost plane on each interior side ("forward halo":
owned rows are complete after one neighbour exchange of the input vector, no second message).
With lexicographic numbering (x slowest) a plane is one contiguous index range per field.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Optional

import numpy as np


@dataclass(frozen=True)
class Slab:
    rank: int
    size: int
    nx: int             # global cells along x
    plane_lo: int       # first owned node plane (global index)
    plane_hi: int       # one past the last owned node plane
    local_plane_lo: int  # first stored plane (ghost included)
    local_plane_hi: int  # one past the last stored plane

    @property
    def cell_lo(self) -> int:   # local cells [cell_lo, cell_hi) in global cell numbering
        return self.local_plane_lo

    @property
    def cell_hi(self) -> int:
        return self.local_plane_hi - 1

    @property
    def n_local_planes(self) -> int:
        return self.local_plane_hi - self.local_plane_lo

    def owned_local_planes(self, degree: int = 1):
        """[begin, end) of the owned node planes of a degree-p space in local plane numbering.  The slab
        is cut at vertex planes; a degree-p space has p node planes per cell layer, the rank owns node
        planes [p*plane_lo, p*plane_hi) of the global lattice (clipped to the p*nx + 1 planes there are)."""
        p = degree
        hi = min(p * self.plane_hi, p * self.nx + 1)
        return p * (self.plane_lo - self.local_plane_lo), hi - p * self.local_plane_lo


def make_slab(rank: int, size: int, nx: int) -> Slab:
    planes = nx + 1
    if size > planes:
        raise ValueError("more ranks than node planes")
    lo = (rank * planes) // size
    hi = ((rank + 1) * planes) // size
    llo = max(lo - 1, 0)
    lhi = min(hi + 1, planes)
    # a rank owning the single last plane would have no cell of its own: keep at least one cell
    if lhi - llo < 2:
        llo = max(llo - 1, 0)
    return Slab(rank, size, nx, lo, hi, llo, lhi)


def halo_lists(slab: Slab, plane_nodes: int, degree: int = 1):
    """[(peer, send_local_nodes, recv_local_nodes)] for one scalar field (both fields use the same
    lists).  Local node id = (global_node_plane - p*local_plane_lo) * plane_nodes + in_plane_index.
    Degree p: rows of an owned vertex plane reach p node planes down and up, rows of the other planes
    stay inside their cell layer, so the rank needs the p node planes below its first owned one and the
    single (vertex) node plane above its last owned one."""
    out = []
    p = degree
    base = np.arange(plane_nodes, dtype=np.int64)

    def planes(g0, g1):  # global node planes [g0, g1)
        gp = np.arange(g0, g1, dtype=np.int64)
        return ((gp[:, None] - p * slab.local_plane_lo) * plane_nodes + base[None, :]).ravel().astype(np.int32)

    if slab.rank > 0 and slab.plane_lo > 0:
        # lower neighbour owns the p node planes below p*plane_lo (my lower ghosts) and needs my first owned plane
        out.append((slab.rank - 1, planes(p * slab.plane_lo, p * slab.plane_lo + 1),
                    planes(p * slab.plane_lo - p, p * slab.plane_lo)))
    if slab.rank < slab.size - 1 and slab.plane_hi <= slab.nx:
        out.append((slab.rank + 1, planes(p * slab.plane_hi - p, p * slab.plane_hi),
                    planes(p * slab.plane_hi, p * slab.plane_hi + 1)))
    return out


class SlabComm:
    """Process-group wrapper handed to `UnitCubeMesh(..., comm=...)`."""

    def __init__(self, rank: int, size: int, device: int = 0, backend: Optional[str] = None):
        self.rank, self.size, self.device = rank, size, device
        self._dist = None
        if size > 1:
            import torch
            import torch.distributed as dist

            if not dist.is_initialized():
                if backend is None:
                    backend = "nccl" if torch.cuda.is_available() else "gloo"
                os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
                os.environ.setdefault("MASTER_PORT", "29511")
                kw = {}
                if backend == "nccl":
                    torch.cuda.set_device(device)
                    kw["device_id"] = torch.device("cuda", device)
                dist.init_process_group(backend=backend, rank=rank, world_size=size, **kw)
            self._dist = dist
            self._backend = dist.get_backend()

    @classmethod
    def from_env(cls, backend: Optional[str] = None) -> "SlabComm":
        return cls(int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
                   int(os.environ.get("LOCAL_RANK", "0")), backend)

    # -- partition
    def slab(self, nx: int) -> Slab:
        return make_slab(self.rank, self.size, nx)

    # -- plumbing collectives (host scalars / small objects only)
    def _tensor(self, values, dtype):
        import torch

        dev = torch.device("cuda", self.device) if self._backend == "nccl" else torch.device("cpu")
        return torch.tensor(values, dtype=dtype, device=dev)

    def barrier(self):
        if self._dist is not None:
            self._dist.barrier()

    def max_float(self, v: float) -> float:
        if self._dist is None:
            return float(v)
        import torch

        t = self._tensor([float(v)], torch.float64)
        self._dist.all_reduce(t, op=self._dist.ReduceOp.MAX)
        return float(t.item())

    def sum_int(self, v: int) -> int:
        if self._dist is None:
            return int(v)
        import torch

        t = self._tensor([int(v)], torch.int64)
        self._dist.all_reduce(t, op=self._dist.ReduceOp.SUM)
        return int(t.item())

    def broadcast_bytes(self, payload: Optional[bytes], nbytes: int) -> bytes:
        if self._dist is None:
            return payload
        import torch

        t = self._tensor(list(payload) if self.rank == 0 else [0] * nbytes, torch.uint8)
        self._dist.broadcast(t, src=0)
        return bytes(t.cpu().tolist())

    def all_gather_bytes(self, payload: bytes):
        if self._dist is None:
            return [payload]
        out = [None] * self.size
        self._dist.all_gather_object(out, payload)
        return out

    # -- wire a libdppb200 handle into the slab decomposition
    def attach(self, handle, space_data, V):
        from .backend import nccl_unique_id

        slab = space_data.slab
        if slab is None or self.size == 1:
            return
        degree = int(space_data.degree)
        plane_nodes = int(np.prod(V.grid_nodes[1:]))
        uid = self.broadcast_bytes(nccl_unique_id() if self.rank == 0 else None, 128)
        ob, oe = slab.owned_local_planes(degree)
        handle.comm_init(self.rank, self.size, uid, ob * plane_nodes, oe * plane_nodes)
        for peer, send, recv in halo_lists(slab, plane_nodes, degree):
            handle.comm_add_neighbor(peer, send, recv)
        # peer-memory fast path (CUDA IPC over NVLink): exchange the handles of every rank's residual vector
        # and mailbox; the library falls back to NCCL by itself if any rank cannot take part
        if self._backend == "nccl":
            handle.comm_ipc_import(self.all_gather_bytes(handle.comm_ipc_export()))
        self.agree_on_protocol(handle)

    def agree_on_protocol(self, handle):
        """Every rank must run the same reduction / halo protocol: peer memory or NCCL, fused two-kernel CG (two
        reductions per iteration) or the unfused sequence (three), one kernel family.  Each is a rank-local fact
        (IPC import result; uniform-grid test of the rank's own slab), so the ranks compare and settle on the
        common denominator instead of hanging in mismatched collectives."""
        info = handle.info()
        mine = bytes([info.peer_memory & 0xFF, 1 if handle.fused_cg_supported() else 0, info.kernel_family & 0xFF])
        votes = self.all_gather_bytes(mine)
        if len({v[0] for v in votes}) != 1:
            handle.comm_ipc_disable()
        if len({v[1] for v in votes}) != 1:
            handle.set_fused_cg(False)
        if len({v[2] for v in votes}) != 1:
            raise RuntimeError(f"slab ranks disagree on the kernel family {[v[2] for v in votes]}: the mesh must be "
                               "structured on every rank or on none")

    def destroy(self):
        if self._dist is not None and self._dist.is_initialized():
            self._dist.destroy_process_group()
