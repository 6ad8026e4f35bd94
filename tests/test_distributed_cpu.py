"""Host-side slab logic for N > 1, on CPU with the gloo backend (world_size 2 and 3):
partition covers every node plane once, halo send/recv lists pair up between neighbours, ghost
planes receive exactly the neighbour's owned values, and owned-row dot products allreduce to the
global dot -- the two communication primitives libdppb200 implements with NCCL (csrc/comm.cu)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from perphil_b200.distributed import halo_lists, make_slab


@pytest.mark.parametrize("size,nx", [(1, 4), (2, 8), (3, 8), (4, 16), (8, 256), (8, 9), (5, 4)])
def test_partition_covers_planes_once(size, nx):
    owned = []
    for r in range(size):
        s = make_slab(r, size, nx)
        assert 0 <= s.local_plane_lo <= s.plane_lo <= s.plane_hi <= s.local_plane_hi <= nx + 1
        assert s.plane_lo - s.local_plane_lo <= 2 and s.local_plane_hi - s.plane_hi <= 1
        assert s.n_local_planes >= 2 and s.cell_hi > s.cell_lo
        owned += list(range(s.plane_lo, s.plane_hi))
    assert owned == list(range(nx + 1))


@pytest.mark.parametrize("degree", [1, 2])
@pytest.mark.parametrize("size,nx", [(2, 8), (4, 16), (8, 256), (8, 192), (3, 7)])
def test_halo_lists_pair_up(size, nx, degree):
    plane_nodes = 7
    p = degree
    slabs = [make_slab(r, size, nx) for r in range(size)]
    lists = [halo_lists(s, plane_nodes, degree) for s in slabs]
    # owned node planes of the degree-p lattice cover [0, p*nx] exactly once
    owned = np.zeros(p * nx + 1, int)
    for s in slabs:
        ob, oe = s.owned_local_planes(degree)
        owned[p * s.local_plane_lo + ob: p * s.local_plane_lo + oe] += 1
        assert 0 <= ob < oe <= p * (s.local_plane_hi - s.local_plane_lo - 1) + 1
    assert np.all(owned == 1)
    for r, s in enumerate(slabs):
        for peer, send, recv in lists[r]:
            back = [x for x in lists[peer] if x[0] == r]
            assert len(back) == 1
            _, psend, precv = back[0]
            # what I send is what the peer receives, in global numbering
            gl = lambda sl, loc: loc + p * sl.local_plane_lo * plane_nodes
            assert np.array_equal(gl(s, send), gl(slabs[peer], precv))
            assert np.array_equal(gl(s, recv), gl(slabs[peer], psend))
            # sends are owned, receives are ghosts
            ob, oe = s.owned_local_planes(degree)
            assert np.all((send >= ob * plane_nodes) & (send < oe * plane_nodes))
            assert np.all((recv < ob * plane_nodes) | (recv >= oe * plane_nodes))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, size, port, nx, ny, nz, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=size)
    try:
        import perphil_b200 as pb
        from perphil_b200.distributed import SlabComm

        comm = SlabComm(rank, size, backend="gloo")
        mesh = pb.UnitCubeMesh(nx, ny, nz, comm=comm)
        _, V = pb.create_function_spaces(mesh)
        slab = mesh.slab
        plane_nodes = (ny + 1) * (nz + 1)
        # global reference vector, local copy with ghosts poisoned
        rng = np.random.default_rng(0)
        xg = rng.standard_normal((nx + 1) * plane_nodes)
        lo, hi = slab.local_plane_lo * plane_nodes, slab.local_plane_hi * plane_nodes
        x = xg[lo:hi].copy()
        ob, oe = slab.owned_local_planes()
        ghost = np.ones(x.size, bool)
        ghost[ob * plane_nodes: oe * plane_nodes] = False
        x[ghost] = np.nan
        # coordinates of the local lattice are the global ones
        assert np.allclose(V.node_coordinates[:, 0].reshape(slab.n_local_planes, -1)[:, 0],
                           np.linspace(0, 1, nx + 1)[slab.local_plane_lo: slab.local_plane_hi])
        # halo exchange with the lists libdppb200 receives
        reqs, bufs = [], []
        for peer, send, recv in halo_lists(slab, plane_nodes):
            reqs.append(dist.isend(torch.from_numpy(x[send].copy()), peer))
            buf = torch.empty(recv.size, dtype=torch.float64)
            bufs.append((recv, buf))
            reqs.append(dist.irecv(buf, peer))
        for r in reqs:
            r.wait()
        for recv, buf in bufs:
            x[recv] = buf.numpy()
        ok_halo = bool(np.array_equal(x, xg[lo:hi]))
        # owned dot + allreduce == global dot
        part = torch.tensor([float(x[~ghost] @ x[~ghost])], dtype=torch.float64)
        dist.all_reduce(part)
        ok_dot = bool(abs(part.item() - xg @ xg) <= 1e-12 * (xg @ xg))
        # boundary nodes: slab interfaces are not boundary
        gb = np.zeros(((nx + 1), ny + 1, nz + 1), bool)
        gb[0] = gb[-1] = True; gb[:, 0] = gb[:, -1] = True; gb[:, :, 0] = gb[:, :, -1] = True
        loc = gb[slab.local_plane_lo: slab.local_plane_hi].ravel()
        ok_bnd = bool(np.array_equal(np.flatnonzero(loc), V.boundary_nodes))
        out[rank] = (ok_halo, ok_dot, ok_bnd, comm.sum_int(oe - ob), comm.max_float(float(rank)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("size", [2, 3])
def test_gloo_halo_and_allreduce(size):
    nx, ny, nz = 7, 3, 2
    mgr = mp.Manager()
    out = mgr.dict()
    port = _free_port()
    mp.spawn(_worker, args=(size, port, nx, ny, nz, out), nprocs=size, join=True)
    assert len(out) == size
    for r in range(size):
        ok_halo, ok_dot, ok_bnd, planes, mx = out[r]
        assert ok_halo and ok_dot and ok_bnd
        assert planes == nx + 1 and mx == float(size - 1)
