"""Synthetic GENERAL hexahedral meshes for measurements and tests: the N^3 unit-cube lattice with randomly
perturbed interior vertices (non-affine cells), randomly permuted node numbering and randomly permuted cell
order -- what the library sees when a mesh is not a tensor grid.  Deterministic (seeded)."""
import numpy as np


def shuffled_distorted_hex(N, distort=0.25, seed=0, shuffle=True):
    """(cell_node_map [N^3, 8] int32, coords [(N+1)^3, 3], boundary_nodes int32) of a Q1 hex mesh."""
    n = N + 1
    rng = np.random.default_rng(seed)
    ax = np.linspace(0.0, 1.0, n)
    X = np.stack(np.meshgrid(ax, ax, ax, indexing="ij"), axis=-1).reshape(-1, 3)
    idx = np.arange(n ** 3, dtype=np.int64).reshape(n, n, n)
    corners = [idx[a:a + N, b:b + N, c:c + N].reshape(-1) for a in (0, 1) for b in (0, 1) for c in (0, 1)]
    cnm = np.stack(corners, axis=1)                     # local order a*4 + b*2 + c (x slowest)
    onb = np.zeros((n, n, n), bool)
    onb[0] = onb[-1] = onb[:, 0] = onb[:, -1] = onb[:, :, 0] = onb[:, :, -1] = True
    onb = onb.reshape(-1)
    if distort:
        X[~onb] += distort / N * (rng.random((int((~onb).sum()), 3)) - 0.5)
    if shuffle:
        perm = rng.permutation(n ** 3)                  # old -> new node id
        cperm = rng.permutation(N ** 3)
        cnm = perm[cnm][cperm]
        Xn = np.empty_like(X); Xn[perm] = X
        bn = np.sort(perm[np.flatnonzero(onb)])
        X = Xn
    else:
        bn = np.flatnonzero(onb)
    return np.ascontiguousarray(cnm, dtype=np.int32), np.ascontiguousarray(X), bn.astype(np.int32)
