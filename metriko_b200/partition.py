"""Row partitioning of the face operator for the multi-GPU (one process per GPU) path, SURVEY.md 8e.

The reference has no distributed code at all; this is the host-side plan the partitioned SpMV / PCG needs:

  * every rank owns a contiguous block of face rows (the face order of subdivision / strip meshes is already
    locality-preserving; `morton_order` gives such an order for arbitrary input);
  * the three neighbours of a row are either local or "halo" (owned by a peer). Because the face adjacency is symmetric,
    the rows a peer needs from me are exactly my rows that have a neighbour in the peer's block, so send and receive lists
    are computed locally, without a set-up exchange, both sorted by global row id (hence in matching order);
  * local rows are renumbered interior-first (all neighbours local), boundary-last, so that the interior SpMV can overlap
    the halo exchange;
  * x on a rank is laid out [owned rows (permuted) | halo slots grouped by peer].

`HaloPlan` is pure numpy and is shared by the CUDA path (metriko_b200/dist.py -> C-ABI + NCCL) and by the CPU emulation
used in the world_size-2 gloo tests (tests/test_partition_gloo.py).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np


def block_ranges(n: int, world: int) -> np.ndarray:
    """Row offsets [world + 1] of a balanced contiguous partition."""
    return (np.arange(world + 1, dtype=np.int64) * n) // world


def morton_order(points: np.ndarray, bits: int = 10) -> np.ndarray:
    """Permutation sorting 3-D points (e.g. face centroids) along a Z-order curve: neighbours in space end up close in index."""
    p = np.asarray(points, np.float64)
    lo, hi = p.min(0), p.max(0)
    q = np.minimum(((p - lo) / np.maximum(hi - lo, 1e-300) * (1 << bits)).astype(np.uint64), (1 << bits) - 1)
    code = np.zeros(len(p), np.uint64)
    for b in range(bits):
        for d in range(3):
            code |= ((q[:, d] >> np.uint64(b)) & np.uint64(1)) << np.uint64(3 * b + d)
    return np.argsort(code, kind="stable")


@dataclass
class HaloPlan:
    rank: int
    world: int
    n_global: int
    row0: int                      # first owned global row
    n_local: int
    n_interior: int                # local rows [0, n_interior) have no halo neighbour
    perm: np.ndarray               # local (new) index -> owned offset (old), i.e. global row = row0 + perm[i]
    inv_perm: np.ndarray           # owned offset -> local index
    nbr_local: np.ndarray          # [3, n_local] int32 column index into x_ext = [local | halo]
    halo_global: np.ndarray        # [n_halo] global row ids of the halo slots (sorted, grouped by owner)
    recv_counts: np.ndarray        # [world] halo slots owned by each peer
    send_counts: np.ndarray        # [world] entries sent to each peer
    send_idx: np.ndarray           # [sum send_counts] LOCAL row indices to pack, grouped by peer, sorted by global id
    peers: list = field(default_factory=list)

    @property
    def n_halo(self) -> int:
        return int(self.halo_global.shape[0])

    def owned_global(self) -> np.ndarray:
        return self.row0 + self.perm

    def scatter_from_global(self, v: np.ndarray) -> np.ndarray:
        """Global vector -> this rank's local (permuted) part."""
        return np.ascontiguousarray(v[self.owned_global()])

    def gather_to_global(self, v_local: np.ndarray, out: np.ndarray) -> None:
        out[self.owned_global()] = v_local


def make_halo_plan(nbr: np.ndarray, rank: int, world: int, ranges: np.ndarray | None = None) -> HaloPlan:
    """nbr: [3, nF] global neighbour ids of every face row (the ELL-3 column indices of the face operator)."""
    nbr = np.asarray(nbr)
    nF = nbr.shape[1]
    if ranges is None:
        ranges = block_ranges(nF, world)
    r0, r1 = int(ranges[rank]), int(ranges[rank + 1])
    mine = nbr[:, r0:r1]                                   # [3, n_local] global ids
    n_local = r1 - r0
    is_local = (mine >= r0) & (mine < r1)
    boundary = ~is_local.all(axis=0)
    # interior first, boundary last; stable inside each class so the given (locality-preserving) order survives
    perm = np.concatenate([np.nonzero(~boundary)[0], np.nonzero(boundary)[0]]).astype(np.int64)
    inv_perm = np.empty(n_local, np.int64)
    inv_perm[perm] = np.arange(n_local)
    n_interior = int((~boundary).sum())
    # halo slots: sorted unique external ids (sorted => grouped by owner, owners being contiguous ranges)
    ext = np.unique(mine[~is_local])
    owner = np.searchsorted(ranges, ext, side="right") - 1
    recv_counts = np.bincount(owner, minlength=world).astype(np.int64)
    # column indices into x_ext
    col = np.empty((3, n_local), np.int64)
    loc = mine - r0
    col[is_local] = inv_perm[loc[is_local]]
    col[~is_local] = n_local + np.searchsorted(ext, mine[~is_local])
    nbr_local = np.ascontiguousarray(col[:, perm].astype(np.int32))
    # send lists: my rows with a neighbour in peer p's block, sorted by global id (= the order of p's halo slots)
    send_idx, send_counts = [], np.zeros(world, np.int64)
    for p in range(world):
        if p == rank:
            continue
        hit = ((mine >= ranges[p]) & (mine < ranges[p + 1])).any(axis=0)
        rows = np.nonzero(hit)[0]                           # owned offsets, ascending = ascending global id
        send_counts[p] = len(rows)
        send_idx.append(inv_perm[rows])
    send_idx = np.concatenate(send_idx).astype(np.int32) if send_idx else np.zeros(0, np.int32)
    peers = [p for p in range(world) if p != rank and (send_counts[p] or recv_counts[p])]
    return HaloPlan(rank, world, nF, r0, n_local, n_interior, perm, inv_perm, nbr_local, ext.astype(np.int64), recv_counts, send_counts,
                    send_idx, peers)


def local_stencil(plan: HaloPlan, val: np.ndarray, mass: np.ndarray | None = None):
    """Slices the global ELL-3 values [3, nF] (and the mass vector) to this rank's permuted rows."""
    rows = plan.owned_global()
    v = np.ascontiguousarray(np.asarray(val)[:, rows])
    m = None if mass is None else np.ascontiguousarray(np.asarray(mass)[rows])
    return v, m
