"""Host side of the multi-GPU slab decomposition (one process per GPU).

The reference is single-process; this module is the launch-side plumbing only:
choose slab edges on cell-layer boundaries so that every rank owns about the same number of
particles, hand each rank its particles, and bootstrap the library's NCCL communicator by
broadcasting the unique id with `torch.distributed`.  Halo exchange, migration and the Δt
all-reduce run inside libsphb200 (csrc/sph_slab*.cuh).
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np

INT64_MIN, INT64_MAX = -(2 ** 63), 2 ** 63 - 1


def cell_coord(x: np.ndarray, H_inv: float) -> np.ndarray:
    """map_floor of the reference (src/SPHCellList.jl:56-61): sign(x)·trunc(|x|·H⁻¹ + ½)."""
    x = np.asarray(x, np.float64)
    return (np.sign(x) * np.trunc(np.abs(x) * H_inv + 0.5)).astype(np.int64)


def plan_edges(coords: np.ndarray, world: int, min_width: int = 2, weights: Optional[np.ndarray] = None) -> List[int]:
    """Slab edges e[0] < e[1] < ... < e[world] on cell-layer boundaries; rank r owns layers
    e[r] <= c < e[r+1].  The layers are cut into `world` contiguous groups such that the LARGEST
    group load is as small as possible (the step time is the slowest rank's): equal-width slabs
    are unusable for a dam break (the water starts in one corner of the tank), and rounding each
    particle-count quantile to the nearest layer boundary — the obvious choice — can leave a rank a
    whole layer (7 % at 14 layers per rank) above the mean.  `weights` (per particle, default 1)
    lets cheap particles count less (a wall particle has a fraction of a fluid particle's neighbours).
    Every slab is at least `min_width` layers wide (the exchange protocol needs 2)."""
    coords = np.asarray(coords, np.int64)
    cmin, cmax = int(coords.min()), int(coords.max())
    nlay = cmax - cmin + 1
    if nlay < world * min_width:
        raise ValueError(f"{nlay} cell layers along the slab axis cannot feed {world} ranks of >= {min_width} layers")
    hist = np.bincount(coords - cmin, weights=None if weights is None else np.asarray(weights, np.float64), minlength=nlay)
    cum = np.concatenate([[0.0], np.cumsum(hist, dtype=np.float64)])   # cum[k] = load of layers < k
    load = lambda i, j: cum[j] - cum[i]
    # best[r][k]: smallest achievable maximum load when the first k layers feed r ranks (linear partition DP)
    INF = float("inf")
    best = np.full((world + 1, nlay + 1), INF)
    cut = np.zeros((world + 1, nlay + 1), np.int64)
    best[0][0] = 0.0
    for r in range(1, world + 1):
        for k in range(r * min_width, nlay - (world - r) * min_width + 1):
            lo, hi = (r - 1) * min_width, k - min_width
            if r == 1:
                lo = hi = 0
            j = np.arange(lo, hi + 1)
            cand = np.maximum(best[r - 1][lo:hi + 1], cum[k] - cum[j])
            m = int(np.argmin(cand))
            best[r][k], cut[r][k] = cand[m], j[m]
    edges = [nlay]
    k = nlay
    for r in range(world, 0, -1):
        k = int(cut[r][k])
        edges.append(k)
    edges = edges[::-1]
    assert edges[0] == 0 and all(b - a >= min_width for a, b in zip(edges, edges[1:])), edges
    return [cmin + e for e in edges]


def owner_of(coords: np.ndarray, edges: Sequence[int]) -> np.ndarray:
    """rank index of every particle (layers outside the planned range go to the end ranks)."""
    inner = np.asarray(edges[1:-1], np.int64)
    return np.searchsorted(inner, np.asarray(coords, np.int64), side="right")


def slab_bounds(edges: Sequence[int], rank: int):
    world = len(edges) - 1
    lo = INT64_MIN if rank == 0 else int(edges[rank])
    hi = INT64_MAX if rank == world - 1 else int(edges[rank + 1])
    return lo, hi


def best_axis(positions: np.ndarray, H_inv: float, world: int) -> int:
    """Default slab axis: among y (and z) the one with the most cell layers.  x is accepted too
    (pass axis=0; the rows of cells then run along y), but a flow that runs along x — a dam break —
    keeps re-distributing particles along it, while a y split stays balanced through the run."""
    D = positions.shape[1]
    best, best_layers = 1, -1
    for ax in range(1, D):
        c = cell_coord(positions[:, ax], H_inv)
        layers = int(c.max() - c.min() + 1)
        if layers > best_layers:
            best, best_layers = ax, layers
    return best


class SlabDecomposition:
    """Distribute one particle table over `world` ranks and join the library communicator.

        dec = SlabDecomposition(sim, particles, H_inv, rank, world, axis=1); dec.setup()
        sim.step(...)                       # collective: every rank calls it
        state = dec.gather(order="id")      # rank 0 gets the whole table back
    """

    def __init__(self, sim, particles, H_inv: float, rank: int, world: int, axis: Optional[int] = None,
                 edges: Optional[Sequence[int]] = None, boundary_weight: float = 1.0):
        self.sim, self.parts, self.H_inv, self.rank, self.world = sim, particles, float(H_inv), rank, world
        pos = np.asarray(particles.Position)
        self.axis = best_axis(pos, self.H_inv, world) if axis is None else int(axis)
        self.coords = cell_coord(pos[:, self.axis], self.H_inv)
        self.boundary_weight = float(boundary_weight)
        w = None if boundary_weight == 1.0 else np.where(np.asarray(particles.Type) == 1, 1.0, self.boundary_weight)
        self.edges = list(edges) if edges is not None else plan_edges(self.coords, world, weights=w)
        self.mine = np.nonzero(owner_of(self.coords, self.edges) == rank)[0]
        self.n_owned = int(self.mine.size)

    def broadcast_unique_id(self) -> bytes:
        import torch
        import torch.distributed as dist
        from .simulation import comm_unique_id
        if self.world == 1 and not dist.is_initialized():
            return comm_unique_id()
        box = [comm_unique_id() if self.rank == 0 else None]
        dist.broadcast_object_list(box, src=0, device=torch.device("cuda", torch.cuda.current_device())
                                   if dist.get_backend() == "nccl" else None)
        return box[0]

    def join(self):
        """communicator + slab bounds, no upload"""
        uid = self.broadcast_unique_id()
        self.sim.comm_init(uid, self.rank, self.world, self.axis)
        lo, hi = slab_bounds(self.edges, self.rank)
        self.sim.set_slab(lo, hi)
        return self

    def ghost_node_table(self):
        """(points, particle IDs) of the particles that have a ghost node (nonzero GhostPoints row, Q10),
        ascending ID — the whole table, identical on every rank."""
        gp = np.asarray(self.parts.GhostPoints)
        has = np.any(gp != 0, axis=1)
        ids = np.asarray(self.parts.ID, np.int64)[has]
        o = np.argsort(ids, kind="stable")
        return gp[has][o], ids[o]

    def setup(self):
        self.join()
        if getattr(getattr(self.sim, "params", None), "mdbc", 0):
            self.sim.set_ghost_nodes(*self.ghost_node_table())
        self.sim.upload(self.parts.permuted(self.mine))
        return self

    def imbalance(self) -> float:
        """max / mean of the owned particle counts over the ranks (collective)"""
        import torch.distributed as dist
        n = int(self.sim.num_particles)
        if self.world == 1:
            return 1.0
        counts = [None] * self.world
        dist.all_gather_object(counts, n)
        return max(counts) / (sum(counts) / len(counts))

    def rebalance(self, threshold: float = 1.15) -> bool:
        """Re-plan the slab edges from the CURRENT particle distribution when the ranks' loads have
        drifted apart by more than `threshold` (max / mean), and redistribute.  Collective; goes
        through the host (download, all-gather, upload) — meant for the rare re-plan of a long run,
        not for the step loop, whose migration only ever moves particles between adjacent slabs.
        The simulation clock is carried over.  Returns whether a re-plan happened."""
        import torch.distributed as dist
        from .preprocess import make_particles
        if self.world == 1 or self.imbalance() <= threshold:
            return False
        rep = self.sim.report()
        st = self.sim.download(fields=("Position", "Velocity", "Acceleration", "Density", "ID", "Type", "GroupMarker"))
        parts = [None] * self.world
        dist.all_gather_object(parts, st)
        cat = {k: np.concatenate([p[k] for p in parts]) for k in st}
        coords = cell_coord(cat["Position"][:, self.axis], self.H_inv)
        w = None if self.boundary_weight == 1.0 else np.where(cat["Type"] == 1, 1.0, self.boundary_weight)
        self.edges = plan_edges(coords, self.world, weights=w)
        mine = np.nonzero(owner_of(coords, self.edges) == self.rank)[0]
        mine = mine[np.argsort(cat["ID"][mine], kind="stable")]          # the table is kept in ascending ID order
        sub = make_particles(cat["Position"][mine], cat["Density"][mine], cat["Type"][mine], cat["GroupMarker"][mine],
                             cat["ID"][mine], velocity=cat["Velocity"][mine], dtype=cat["Position"].dtype, sort_by_id=False)
        sub.Acceleration[:] = cat["Acceleration"][mine]
        lo, hi = slab_bounds(self.edges, self.rank)
        self.sim.set_slab(lo, hi)
        self.sim.upload(sub)
        self.sim.set_time(rep["total_time"], rep["iteration"])
        self.n_owned = int(mine.size)
        return True

    def gather(self, order: str = "id", fields=("Position", "Velocity", "Density", "Pressure", "ID")):
        """All ranks' owned particles on rank 0 (None elsewhere)."""
        import torch.distributed as dist
        st = self.sim.download(fields=fields)
        if self.world == 1:
            parts = [st]
        else:
            parts = [None] * self.world if self.rank == 0 else None
            dist.gather_object(st, parts, dst=0)
            if self.rank != 0:
                return None
        out = {k: np.concatenate([p[k] for p in parts]) for k in fields}
        if order == "id":
            o = np.argsort(out["ID"], kind="stable")
            out = {k: v[o] for k, v in out.items()}
        return out
