"""Host logic of the multi-GPU slab decomposition (sphexample_b200/slab.py), on CPU.
The N > 1 plumbing (unique-id broadcast, partition, gather) runs under torch.distributed with the
gloo backend, world_size 2; the device side is covered by scripts/slab_parity.py on 2 GPUs."""
import os
import socket

import numpy as np
import pytest

import util
from sphexample_b200 import slab


def test_cell_coord_is_round_half_away_from_zero():
    H_inv = 1.0 / 0.08
    x = np.array([0.0, 0.039, 0.04, 0.041, -0.039, -0.04, -0.041, 0.12, -0.12])
    assert slab.cell_coord(x, H_inv).tolist() == [0, 0, 1, 1, 0, -1, -1, 2, -2]   # src/SPHCellList.jl:56-61


@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_plan_edges_balances_by_particle_count(world):
    case = util.case_3d_small("float64")
    H_inv = util.params_of(case).H_inv
    c = slab.cell_coord(case.particles.Position[:, 1], H_inv)
    if c.max() - c.min() + 1 < 2 * world:
        with pytest.raises(ValueError):
            slab.plan_edges(c, world)
        return
    e = slab.plan_edges(c, world)
    assert len(e) == world + 1 and e[0] == c.min() and e[-1] == c.max() + 1
    assert all(b - a >= 2 for a, b in zip(e, e[1:]))          # protocol needs >= 2 layers per slab
    own = slab.owner_of(c, e)
    counts = np.bincount(own, minlength=world)
    assert counts.sum() == len(c) and counts.min() > 0
    # every particle's layer lies inside its owner's [lo, hi)
    for r in range(world):
        lo, hi = slab.slab_bounds(e, r)
        assert np.all((c[own == r] >= lo) & (c[own == r] < hi))
    # balance: no rank above 1.6x the mean on this coarse (11-layer) case
    assert counts.max() <= 1.6 * len(c) / world + 1


def test_skewed_distribution_is_not_split_by_width():
    # dam break: most particles in the first quarter -> equal-width slabs would be unusable
    c = np.concatenate([np.repeat(np.arange(0, 10), 1000), np.repeat(np.arange(10, 40), 10)])
    e = slab.plan_edges(c, 4)
    counts = np.bincount(slab.owner_of(c, e), minlength=4)
    assert counts.max() < 0.4 * len(c)
    assert e[1] < 10 and e[2] < 10


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


class _FakeSim:
    """records what SlabDecomposition hands to the library (no GPU here)"""

    def __init__(self):
        self.calls = []

    def comm_init(self, uid, rank, world, axis):
        self.calls.append(("comm_init", bytes(uid), rank, world, axis))

    def set_slab(self, lo, hi):
        self.calls.append(("set_slab", lo, hi))

    def upload(self, parts):
        self.uploaded = parts

    def download(self, fields):
        p = self.uploaded
        return {k: np.asarray(getattr(p, k)) for k in fields}

    @property
    def num_particles(self):
        return len(self.uploaded)

    def report(self):
        return {"total_time": 0.25, "iteration": 77}

    def set_time(self, t, it):
        self.calls.append(("set_time", t, it))


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from sphexample_b200 import simulation
        simulation.comm_unique_id = lambda: bytes(range(128)) if rank == 0 else b"\xff" * 128   # no NCCL on CPU
        case = util.case_3d_small("float32")
        H_inv = util.params_of(case).H_inv
        sim = _FakeSim()
        dec = slab.SlabDecomposition(sim, case.particles, H_inv, rank, world, axis=1).setup()
        uid = sim.calls[0][1]
        st = dec.gather(order="id", fields=("Position", "Density", "ID"))
        # drift: push a third of rank 0's particles across the slab face, then re-plan from the current state
        before = dec.imbalance()
        if rank == 0:
            up = sim.uploaded
            k = len(up) // 3
            up.Position[:k, 1] += (dec.edges[-1] - dec.edges[0]) * 0.6 / H_inv
        moved = dec.rebalance(threshold=1.0)
        after = dec.imbalance()
        st2 = dec.gather(order="id", fields=("ID",))
        ok2 = None if st2 is None else (st2["ID"].tolist() == np.sort(case.particles.ID).tolist())
        in_slab = bool(np.all(slab.owner_of(slab.cell_coord(sim.uploaded.Position[:, 1], H_inv), dec.edges) == rank))
        q.put((rank, uid, sim.calls[1], dec.n_owned, (before, moved, after, ok2, in_slab, sim.calls[-1]), None if st is None else
               (st["ID"].tolist() == np.sort(case.particles.ID).tolist(),
                bool(np.array_equal(st["Position"], case.particles.Position[np.argsort(case.particles.ID, kind="stable")])))))
    finally:
        dist.destroy_process_group()


def test_two_ranks_partition_broadcast_and_gather_with_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, uid0, slab0, n0, reb0, chk0), (r1, uid1, slab1, n1, reb1, chk1) = res
    # rebalance: collective re-plan keeps every particle, puts each rank's share inside its new slab, carries the clock
    assert reb0[1] is True and reb1[1] is True and reb0[3] is True and reb0[4] and reb1[4]
    assert reb0[2] <= 1.25 and reb0[5] == ("set_time", 0.25, 77) and reb1[5] == ("set_time", 0.25, 77)
    assert uid0 == uid1 == bytes(range(128))                  # rank 0's id reached rank 1
    assert slab0[1] == slab.INT64_MIN and slab1[2] == slab.INT64_MAX and slab0[2] == slab1[1]   # adjacent slabs
    case = util.case_3d_small("float32")
    assert min(n0, n1) > 0
    assert chk0 == (True, True) and chk1 is None              # rank 0 got the whole table back, by ID


def test_ghost_node_table_is_the_nonzero_ghost_rows_by_ascending_id():
    """slab-mode SimpleMDBC hands every rank the same static node table (sphb200_set_ghost_nodes)"""
    case = util.case_c5("float64")
    p = case.particles
    shuffled = p.permuted(np.random.default_rng(3).permutation(len(p)))
    dec = slab.SlabDecomposition(_FakeSim(), shuffled, util.params_of(case).H_inv, 0, 2, axis=0)
    gp, gid = dec.ghost_node_table()
    has = np.any(p.GhostPoints != 0, axis=1)
    assert gp.shape == (int(has.sum()), 2) and gp.shape[0] > 500
    assert np.all(np.diff(gid) > 0) and np.array_equal(gid, np.sort(p.ID[has]))
    o = np.argsort(p.ID[has], kind="stable")
    assert np.array_equal(gp, p.GhostPoints[has][o])


def _mdbc_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from sphexample_b200 import simulation
        simulation.comm_unique_id = lambda: bytes(range(128)) if rank == 0 else b"\xff" * 128
        case = util.case_c5("float64")
        p = util.params_of(case)

        class Sim(_FakeSim):
            params = p

            def set_ghost_nodes(self, gp, ids):
                self.calls.append(("set_ghost_nodes", np.asarray(gp).tobytes(), np.asarray(ids).tobytes(), len(ids)))
        sim = Sim()
        dec = slab.SlabDecomposition(sim, case.particles, p.H_inv, rank, world, axis=0).setup()
        names = [c[0] for c in sim.calls]
        own_ghosts = int(np.any(sim.uploaded.GhostPoints != 0, axis=1).sum())
        q.put((rank, names, sim.calls[names.index("set_ghost_nodes")][1:], dec.n_owned, own_ghosts))
    finally:
        dist.destroy_process_group()


def test_every_rank_hands_over_the_same_ghost_node_table_with_gloo():
    """slab-mode SimpleMDBC: comm_init, set_slab, then the WHOLE ghost-node table on every rank (not just the nodes of
    the rank's own particles), then the rank's share of the particles"""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_mdbc_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, names0, tab0, n0, g0), (r1, names1, tab1, n1, g1) = res
    assert names0 == names1 == ["comm_init", "set_slab", "set_ghost_nodes"]
    assert tab0 == tab1 and tab0[2] > 500                      # identical bytes on both ranks
    assert n0 + n1 == 3027 and g0 + g1 == tab0[2] and 0 < g0 < tab0[2]   # each rank owns only a part of the nodes' particles
