"""Bank-aware ordering of neighbour-list entries (csrc/sph_listorder.h) on the CPU: the product's
own code through the shim.  It must hand back every entry exactly once (a permutation: nothing
lost, nothing doubled — the physics only sees a different summation order) and it should put lane
q's k-th entry into bank group (q + k) mod 8 whenever that group still has entries."""
import ctypes as C

import numpy as np
import pytest

from test_physics_host import shim  # noqa: F401  (fixture)


def rotate(shim, entries, q):  # noqa: F811
    e = np.ascontiguousarray(entries, np.uint16)
    out = np.zeros_like(e)
    shim.shim_bank_rotate.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    assert shim.shim_bank_rotate(e.ctypes.data, len(e), q, out.ctypes.data) == len(e)
    return out


@pytest.mark.parametrize("m", [0, 1, 7, 8, 9, 33, 63, 64])
def test_every_entry_comes_back_exactly_once(shim, m):  # noqa: F811
    rng = np.random.default_rng(m)
    for q in range(8):
        idx = rng.choice(2000, size=m, replace=False).astype(np.uint16)
        ent = idx | (rng.integers(0, 2, m).astype(np.uint16) << 15)       # role bit rides along
        out = rotate(shim, ent, q)
        assert sorted(out.tolist()) == sorted(ent.tolist())


def test_skewed_batches_survive(shim):  # noqa: F811
    for q in range(8):
        ent = (np.arange(64, dtype=np.uint16) * 8 + 3)                    # all in one bank group
        assert sorted(rotate(shim, ent, q).tolist()) == sorted(ent.tolist())
        ent = np.concatenate([np.arange(60, dtype=np.uint16) * 8, np.array([1, 2, 3, 4], np.uint16)])
        assert sorted(rotate(shim, ent, q).tolist()) == sorted(ent.tolist())


def test_positions_follow_the_lane_rotation_while_groups_last(shim):  # noqa: F811
    rng = np.random.default_rng(3)
    hits = total = 0
    for q in range(8):
        ent = np.sort(rng.choice(2000, size=64, replace=False)).astype(np.uint16)
        out = rotate(shim, ent, q)
        want = (q + np.arange(64)) % 8
        hits += int(((out & 7) == want).sum())
        total += 64
        # stable inside a bank group: window order is kept among the entries of one group
        for r in range(8):
            grp = out[(out & 7) == r]
            assert np.all(np.diff(grp.astype(int)) > 0)
    assert hits / total > 0.7


def test_quarter_warp_conflicts_drop(shim):  # noqa: F811
    """8 lanes with overlapping random lists: wavefronts per load (largest number of distinct indices
    per bank group) before and after"""
    rng = np.random.default_rng(11)

    def wavefronts(lists):
        w = 0
        for k in range(64):
            u = np.unique([l[k] & 0x7fff for l in lists])
            w += np.bincount(u % 8, minlength=8).max()
        return w / 64.0

    before = after = 0.0
    for trial in range(20):
        base = rng.choice(1800, size=120, replace=False)
        lists = [np.sort(rng.choice(base, size=64, replace=False)).astype(np.uint16) for _ in range(8)]
        before += wavefronts(lists)
        after += wavefronts([rotate(shim, l, q) for q, l in enumerate(lists)])
    assert after < 0.8 * before      # (the lattice model of scripts/sim_list_conflicts.py: 2.5 -> 1.3)
