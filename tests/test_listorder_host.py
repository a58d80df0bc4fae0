"""Bank-aware ("rainbow") ordering of neighbour-list entries (csrc/sph_listorder.h) on the CPU: the
product's own code through the shim.  It must hand back every real entry exactly once (a
permutation: nothing lost, nothing doubled — the physics only sees a different summation order),
pad only with sentinels of the bank group the slot expects, and put slot u of every chunk of lane
q into bank group (q + u) mod 8 unless the entry comes from an over-full group."""
import ctypes as C

import numpy as np
import pytest

from test_physics_host import shim  # noqa: F401  (fixture)


def order(shim, entries, q, total8, ovf_cap=32):  # noqa: F811
    e = np.ascontiguousarray(entries, np.uint16)
    assert len(e) % 8 == 0
    out = np.full_like(e, 0xFFFF)
    shim.shim_rainbow_order.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_uint, C.c_int, C.c_void_p]
    ok = shim.shim_rainbow_order(e.ctypes.data, len(e), q, total8, ovf_cap, out.ctypes.data)
    return bool(ok), out


def padded(idx, roles, total8):
    ent = (idx.astype(np.uint16) | (roles.astype(np.uint16) << 15))
    pad = (-len(ent)) % 8
    return np.concatenate([ent, np.full(pad, total8, np.uint16)])


@pytest.mark.parametrize("m", [0, 1, 7, 8, 9, 33, 64, 190, 317])
def test_every_entry_comes_back_exactly_once(shim, m):  # noqa: F811
    rng = np.random.default_rng(m)
    total8 = 2000
    for q in range(8):
        idx = rng.choice(total8, size=m, replace=False)
        ent = padded(idx, rng.integers(0, 2, m), total8)          # role bit rides along
        ok, out = order(shim, ent, q, total8, ovf_cap=256)
        assert ok
        real = out[(out & 0x7fff) < total8]
        assert sorted(real.tolist()) == sorted(ent[(ent & 0x7fff) < total8].tolist())
        # the padding is the sentinel of the bank group the slot expects
        for k, e in enumerate(out):
            if (e & 0x7fff) >= total8:
                assert e == total8 + (q + k) % 8


def test_skewed_lists_survive_or_report(shim):  # noqa: F811
    total8 = 4096
    for q in range(8):
        ent = (np.arange(64, dtype=np.uint16) * 8 + 3)            # all in one bank group: 56 overflow entries
        ok, out = order(shim, ent, q, total8, ovf_cap=64)
        assert ok and sorted(out.tolist()) == sorted(ent.tolist())
        ok, _ = order(shim, ent, q, total8, ovf_cap=32)           # scratch too small: reported, list kept as built
        assert not ok


def test_slots_follow_the_lane_rotation(shim):  # noqa: F811
    rng = np.random.default_rng(3)
    hits = total = 0
    total8 = 2000
    for q in range(8):
        idx = np.sort(rng.choice(total8, size=192, replace=False))
        ok, out = order(shim, padded(idx, np.zeros(192, int), total8), q, total8)
        assert ok
        want = (q + np.arange(len(out))) % 8
        hits += int((((out & 0x7fff) & 7) == want).sum())
        total += len(out)
    assert hits / total > 0.9


def test_quarter_warp_conflicts_drop(shim):  # noqa: F811
    """8 lanes with overlapping random lists: wavefronts per load (largest number of distinct indices
    per bank group) before and after"""
    rng = np.random.default_rng(11)
    total8 = 1800

    def wavefronts(lists):
        w = 0
        n = max(len(l) for l in lists)
        for k in range(n):
            u = np.unique([l[k] & 0x7fff for l in lists if k < len(l)])
            w += np.bincount(u % 8, minlength=8).max()
        return w / n

    before = after = 0.0
    for trial in range(20):
        base = rng.choice(total8, size=300, replace=False)
        lists = [padded(np.sort(rng.choice(base, size=int(rng.integers(150, 200)), replace=False)), np.zeros(1, int), total8)
                 for _ in range(8)]
        before += wavefronts(lists)
        after += wavefronts([order(shim, l, q, total8)[1] for q, l in enumerate(lists)])
    assert after < 0.6 * before      # (lattice model, scripts/sim_list_conflicts.py: 9.3 -> 4.4 wavefronts per LDS.128)
