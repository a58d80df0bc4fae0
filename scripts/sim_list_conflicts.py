"""CPU model of the list kernel's shared-memory gathers (no GPU needed): how many wavefronts do the
LDS.128 loads of one warp cost for different orderings of the per-particle neighbour lists?

Geometry: a jittered fluid lattice with H = 2*sqrt(3)*dp (the 3D dam-break constants), cell-sorted
like the device table (x fastest), bricks of 128 consecutive particles of a row, window = the 9 row
spans concatenated (4-aligned), list = window indices within H + skin.  Model: a 16-byte element
lives in bank group (index mod 8); the 8 lanes of a quarter warp are served in as many wavefronts as
the largest number of DISTINCT indices that share a bank group (equal indices broadcast).
Measured on B200 (profiles/r1m): 12.5 wavefronts per LDS.128 = 3.1 per quarter warp.

  python scripts/sim_list_conflicts.py [cells_per_axis] [skin]"""
import sys

import numpy as np

nc = int(sys.argv[1]) if len(sys.argv) > 1 else 6
skin = float(sys.argv[2]) if len(sys.argv) > 2 else 0.1
rng = np.random.default_rng(0)
dp = 1.0
H = 2.0 * np.sqrt(3.0) * dp
L = nc * H
g = np.arange(0.5 * dp, L, dp)
pos = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3) + rng.uniform(-0.15, 0.15, (len(g) ** 3, 3)) * dp
cell = np.floor(pos / H).astype(int)
key = (cell[:, 2] * nc + cell[:, 1]) * nc + cell[:, 0]
order = np.argsort(key, kind="stable")
pos, cell, key = pos[order], cell[order], key[order]
n = len(pos)
start = np.searchsorted(key, np.arange(nc ** 3 + 1))
Hs2 = (H * (1 + skin)) ** 2


def quarter_wavefronts(idx8):
    """idx8: window indices of the (active) lanes of one quarter warp at one list position"""
    u = np.unique(idx8)
    return 0 if len(u) == 0 else int(np.bincount(u % 8, minlength=8).max())


def greedy_rotation(lst, q):
    """lane q: reorder so that position k prefers bank group (q + k) % 8; batches of 64 as the build flushes"""
    out = []
    for b0 in range(0, len(lst), 64):
        batch = list(lst[b0:b0 + 64])
        buckets = [[e for e in batch if e % 8 == r] for r in range(8)]
        k0 = len(out)
        for k in range(len(batch)):
            r = (q + k0 + k) % 8
            if not buckets[r]:
                r = int(np.argmax([len(b) for b in buckets]))
            out.append(buckets[r].pop(0))
    return out


tot = {"window order": 0, "greedy rotation": 0}
slots = 0
entries = 0
nbricks = 0
for cz in range(1, nc - 1):
    for cy in range(1, nc - 1):
        row0 = (cz * nc + cy) * nc
        p0, p1 = start[row0 + 1], start[row0 + nc - 1]          # interior cells of the row
        for t0 in range(p0, p1, 128):
            t1 = min(t0 + 128, p1)
            cx0, cx1 = cell[t0, 0], cell[t1 - 1, 0]
            win = []
            for dz in (-1, 0, 1):
                for dy in (-1, 0, 1):
                    rk = ((cz + dz) * nc + cy + dy) * nc
                    w0, w1 = start[rk + cx0 - 1] & ~3, (start[rk + cx1 + 2] + 3) & ~3
                    win.extend(range(w0, min(w1, n)))
            win = np.array(win)
            wpos = pos[win]
            lists = []
            for i in range(t0, t1):
                d2 = ((wpos - pos[i]) ** 2).sum(1)
                ok = (d2 <= Hs2) & (np.abs(cell[win, 0] - cell[i, 0]) <= 1)
                lists.append(np.nonzero(ok)[0])
            nbricks += 1
            for w in range(0, t1 - t0, 32):
                lanes = lists[w:w + 32]
                variants = {"window order": [list(l) for l in lanes],
                            "greedy rotation": [greedy_rotation(l, q % 8) for q, l in enumerate(lanes)]}
                m = max(len(l) for l in lanes)
                slots += m * len(lanes)
                entries += sum(len(l) for l in lanes)
                for name, ls in variants.items():
                    for k in range(m):
                        for qw in range(0, len(ls), 8):
                            idx = [l[k] for l in ls[qw:qw + 8] if k < len(l)]
                            tot[name] += quarter_wavefronts(np.array(idx, dtype=int))
        if nbricks >= 12:
            break
    if nbricks >= 12:
        break
qsteps = slots / 8.0
print(f"bricks {nbricks}, mean list length {entries / max(1, slots) * (slots / max(1, slots)):.2f} (entries {entries}, lane slots {slots})")
for name, v in tot.items():
    print(f"{name:16s}: {v / qsteps:5.2f} wavefronts per quarter-warp load (ideal 1.00; measured on B200 with window order: 3.1)")
