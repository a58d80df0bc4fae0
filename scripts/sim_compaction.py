"""CPU model of the interaction kernel's two-phase traversal: for a sample of warps (32 consecutive
cell-sorted targets of a row brick) replay the candidate walk and measure, for several list/flush
policies, the lane utilisation of the pair body (phase 2) and the warp-level trip counts."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sphexample_b200 import cases
from oracle import brute_force

dp = float(sys.argv[1]) if len(sys.argv) > 1 else 0.0085
case = cases.case_dam_break_3d(dp, "float64")
pos = case.particles.Position
H = case.kernel.H; H2 = case.kernel.H2
cells = brute_force.cell_coords(pos, 1.0 / H)
cmin = cells.min(0) - 1
ext = cells.max(0) - cells.min(0) + 3
nx, ny, nz = ext
c = cells - cmin
key = (c[:, 2] * ny + c[:, 1]) * nx + c[:, 0]
order = np.argsort(key, kind="stable")
pos = pos[order]; key = key[order]
ncell = nx * ny * nz
cs = np.searchsorted(key, np.arange(ncell + 1))
N = len(pos)
BT = 128
rng = np.random.default_rng(0)
# bricks: rows cut into <= BT targets
rows = ny * nz
bricks = []
for r in range(rows):
    p0, p1 = cs[r * nx], cs[(r + 1) * nx]
    for t0 in range(p0, p1, BT):
        bricks.append((t0, min(t0 + BT, p1)))
print("N", N, "bricks", len(bricks), "mean targets/brick", N / len(bricks))
sample = rng.choice(len(bricks), size=min(400, len(bricks)), replace=False)

def policies():
    return {"cap64_any60": dict(cap=64, mode="any"), "cap32_any28": dict(cap=32, mode="any"),
            "cap128_any124": dict(cap=128, mode="any"), "ring64_thr24": dict(cap=64, mode="ring", thr=24),
            "ring64_thr28": dict(cap=64, mode="ring", thr=28), "ring32_thr24": dict(cap=32, mode="ring", thr=24),
            "unlimited": dict(cap=10**9, mode="any")}

stats = {k: [0, 0] for k in policies()}   # accepted pairs, rounds*32
tests_warp = 0; tests_lane = 0; acc_total = 0; any_ok = 0; warps = 0
for b in sample:
    t0, t1 = bricks[b]
    rowbase = (key[t0] // nx) * nx
    for w0 in range(t0, t1, 32):
        idx = np.arange(w0, min(w0 + 32, t1))
        warps += 1
        cx = key[idx] - rowbase
        seqs = []   # list of boolean [lanes] per candidate in walk order
        for ds in (-1, 0, 1):
            for dm in (-1, 0, 1):
                rk = rowbase + (ds * ny + dm) * nx
                lo = cs[rk + cx - 1]; hi = cs[rk + cx + 2]
                jb, je = lo.min(), hi.max()
                if je <= jb: continue
                j = np.arange(jb, je)
                d = pos[idx][:, None, :] - pos[j][None, :, :]
                r2 = (d * d).sum(2)
                ok = (r2 <= H2) & (j[None, :] >= lo[:, None]) & (j[None, :] < hi[:, None]) & (j[None, :] != idx[:, None])
                tests_warp += len(j); tests_lane += int(((j[None, :] >= lo[:, None]) & (j[None, :] < hi[:, None])).sum())
                any_ok += int(ok.any(0).sum())
                seqs.append(ok.T)   # [cand, lane]
        if not seqs: continue
        ok_all = np.concatenate(seqs, 0)
        acc_total += int(ok_all.sum())
        for name, pol in policies().items():
            cnt = np.zeros(ok_all.shape[1], int)
            rounds = 0
            cap = pol["cap"]
            for g in range(0, len(ok_all), 4):
                cnt += ok_all[g:g + 4].sum(0)
                if cnt.max() > cap - 4:
                    if pol["mode"] == "any":
                        rounds += cnt.max(); cnt[:] = 0
                    else:
                        # ring: pop rounds while >= thr lanes have entries; then keep going until max <= cap/2
                        while True:
                            have = (cnt > 0).sum()
                            if have == 0: break
                            if have < pol["thr"] and cnt.max() <= cap // 2: break
                            rounds += 1; cnt = np.maximum(cnt - 1, 0)
            rounds += cnt.max()
            stats[name][0] += int(ok_all.sum()); stats[name][1] += rounds * 32
print(f"warps {warps}: warp-level candidates/warp {tests_warp / warps:.0f}, in-window tests/lane {tests_lane / warps / 32:.0f}, "
      f"accepted/lane {acc_total / warps / 32:.1f}, P(any lane ok | candidate) {any_ok / tests_warp:.3f}")
for k, (a, r) in stats.items():
    print(f"{k:16s} phase-2 lane utilisation {a / max(r, 1):.3f}")

# ---- interleaved walk: segments take `chunk` candidates of every row in turn ----------------------
def walk_stats(order_fn, label, cap=64, stage=None):
    acc = 0; rounds_tot = 0
    for b in sample:
        t0, t1 = bricks[b]
        rowbase = (key[t0] // nx) * nx
        for w0 in range(t0, t1, 32):
            idx = np.arange(w0, min(w0 + 32, t1))
            cx = key[idx] - rowbase
            rowsok = []
            for ds in (-1, 0, 1):
                for dm in (-1, 0, 1):
                    rk = rowbase + (ds * ny + dm) * nx
                    lo = cs[rk + cx - 1]; hi = cs[rk + cx + 2]
                    jb, je = lo.min(), hi.max()
                    if je <= jb: continue
                    j = np.arange(jb, je)
                    d = pos[idx][:, None, :] - pos[j][None, :, :]
                    r2 = (d * d).sum(2)
                    ok = (r2 <= H2) & (j[None, :] >= lo[:, None]) & (j[None, :] < hi[:, None]) & (j[None, :] != idx[:, None])
                    rowsok.append(ok.T)
            if not rowsok: continue
            seq = order_fn(rowsok)
            cnt = np.zeros(seq.shape[1], int); rounds = 0
            for g in range(0, len(seq), 4):
                cnt += seq[g:g + 4].sum(0)
                if cnt.max() > cap - 4 or (stage and (g // stage) != ((g + 4) // stage)):
                    rounds += cnt.max(); cnt[:] = 0
            rounds += cnt.max()
            acc += int(seq.sum()); rounds_tot += rounds * 32
    print(f"{label:40s} utilisation {acc / max(rounds_tot, 1):.3f}")

def sequential(rowsok): return np.concatenate(rowsok, 0)
def interleave(chunk):
    def f(rowsok):
        out = []; k = 0
        while True:
            got = False
            for r in rowsok:
                seg = r[k:k + chunk]
                if len(seg): out.append(seg); got = True
            if not got: break
            k += chunk
        return np.concatenate(out, 0)
    return f
walk_stats(sequential, "sequential cap64")
walk_stats(sequential, "sequential cap64 stage320", stage=320)
walk_stats(sequential, "sequential cap64 stage1024", stage=1024)
for ch in (4, 8, 16, 32):
    walk_stats(interleave(ch), f"interleave chunk{ch} cap64")
walk_stats(interleave(8), "interleave chunk8 cap32", cap=32)
walk_stats(interleave(8), "interleave chunk8 cap64 stage320", stage=320)
walk_stats(interleave(8), "interleave chunk8 cap48", cap=48)
