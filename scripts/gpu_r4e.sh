#!/bin/bash
# per-kernel durations of the small cases with the lean sequence + split lists (plain launches under ncu)
O=gpurun_out/r4e; mkdir -p $O
SMALL_PROFILE_NCU=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $O/small_launches.csv python scripts/small_profile.py c1 c5 > $O/ncu.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(l for l in open("gpurun_out/r4e/small_launches.csv") if l.startswith('"'))]
hdr = rows[0]; ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
seq = [(r[ki][:70], float(r[vi].replace(",", ""))) for r in rows[1:]]
print(len(seq), "launches")
for lo, hi, name in ((0, 500, "C1"), (len(seq) - 400, len(seq), "C5")):
    print(name)
    agg = collections.OrderedDict()
    for k, v in seq[lo:hi]:
        agg.setdefault(k, []).append(v)
    for k, v in agg.items():
        v.sort(); print(f"{k:72s} n={len(v):4d} median={v[len(v)//2]/1e3:8.2f} us max={v[-1]/1e3:8.2f}")
PY
