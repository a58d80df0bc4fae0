"""Per-kernel fingerprint of the compiled library: sha1 of the SASS instruction stream (addresses and
encodings stripped).  Two builds with equal fingerprints execute the same device code, so a kernel
validated on hardware at commit A is known to be untouched at commit B without a GPU.
  python scripts/sass_fingerprint.py [libsphb200.so] > fingerprints.txt ; diff them"""
import hashlib
import re
import subprocess
import sys

so = sys.argv[1] if len(sys.argv) > 1 else "sphexample_b200/lib/libsphb200.so"
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
cur, acc = None, {}
for line in txt.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        acc[cur] = []
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(.*?;)", line)
    if m and cur:
        acc[cur].append(re.sub(r"\s+", " ", m.group(1)))
dem = subprocess.run(["c++filt"], input="\n".join(acc), capture_output=True, text=True).stdout.splitlines()
for name, d in sorted(zip(acc, dem), key=lambda x: x[1]):
    h = hashlib.sha1("\n".join(acc[name]).encode()).hexdigest()[:12]
    short = re.sub(r"\(.*", "", d).replace("void ", "").replace("sph::", "").replace("(anonymous namespace)::", "")
    print(f"{h} {len(acc[name]):6d}  {short}")
