"""Instruction histogram of the shipped library's SASS (cuobjdump -sass): which tensor-core / TMEM / TMA / bulk-copy
mnemonics the sm_100a kernels contain, per kernel family.   python profiles/sass_histogram.py > profiles/sass_r2.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "cgl-gan_b200", "lib", "libcgl_b200.so")
WATCH = ["UTCHMMA", "UTCQMMA", "UTCBAR", "UTCCP", "LDTM", "STTM", "UTCATOMSWS", "UTMALDG", "UTMASTG", "UBLKCP", "UBLKPF",
         "LDGSTS", "SYNCS", "UCGABAR", "MUFU", "FFMA", "HMMA", "LDG", "STG", "LDS", "STS", "REDUX", "ATOMG", "ATOMS"]

out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
fam = collections.defaultdict(collections.Counter)
ninstr = collections.Counter()
cur = None
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"<.*", "", name).replace("cgl::", "").replace("(anonymous namespace)::", "")
        cur = re.sub(r"\(.*", "", cur).split()[-1]
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        op = m.group(1).split(".")[0]
        ninstr[cur] += 1
        for w in WATCH:
            if op == w or (w in ("UTCHMMA", "LDTM", "STTM", "UBLKPF", "UBLKCP", "UTCBAR") and op.startswith(w)):
                fam[cur][w] += 1
print(f"# SASS instruction histogram of `{os.path.relpath(LIB, ROOT)}` (cuobjdump -sass, all template instances summed)\n")
tot = collections.Counter()
for k in fam:
    tot.update(fam[k])
print("whole library: " + ", ".join(f"{w} {tot[w]}" for w in WATCH if tot[w]) + "\n")
print("| kernel family | instructions | " + " | ".join(WATCH[:16]) + " |")
print("|---|---:|" + "---:|" * 16)
for k in sorted(ninstr, key=lambda n: -ninstr[n]):
    print(f"| `{k}` | {ninstr[k]} | " + " | ".join(str(fam[k][w]) if fam[k][w] else "" for w in WATCH[:16]) + " |")
