"""Top stalled SASS lines of one kernel from `ncu -i rep --page source --csv --kernel-id ::regex:NAME:N > src.csv`.
    python profiles/top_stalls.py src.csv [n]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = []
hdr = None
for r in rows:
    if r and r[0] == "Address":
        hdr = r
        continue
    if hdr is None or len(r) != len(hdr):
        continue
    out.append(r)
iS = hdr.index("Warp Stall Sampling (All Samples)")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[iS] or 0) for r in out)
print("total samples", tot, " instructions", len(out))
agg = {}
for r in out:
    for i in stall_cols:
        agg[hdr[i]] = agg.get(hdr[i], 0) + int(r[i] or 0)
print("by reason:", sorted(((v, k) for k, v in agg.items() if v), reverse=True)[:8])
for idx, r in sorted(enumerate(out), key=lambda t: -int(t[1][iS] or 0))[:n]:
    reasons = sorted(((int(r[i] or 0), hdr[i][6:]) for i in stall_cols), reverse=True)[:2]
    print(f"{int(r[iS]):6d} #{idx:4d} {r[1].strip()[:64]:64s} {reasons}")
