"""Aggregate `ncu --page source --csv` (SASS view) of one kernel: stall reasons and top instructions.

    ncu -i prof.ncu-rep --page source --csv --kernel-name regex:<name> > src.csv
    python profiles/stalls.py src.csv [top_n]
"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
print(rows[0][1][:140])
hdr = rows[1]
idx = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) == len(hdr) and r[idx["# Samples"]].isdigit()]
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = {h: 0 for h in stall_cols}
samples = 0
insts = 0
for d in data:
    for h in stall_cols:
        tot[h] += int(d[idx[h]] or 0)
    samples += int(d[idx["# Samples"]] or 0)
    insts += int(d[idx["Instructions Executed"]] or 0)
print("samples", samples, "warp-instructions", insts)
for h, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    if v:
        print("  %-26s %7d %5.1f%%" % (h, v, 100.0 * v / max(samples, 1)))
print("top instructions by samples:")
for d in sorted(data, key=lambda d: -int(d[idx["# Samples"]] or 0))[:topn]:
    st = sorted(((int(d[idx[h]] or 0), h) for h in stall_cols), reverse=True)[:2]
    print("  %6s %9s  %-60s %s" % (d[idx["# Samples"]], d[idx["Instructions Executed"]], d[idx["Source"]].strip()[:60],
                                  ", ".join("%s=%d" % (h[6:], v) for v, h in st if v)))
