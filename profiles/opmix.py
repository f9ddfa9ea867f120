"""Instruction mix (warp-instructions executed by opcode) from `ncu --page source --csv` output."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
idx = {h: i for i, h in enumerate(hdr)}
seen = set()
mix = collections.Counter()
for r in rows[2:]:
    if len(r) != len(hdr) or not r[idx["# Samples"]].isdigit():
        continue
    key = r[idx["Address"]]
    if key in seen:
        continue
    seen.add(key)
    src = r[idx["Source"]].strip()
    toks = src.split()
    op = toks[1] if toks[0].startswith("@") else toks[0]
    mix[op.split(".")[0] + ("." + op.split(".")[1] if op.startswith(("LDS", "STS", "LDG", "STG", "LDGSTS")) and "." in op else "")] += int(r[idx["Instructions Executed"]])
tot = sum(mix.values())
print("total warp-instructions", tot)
for op, n in mix.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 30):
    print("  %-14s %12d %5.1f%%" % (op, n, 100.0 * n / tot))
