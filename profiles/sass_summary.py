"""Per-kernel counts of the SASS mnemonics that prove the Blackwell-native paths (B200_PROFILING.md):
UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG/UTMAREDG/UBLKCP = TMA, HMMA = mma.sync (legacy tensor
path), LDGSTS = cp.async, SYNCS = mbarrier, FFMA2 = packed fp32 FMA.   python profiles/sass_summary.py > profiles/sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "pinthememory_b200", "_lib", "libpinmem_b200.so")
PAT = [("UTC.*MMA", r"\bUTC\w*MMA"), ("LDTM", r"\bLDTM"), ("STTM", r"\bSTTM"), ("UTMALDG", r"\bUTMALDG"), ("UTMASTG", r"\bUTMASTG"),
       ("UTMAREDG", r"\bUTMAREDG"), ("UBLKCP", r"\bUBLKCP"), ("HMMA", r"\bHMMA"), ("LDGSTS", r"\bLDGSTS"), ("SYNCS", r"\bSYNCS"),
       ("FFMA2", r"\bFFMA2"), ("MUFU", r"\bMUFU"), ("RED/ATOM", r"\b(RED|ATOMG|ATOMS)\b")]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    counts, order, cur = collections.defaultdict(collections.Counter), [], None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = re.sub(r"\(.*", "", cur).replace("void ", "")
            order.append(cur)
            continue
        if cur is None:
            continue
        for name, pat in PAT:
            if re.search(pat, line):
                counts[cur][name] += 1
    names = [n for n, _ in PAT]
    print("# cuobjdump -sass pinthememory_b200/_lib/libpinmem_b200.so (sm_100a): instruction counts per kernel")
    print("%-78s " % "kernel" + " ".join("%8s" % n for n in names))
    tot = collections.Counter()
    for k in order:
        if not any(counts[k].values()):
            continue
        print("%-78s " % k[:78] + " ".join("%8d" % counts[k][n] for n in names))
        tot.update(counts[k])
    print("%-78s " % "TOTAL" + " ".join("%8d" % tot[n] for n in names))


if __name__ == "__main__":
    sys.exit(main())
