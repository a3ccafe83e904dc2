"""Summarise an `ncu --page source --csv` dump by __syncthreads phase: instruction share,
stall-sample share and top opcodes.  Usage: python profiles/sass_phases.py dump.csv"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
iS, iE, iSm = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
phase, tot, totS = 0, 0, 0
ph = collections.defaultdict(lambda: [0, 0, collections.Counter(), collections.Counter()])
for r in rows[hi + 1:]:
    if len(r) <= iE or not r[iE].strip().isdigit():
        continue
    src, n, s = r[iS], int(r[iE]), int(r[iSm] or 0)
    toks = src.split()
    op = toks[1] if toks[0].startswith("@") else toks[0]
    ph[phase][0] += n; ph[phase][1] += s
    ph[phase][2][op.split(".")[0]] += n
    ph[phase][3][op.split(".")[0]] += s
    tot += n; totS += s
    if "BAR.SYNC" in src:
        phase += 1
print("total warp-instructions", tot, "stall samples", totS)
for p, (n, s, c, cs) in sorted(ph.items()):
    print(f"phase {p}: inst {n} ({n / tot:.1%})  samples {s} ({s / max(totS, 1):.1%})")
    print("    inst   :", c.most_common(10))
    print("    samples:", cs.most_common(6))
