"""Attribute `ncu --page source --csv` stall samples of one kernel to CUDA source lines.

    python profiles/sass_lines.py <ncu_source.csv> <object-or-cubin> <kernel-substring> [top]

ncu's CLI source page is SASS only; the line table comes from `nvdisasm -g` on the cubin
(built with -lineinfo).  Instructions are matched by their order inside the function.
"""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

csv_path, obj, kname = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
rows = list(csv.reader(open(csv_path)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
iS, iSm, iE = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
sass = [(int(r[iSm] or 0), int(r[iE] or 0), r[iS].strip()) for r in rows[hi + 1:] if len(r) > iE and r[iE].strip().isdigit()]
# ncu lists the kernel once per captured launch: keep the first copy
first = sass[0][2]
n = next((k for k in range(1, len(sass)) if sass[k][2] == first and sass[k:k + 5] and [x[2] for x in sass[k:k + 5]] == [x[2] for x in sass[:5]]), len(sass))
sass = sass[:n]

tmp = tempfile.mkdtemp()
cubin = obj
if not obj.endswith(".cubin"):
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
    cubin = os.path.join(tmp, sorted(f for f in os.listdir(tmp) if f.endswith(".cubin"))[0])
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
lines, cur, infunc = [], None, False
for ln in dis:
    if ln.startswith(".text.") or re.match(r"\s*\.section\s+\.text\.", ln):
        infunc = kname in ln
        continue
    if not infunc:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
        lines.append(cur)
if len(lines) != len(sass):
    print(f"warning: {len(lines)} instructions in nvdisasm vs {len(sass)} in the ncu dump; matching by order")
agg = collections.defaultdict(lambda: [0, 0])
for (s, e, _), loc in zip(sass, lines):
    agg[loc][0] += s
    agg[loc][1] += e
tot = sum(v[0] for v in agg.values()) or 1
srcs = {}
print(f"total samples {tot}, instructions executed {sum(v[1] for v in agg.values())}")
for loc, (s, e) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    text = ""
    if loc:
        path = next((os.path.join(d, loc[0]) for d, _, fs in os.walk(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))) if loc[0] in fs), None)
        if path:
            srcs.setdefault(path, open(path).read().splitlines())
            text = srcs[path][loc[1] - 1].strip()[:90] if loc[1] - 1 < len(srcs[path]) else ""
    print(f"{s:7d} {s / tot:6.1%} exec {e:11d}  {loc[0] if loc else '?'}:{loc[1] if loc else 0}  {text}")
