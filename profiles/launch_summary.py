"""Summarise ONE training step out of an `ncu --metrics gpu__time_duration.sum --csv` launch list of bench.py (eager, --no-graph):
the launches from the last front-end kernel (logmel) up to the next one / the end, grouped by kernel name.
usage: python profiles/launch_summary.py launches.csv [occurrence_from_end=2]"""
import csv, re, sys
from collections import OrderedDict

path = sys.argv[1]
back = int(sys.argv[2]) if len(sys.argv) > 2 else 2
rows = []
with open(path, newline="") as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") == "gpu__time_duration.sum":
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        us = v / 1e3 if unit in ("ns", "nsecond") else v if unit in ("us", "usecond") else v * 1e3
        rows.append((r["Kernel Name"], us))
starts = [i for i, (k, _) in enumerate(rows) if "logmel" in k]
s = starts[-back]
e = starts[-back + 1] if back > 1 else len(rows)
# the step starts at the memset / fill in front of the logmel kernel
while s > 0 and ("Memset" in rows[s - 1][0] or "memset" in rows[s - 1][0]):
    s -= 1
step = rows[s:e]
# bench.py writes a 256 MiB buffer between timed steps (L2 flush, outside the event pairs): not part of the step
flush = [r for r in step if "FillFunctor<unsigned char>" in r[0]]
step = [r for r in step if "FillFunctor<unsigned char>" not in r[0]]
agg = OrderedDict()
for k, us in step:
    short = re.sub(r"\(.*", "", k).replace("void ", "").replace("(anonymous namespace)::", "")
    short = re.sub(r"unnamed>::", "", short)[:110]
    a = agg.setdefault(short, [0.0, 0])
    a[0] += us
    a[1] += 1
tot = sum(us for _, us in step)
if flush:
    print(f"(excluded: {len(flush)} L2-flush fill of bench.py between steps, {sum(us for _, us in flush):.1f} us)")
print(f"one training step: {len(step)} launches, sum of kernel durations {tot / 1e3:.3f} ms (serialised under ncu, cold caches)\n")
for k, (us, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"  {us:9.1f} us  {100 * us / tot:5.1f} %  x{n:<3d} {k}")
ours = sum(us for k, us in step if "mlvae" in k)
n_ours = sum(1 for k, _ in step if "mlvae" in k)
lib = sum(us for k, us in step if any(t in k for t in ("nvjet", "cutlass", "cudnn", "RNN_", "gemv", "cublas")))
print(f"\nkernels of libmlvae_b200.so: {ours / 1e3:.3f} ms = {100 * ours / tot:.1f} % of the step ({n_ours} launches); "
      f"cuBLAS / cuDNN kernels: {lib:.1f} us; other (torch elementwise / fill / copy): {len(step) - n_ours} launches, {tot - ours - lib:.1f} us")
