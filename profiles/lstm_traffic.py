"""Write profiles/r02_lstm_traffic.json from an `ncu --page raw --csv` dump of the `--set full` capture of tests/probes/lstm_ncu.py
(one lstm_fwd_kernel + one lstm_bwd_kernel launch at the benched shape).  bench.py reads the JSON for `roofline.traffic`.
usage: ncu -i lstm_full.ncu-rep --page raw --csv > raw.csv; python profiles/lstm_traffic.py raw.csv profiles/<committed copy of raw.csv>"""
import csv, json, os, sys

raw, committed = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else sys.argv[1])
rows = list(csv.reader(open(raw)))
hdr, units = rows[0], rows[1]
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
dur = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}
per = {}
for r in rows[2:]:
    d = dict(zip(hdr, r))
    name = "lstm_fwd_kernel" if "lstm_fwd" in d["Kernel Name"] else "lstm_bwd_kernel" if "lstm_bwd" in d["Kernel Name"] else None
    if not name:
        continue
    def val(k):
        return float(d[k].replace(",", "")) * scale.get(units[hdr.index(k)], 1.0)
    tp = [d[k] for k in hdr if k.endswith("sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed") and d[k] not in ("", "no data")]
    per[name] = {"dram_read": val("dram__bytes_read.sum"), "dram_write": val("dram__bytes_write.sum"),
                 "duration_us_under_ncu": round(float(d["gpu__time_duration.sum"].replace(",", "")) * dur.get(units[hdr.index("gpu__time_duration.sum")], 1.0), 1),
                 "tensor_pipe_active_pct": round(float(tp[0]), 1) if tp else None,
                 "issue_active_pct": round(float(d["sm__issue_active.avg.pct_of_peak_sustained_elapsed"]), 1) if "sm__issue_active.avg.pct_of_peak_sustained_elapsed" in d else None}
mean = sum(v["dram_read"] + v["dram_write"] for v in per.values()) / max(1, len(per))
out = {"c2": {"dram_bytes_per_launch": mean, "per_kernel": per,
              "algorithmic_bytes": {"lstm_fwd_kernel": "P 262.1 MB read; Y 65.5 + C 131.1 + activated gates 262.1 MB written = 720.9 MB (B=64, T=500, H=512)",
                                    "lstm_bwd_kernel": "gates 262.1 + C 131.1 + dY 65.5 MB read; dA 262.1 MB written = 720.9 MB"},
              "source": f"{committed}: ncu --set full --clock-control none of tests/probes/lstm_ncu.py at the benched shape (B=64, T=500, In=64, H=512), "
                        "dram__bytes_read.sum + dram__bytes_write.sum, mean of the forward and the backward launch"}}
p = os.path.join(os.path.dirname(os.path.abspath(__file__)), "r02_lstm_traffic.json")
json.dump(out, open(p, "w"), indent=1)
print(json.dumps(out, indent=1))
