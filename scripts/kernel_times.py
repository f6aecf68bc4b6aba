"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` log: launches, total and mean duration and share per kernel.
   python scripts/kernel_times.py gpurun_out/x.csv [out_summary.csv]"""
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value"); ui = hdr.index("Metric Unit")
t = collections.OrderedDict()
for r in rows[1:]:
    v = float(r[vi].replace(",", "")); v = v / 1000.0 if r[ui] in ("ns", "nsecond") else v
    name = r[ki].split("(")[0].replace("void ", "").replace("las2::", "").replace("<unnamed>::", "").replace("unnamed>::", "").strip()
    t.setdefault(name, []).append(v)
tot = sum(sum(v) for v in t.values())
lines = ["kernel,launches,total_us,share_pct,avg_us"]
for k, v in sorted(t.items(), key=lambda kv: -sum(kv[1])):
    lines.append(f"{k},{len(v)},{sum(v):.1f},{100 * sum(v) / tot:.2f},{sum(v) / len(v):.2f}")
print("\n".join(lines))
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write("\n".join(lines) + "\n")
