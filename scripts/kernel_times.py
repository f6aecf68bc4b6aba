"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` log: mean duration per kernel name, in launch order of first appearance.
   python scripts/kernel_times.py gpurun_out/x.csv"""
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value"); ui = hdr.index("Metric Unit")
t = collections.OrderedDict()
for r in rows[1:]:
    v = float(r[vi].replace(",", "")); v = v / 1000.0 if r[ui] == "ns" else v
    t.setdefault(r[ki].split("(")[0][-60:], []).append(v)
for k, v in t.items(): print(f"{k:62s} n={len(v):4d} mean {sum(v)/len(v):9.2f} us  min {min(v):9.2f}  max {max(v):9.2f}")
